// build.cu — K5 GPU Vamana build (placeholder until the batched build lands).
#include "common.cuh"

int launch_vamana_build(const float *d_X, int64_t N, int D, int R, int L, float alpha, int64_t medoid, uint64_t seed,
                        uint32_t *d_out_adj, int32_t *d_out_deg, int device, cudaStream_t s) {
    dr_set_error("dr_vamana_build: not implemented yet");
    return 4;
}
