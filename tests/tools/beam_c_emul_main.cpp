// TEST INFRASTRUCTURE: host driver around the text of csrc/beam_c.cu's kernel (beam_c_kernel.inc is cut out of the .cu file and
// common.cuh by tests/test_beam_c_emulated.py), executed by 32 lockstep threads (warp_emul.h).
#include "warp_emul.h"
alignas(16) unsigned char beamc_smem[64 * 1024];
#include "beam_c_kernel.inc"

struct LaneArg { BeamCArgs a; int lane; };
static void *lane_main(void *p) {
    LaneArg *la = (LaneArg *)p;
    threadIdx.x = (unsigned)la->lane;
    beam_c_kernel(la->a);
    return nullptr;
}

extern "C" int emul_beam_c(const float *vec, const uint32_t *adj, const uint8_t *codes, const uint8_t *deleted, const float *Q,
                           const float *lut, long long N, int D, int R, int M, long long B, int k, int bw, int dist, int sqrt_out,
                           uint32_t start, int32_t *ids, float *dd, int32_t *hops, int32_t *vis) {
    BeamCArgs a;
    memset(&a, 0, sizeof(a));
    a.vec = vec; a.adj = adj; a.codes = codes; a.deg = nullptr; a.deleted = deleted; a.Q = Q; a.lut = dist == DR_DIST_PQ ? lut : nullptr;
    a.N = N; a.D = D; a.R = R; a.M = M; a.B = B; a.k = k; a.bw = bw; a.dist = dist; a.sqrt_out = sqrt_out; a.start = start;
    a.out_ids = ids; a.out_dist = dd; a.out_hops = hops; a.out_visited = vis;
    a.words = (N + 31) / 32;
    a.bitmap = (uint32_t *)malloc((size_t)a.words * 4);
    memset(a.bitmap, 0xAB, (size_t)a.words * 4);          // the kernel must clear it itself
    const size_t smem = ((size_t)(dist == DR_DIST_PQ ? 0 : D) * 4 + 15) / 16 * 16 + (size_t)(bw + R + 1 + k + 1 + 32) * 8 + 16;
    if (smem > sizeof(beamc_smem)) return 2;
    memset(beamc_smem, 0xCD, sizeof(beamc_smem));
    pthread_barrier_init(&emul_bar, nullptr, 32);
    pthread_t th[32];
    LaneArg la[32];
    for (int i = 0; i < 32; ++i) { la[i].a = a; la[i].lane = i; pthread_create(&th[i], nullptr, lane_main, &la[i]); }
    for (int i = 0; i < 32; ++i) pthread_join(th[i], nullptr);
    pthread_barrier_destroy(&emul_bar);
    free(a.bitmap);
    // the kernel must stay inside the dynamic shared memory launch_beam_c asks for (same expression as in beam_c.cu)
    for (size_t i = smem; i < sizeof(beamc_smem); ++i)
        if (beamc_smem[i] != 0xCD) return 3;
    return 0;
}
