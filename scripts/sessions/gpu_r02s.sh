set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_lut_tc_gpu.py tests/test_build_index_gpu.py -q -m gpu -x 2>&1 | tail -4
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread"
timeout 600 ncu --metrics $M --clock-control none -k regex:kmeans_assign_tc_kernel -s 20 -c 1 --csv --log-file gpurun_out/r02s_kmeans.csv python tests/tools/profile_misc.py kmeans 2>&1 | tail -1
grep -E "time_duration|pipe_tensor|issue_active|inst_executed|registers" gpurun_out/r02s_kmeans.csv | awk -F'","' '{print $(NF-2), $(NF)}'
