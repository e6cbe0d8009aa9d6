set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__grid_size"
run() { timeout 900 ncu --metrics $M --clock-control none -k regex:$2 -s $3 -c $4 --csv --log-file gpurun_out/r02_misc_$1.csv python tests/tools/profile_misc.py $5 > gpurun_out/r02_misc_$1.log 2>&1; tail -1 gpurun_out/r02_misc_$1.log; }
run prune 'prune_kernel' 150 4 build
run buildsearch 'search_kernel' 80 2 build
