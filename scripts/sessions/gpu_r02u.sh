timeout 900 ncu --set full --import-source on --clock-control none -k regex:lut_u8_tc_kernel -c 2 -f -o gpurun_out/r02u_lut python tests/tools/profile_misc.py lut > gpurun_out/r02u.log 2>&1
ls -la gpurun_out/r02u_lut.ncu-rep
