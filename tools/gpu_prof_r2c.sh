#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_c60.json 2> gpurun_out/r2c_c60.err; python tools/show_bench.py gpurun_out/r2c_c60.json
B200QC_DFJ_SIDE_STREAM_SINGLE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_c60_side.json 2> gpurun_out/r2c_c60_side.err; python tools/show_bench.py gpurun_out/r2c_c60_side.json
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k 'regex:jk_reg_kernel<\(int\)1, \(int\)1, \(int\)1, \(int\)1' -c 2 -f -o gpurun_out/r2b_jk_pppp python tools/bench_jk.py taxol_like:3-21g:noshared > gpurun_out/r2b_ncu_pppp.log 2>&1
timeout 600 $NCU -k 'regex:jk_reg_kernel<\(int\)0, \(int\)0, \(int\)0, \(int\)0' -c 2 -f -o gpurun_out/r2b_jk_ssss python tools/bench_jk.py taxol_like:3-21g:noshared > gpurun_out/r2b_ncu_ssss.log 2>&1
timeout 600 $NCU -k 'regex:jk_reg_kernel<\(int\)2, \(int\)1, \(int\)2, \(int\)1' -c 2 -f -o gpurun_out/r2b_jk_dpdp python tools/bench_jk.py carbon_cluster20:def2-svp:noshared > gpurun_out/r2b_ncu_dpdp.log 2>&1
ls -la gpurun_out/*.ncu-rep
