#!/bin/bash
# round 2, second half: full GPU suite, K1 timing, ncu full-set captures of the staged K1 and of J/K register kernels
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2b_tests.log | cut -c1-300
timeout 200 python tools/time_k1.py c60 2>&1 | tail -1 | tee gpurun_out/r2b_k1.txt
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k regex:ao_eval_sb2 -c 1 -f -o gpurun_out/r2b_k1 python tools/time_k1.py c60 > gpurun_out/r2b_ncu_k1.log 2>&1
timeout 600 $NCU -k 'regex:jk_reg_kernel<1, 1, 1, 1' -c 2 -f -o gpurun_out/r2b_jk_pppp python tools/bench_jk.py taxol_like:3-21g:noshared > gpurun_out/r2b_ncu_pppp.log 2>&1
timeout 600 $NCU -k 'regex:jk_reg_kernel<0, 0, 0, 0' -c 2 -f -o gpurun_out/r2b_jk_ssss python tools/bench_jk.py taxol_like:3-21g:noshared > gpurun_out/r2b_ncu_ssss.log 2>&1
timeout 600 $NCU -k 'regex:jk_reg_kernel<2, 1, 2, 1' -c 2 -f -o gpurun_out/r2b_jk_dpdp python tools/bench_jk.py carbon_cluster20:def2-svp:noshared > gpurun_out/r2b_ncu_dpdp.log 2>&1
ls -la gpurun_out/*.ncu-rep
