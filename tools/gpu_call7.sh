#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xcpath.py tests/test_gpu_baseline_parity.py -m gpu -x -q > gpurun_out/c7_tests.log 2>&1
echo "tests rc=$?"
tail -3 gpurun_out/c7_tests.log
timeout 600 python tools/prof_sb.py c60 3 rho 0:327680,0:311296,0:303104,0:294912,0:286720,4:327680,4:294912,1:294912,2:294912,3:294912,0:163840,0:425984 > gpurun_out/c7_prof.log 2>&1
echo "prof rc=$?"
cat gpurun_out/c7_prof.log | grep -v "iter 0"
