#!/bin/bash
# round-2 GPU call 1: new K2 (point-stationary) + fused vb slicer: parity, then bench variants
mkdir -p gpurun_out
echo "== xcpath + baseline parity" > gpurun_out/c1_tests.log
timeout 900 python -m pytest tests/test_gpu_xcpath.py tests/test_gpu_baseline_parity.py -m gpu -x -q >> gpurun_out/c1_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c1_tests.log
echo "== full suite" >> gpurun_out/c1_tests.log
timeout 1200 python -m pytest tests -m gpu -q >> gpurun_out/c1_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c1_tests.log
for v in "new:" "oldrho:B200QC_RHO_I8_BN=64" "unfused:B200QC_VXC_FUSED_VB=0" "nbc12:B200QC_I8_MODE=49152"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_$name.json 2> gpurun_out/c1_bench_$name.err
  echo "bench $name rc=$?" >> gpurun_out/c1_tests.log
done
tail -5 gpurun_out/c1_tests.log
nvidia-smi --query-gpu=name,memory.total --format=csv; free -g | head -2; nproc
