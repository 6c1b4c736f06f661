#!/bin/bash
# round-2 GPU call 4: K2 with multicast clusters + concatenated slices + L2 prefetch; coalesced fused vb slicer with
# verify / repair; K4b with concatenated slice pairs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xcpath.py tests/test_gpu_baseline_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/c4_tests.log 2>&1
echo "tests rc=$?"
tail -5 gpurun_out/c4_tests.log
timeout 600 python tools/prof_sb.py c60 3 both 0:131072,0:262144,0:393216,1:262144,2:262144,3:262144,2:393216,3:393216 > gpurun_out/c4_prof.log 2>&1
echo "prof rc=$?"
cat gpurun_out/c4_prof.log | grep -v "iter 0"
for v in "cl2:" "cl4:B200QC_I8_MODE=393216" "cl1:B200QC_I8_MODE=131072"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench_$name.json 2> gpurun_out/c4_bench_$name.err
  echo "bench $name rc=$?"
done
