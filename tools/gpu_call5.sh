#!/bin/bash
# round-2 GPU call 5: K2 epilogue on the 16x256b fragment layout, magic-number slicing digits
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xcpath.py tests/test_gpu_baseline_parity.py tests/test_gpu_fullsize.py tests/test_gpu_dfk.py -m gpu -x -q > gpurun_out/c5_tests.log 2>&1
echo "tests rc=$?"
tail -5 gpurun_out/c5_tests.log
timeout 600 python tools/prof_sb.py c60 3 both 0:131072,0:262144,0:393216,1:262144,2:262144,3:262144 > gpurun_out/c5_prof.log 2>&1
echo "prof rc=$?"
cat gpurun_out/c5_prof.log | grep -v "iter 0"
for v in "cl2:" "cl4:B200QC_I8_MODE=393216" "cl1:B200QC_I8_MODE=131072" "unfused:B200QC_VXC_FUSED_VB=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench_$name.json 2> gpurun_out/c5_bench_$name.err
  echo "bench $name rc=$?"
done
python - <<'PY'
import json
for n in ("cl2","cl4","cl1","unfused"):
    try:
        d=json.loads(open("gpurun_out/c5_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
    except Exception as e: print(n, "ERR", e)
PY
