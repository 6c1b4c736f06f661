#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xcpath.py tests/test_gpu_baseline_parity.py -m gpu -x -q > gpurun_out/c8_tests.log 2>&1
echo "tests rc=$?"
tail -3 gpurun_out/c8_tests.log
timeout 600 python tools/prof_sb.py c60 3 rho 0:0,5:0,0:131072,0:393216,1:0,2:0,3:0,0:311296 > gpurun_out/c8_prof.log 2>&1
echo "prof rc=$?"
cat gpurun_out/c8_prof.log | grep -v "iter 0"
for v in "def:" "loose5:B200QC_I8_MODE=5242880" "loose6:B200QC_I8_MODE=6291456" "loose7:B200QC_I8_MODE=7340032"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c8_bench_$name.json 2> gpurun_out/c8_bench_$name.err
  echo "bench $name rc=$?"
done
python - <<'PY'
import json
for n in ("def","loose5","loose6","loose7"):
    try:
        d=json.loads(open("gpurun_out/c8_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],3), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
    except Exception as e: print(n, "ERR", e)
PY
