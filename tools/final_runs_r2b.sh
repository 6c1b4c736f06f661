#!/bin/bash
# round 2 (second half) final 1-GPU lines: python bench.py per workload, reference arm, smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/fin_c60.json 2> gpurun_out/fin_c60.err; tail -2 gpurun_out/fin_c60.err
for w in taxol-like-b3lyp-4c taxol-like-pbe0-4c; do timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/fin_$w.json 2> gpurun_out/fin_$w.err; tail -2 gpurun_out/fin_$w.err; done
for w in c60-pbe0-df taxol-like-b3lyp-df taxol-like-pbe-df benzene-lda-4c benzene-scan-4c; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fin_$w.json 2> gpurun_out/fin_$w.err; tail -2 gpurun_out/fin_$w.err; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_ref.json 2>gpurun_out/fin_ref.err
python - <<PY
import json
for w in ("c60","taxol-like-b3lyp-4c","taxol-like-pbe0-4c","c60-pbe0-df","taxol-like-b3lyp-df","taxol-like-pbe-df","benzene-lda-4c","benzene-scan-4c","ref"):
    try:
        d=json.loads(open("gpurun_out/fin_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"],3), round(d["e2e"]["value"],3), d.get("gpu_launches"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
        if "kernels" in d: print("   ", {k:round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
        if d.get("roofline"): print("   roofline", d["roofline"]["kernel"], round(d["roofline"]["achieved"],1), d["roofline"]["unit"], round(d["roofline"]["frac"] or 0,3))
    except Exception as e: print(w, "ERR", e)
PY
