timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/final_c60.json 2> gpurun_out/final_c60.err; tail -2 gpurun_out/final_c60.err
for w in c60-pbe0-df taxol-like-b3lyp-df taxol-like-pbe-df benzene-lda-4c; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_$w.json 2> gpurun_out/final_$w.err; tail -2 gpurun_out/final_$w.err; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2>gpurun_out/final_ref.err
python - <<EOF
import json
for w in ("c60","c60-pbe0-df","taxol-like-b3lyp-df","taxol-like-pbe-df","benzene-lda-4c","ref"):
    try:
        d=json.load(open("gpurun_out/final_%s.json"%w))
        print(w, round(d["value"],3), round(d["e2e"]["value"],3), d.get("gpu_launches"), d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
        if "kernels" in d: print("   ", {k:round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
        if d.get("roofline"): print("   roofline", d["roofline"]["kernel"], round(d["roofline"]["achieved"],1), d["roofline"]["unit"], round(d["roofline"]["frac"] or 0,3))
    except Exception as e: print(w, "ERR", e)
EOF
