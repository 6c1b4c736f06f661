"""One-line summary of bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f  e2e %.3f " % (d["ms_per_step"], d.get("e2e", {}).get("value", float("nan"))),
              {k: round(v["ms_per_launch"], 3) for k, v in d.get("kernels", {}).items()})
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
