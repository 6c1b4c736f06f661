#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -2
timeout 300 python bench.py --workload taxol-like-b3lyp-4c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/last_b3lyp4c.json 2> gpurun_out/last_b3lyp4c.err; python tools/show_bench.py gpurun_out/last_b3lyp4c.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/last_c60.json 2> gpurun_out/last_c60.err; python tools/show_bench.py gpurun_out/last_c60.json
