#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c11_tests.log 2>&1
echo "tests rc=$?"; tail -30 gpurun_out/c11_tests.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench1.json 2> gpurun_out/c11_bench1.err
python tools/show_bench.py gpurun_out/c11_bench1.json
