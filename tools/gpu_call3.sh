#!/bin/bash
# round-2 GPU call 3: group-bounded fused vb slicer: parity, full suite, bench (fused / unfused), reference arm at full size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c3_tests.log 2>&1
echo "tests rc=$?"
tail -5 gpurun_out/c3_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c3_ref.json 2> gpurun_out/c3_ref.err
echo "ref rc=$?"
tail -c 1500 gpurun_out/c3_ref.json
nproc; free -g | head -2
