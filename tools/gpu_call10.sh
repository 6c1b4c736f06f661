#!/bin/bash
# full GPU suite on a 2-GPU box, bench at 1 and 2 GPUs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c10_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/c10_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench1.json 2> gpurun_out/c10_bench1.err
echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench2.json 2> gpurun_out/c10_bench2.err
echo "bench2 rc=$?"
B200QC_DFJ_SIDE_STREAM=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench2_noside.json 2> gpurun_out/c10_bench2_noside.err
echo "bench2 noside rc=$?"
python tools/show_bench.py gpurun_out/c10_bench1.json gpurun_out/c10_bench2.json gpurun_out/c10_bench2_noside.json
