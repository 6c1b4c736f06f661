#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 python -m pytest tests/test_gpu_multirank.py -q -x 2>&1 | tail -2
for n in 8 4 2; do
  timeout 200 $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --workload taxol-like-b3lyp-4c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/fin8_b3lyp4c_${n}gpu.json 2> gpurun_out/fin8_b3lyp4c_${n}gpu.err
  python tools/show_bench.py gpurun_out/fin8_b3lyp4c_${n}gpu.json
done
timeout 200 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/fin8_c60_8gpu.json 2> gpurun_out/fin8_c60_8gpu.err
python tools/show_bench.py gpurun_out/fin8_c60_8gpu.json
