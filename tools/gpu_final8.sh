#!/bin/bash
# round-2 multi-GPU run on one 8-GPU box: C60 strong scaling, configs[3] (taxol-like B3LYP, DF-J + DF-K) at 8 GPUs,
# configs[4] sweep points at 2 / 4 / 8 GPUs
mkdir -p gpurun_out
O=gpurun_out
P=29600
run() { # name ngpu workload steps
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus $2 --workload $3 --steps $4 --warmup 3 --no-cpu-baseline > $O/f8_$1.json 2> $O/f8_$1.err
  echo "$1 rc=$?"
}
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $O/f8_tests.log 2>&1; echo "multirank tests rc=$?"; tail -2 $O/f8_tests.log
run c60_8 8 c60-pbe-df 20
run c60_4 4 c60-pbe-df 20
run c60_2 2 c60-pbe-df 10
B200QC_DFJ_SIDE_STREAM=0 run c60_8_noside 8 c60-pbe-df 20
run taxol_b3lyp_8 8 taxol-like-b3lyp-df 10
run cluster72_8 8 cluster72-pbe-df 10
run cluster72_4 4 cluster72-pbe-df 10
run cluster72_2 2 cluster72-pbe-df 5
run cluster143_8 8 cluster143-pbe-df 5
python tools/show_bench.py $O/f8_c60_8.json $O/f8_c60_4.json $O/f8_c60_2.json $O/f8_c60_8_noside.json $O/f8_taxol_b3lyp_8.json $O/f8_cluster72_8.json $O/f8_cluster72_4.json $O/f8_cluster72_2.json $O/f8_cluster143_8.json
for f in $O/f8_*.err; do if [ -s $f ]; then echo "== $f"; tail -3 $f; fi; done
