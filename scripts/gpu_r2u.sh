#!/bin/bash
# Round 2, session U: multi-GPU measurement of the final build. usage: bash scripts/gpu_r2u.sh N [configs]
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2u_gpus_n$N.txt
for mode in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --scaling $mode \
    > gpurun_out/r2u_bench_n${N}_$mode.json 2> gpurun_out/r2u_bench_n${N}_$mode.err; cut -c1-260 gpurun_out/r2u_bench_n${N}_$mode.json; tail -2 gpurun_out/r2u_bench_n${N}_$mode.err
done
if [ -n "$2" ]; then
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_configs.py --configs $2 $3 $4 $5 \
    > gpurun_out/r2u_configs_n$N.jsonl 2> gpurun_out/r2u_configs_n$N.err; cut -c1-420 gpurun_out/r2u_configs_n$N.jsonl; tail -3 gpurun_out/r2u_configs_n$N.err
fi
