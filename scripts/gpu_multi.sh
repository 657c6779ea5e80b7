#!/bin/bash
# Multi-GPU session: bench.py under torchrun on N GPUs (+ optionally the sharded C4/C5 configs). usage: bash scripts/gpu_multi.sh N [configs] [extra bench_configs flags]
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
  > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
if [ -n "$2" ]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_configs.py --configs $2 $3 \
    > gpurun_out/configs_n$N.jsonl 2> gpurun_out/configs_n$N.err; cut -c1-400 gpurun_out/configs_n$N.jsonl; tail -3 gpurun_out/configs_n$N.err
fi
