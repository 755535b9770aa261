#!/usr/bin/env bash
# N=8 (or $1) check of the bench contract, launched the way the driver launches it.
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 4 --warmup 3 \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"; tail -3 gpurun_out/bench_n$N.err; cut -c1-400 gpurun_out/bench_n$N.json
