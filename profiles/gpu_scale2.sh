#!/usr/bin/env bash
# N=2 check of the bench contract, launched the way the driver launches it (one rank per GPU over NCCL); stdout must hold
# exactly one JSON line per run.
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 \
    > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $? stdout lines: $(wc -l < gpurun_out/bench_n2.json)"; cut -c1-200 gpurun_out/bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 \
    > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 exit $? stdout lines: $(wc -l < gpurun_out/bench_ref_n2.json)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --workload train --gpus 2 --steps 4 --warmup 3 \
    > gpurun_out/train_n2.json 2> gpurun_out/train_n2.err; echo "train n2 exit $? stdout lines: $(wc -l < gpurun_out/train_n2.json)"; cut -c1-200 gpurun_out/train_n2.json
