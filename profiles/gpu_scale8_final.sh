#!/usr/bin/env bash
# Final code on 8 GPUs of one box, launched the way the driver launches bench.py: pipeline (both e2e arms) + training step (batch 32).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --api-steps 1 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "pipeline N=8 exit $?"
timeout 600 $TR bench.py --workload train --gpus 8 --train-batch 32 --steps 120 --warmup 5 > gpurun_out/train_n8_b32.json 2> gpurun_out/train_n8_b32.err; echo "train N=8 exit $?"
python - <<'PY'
import json
for f in ("bench_n8", "train_n8_b32"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench_n8.err
