#!/usr/bin/env bash
# Round 2 multi-GPU lines on N GPUs of one box (gpurun --gpus N -- bash profiles/gpu_scale_r2.sh N): the image->FEN pipeline
# (batch-sharded, no collective) and the UNet training step (data-parallel, bucketed all-reduce overlapped with backward),
# each launched the way the driver launches bench.py.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "1" ]; then TR="python"; fi
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --api-steps 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "pipeline N=$N exit $?"
for B in 8 32; do
  timeout 600 $TR bench.py --workload train --gpus $N --train-batch $B --steps $((B == 8 ? 300 : 120)) --warmup 5 > gpurun_out/train_n${N}_b$B.json 2> gpurun_out/train_n${N}_b$B.err; echo "train N=$N b=$B exit $?"
done
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in (f"bench_n{n}", f"train_n{n}_b8", f"train_n{n}_b32"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/train_n${N}_b8.err
