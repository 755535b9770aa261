#!/usr/bin/env bash
# Round-2 evidence on one B200, final code: the driver's bench command (both arms), the side workloads, ncu launch list +
# full-set capture + DRAM traffic of one 148-board pass, the training step's launch list.
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference_n1.err; echo "reference arm exit $?"
timeout 300 python bench.py --workload unet-sweep > gpurun_out/bench_unet_sweep.json 2> gpurun_out/bench_unet_sweep.err; echo "unet-sweep exit $?"
timeout 300 python bench.py --workload classify > gpurun_out/bench_classify.json 2> gpurun_out/bench_classify.err; echo "classify exit $?"
timeout 300 python bench.py --workload decode > gpurun_out/bench_decode.json 2> gpurun_out/bench_decode.err; echo "decode exit $?"
for B in 8 32; do
  timeout 600 python bench.py --workload train --train-batch $B --steps $((B == 8 ? 300 : 120)) --warmup 5 > gpurun_out/train_n1_b$B.json 2> gpurun_out/train_n1_b$B.err; echo "train b=$B exit $?"
done
timeout 300 python profiles/latency.py > gpurun_out/latency_single_image.json 2> gpurun_out/latency.err; echo "latency exit $?"
python - <<'PY'
import json
for f in ("bench_n1", "bench_reference_n1", "bench_unet_sweep", "bench_classify", "bench_decode", "train_n1_b8", "train_n1_b32"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), d["unit"], "ms/step", round(d.get("ms_per_step", 0), 2), d.get("clocks"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "failed", e)
PY
bash profiles/gpu_profile_r2b.sh
