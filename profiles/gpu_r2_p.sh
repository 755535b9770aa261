#!/usr/bin/env bash
# A/B: boards per pipeline chunk (fixed per-launch overheads amortised over more boards).
mkdir -p gpurun_out
run() {
  timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --api-steps 1 --chunk $1 > gpurun_out/bench_chunk$1.json 2> gpurun_out/bench_chunk$1.err; echo "chunk $1 exit $?"
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_chunk{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("chunk", sys.argv[1], round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "api", round(d["e2e_api"]["value"], 1), "ms/step", round(d["ms_per_step"], 2),
          d["config"]["boards_per_gpu_per_step"], {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, d["clocks"])
except Exception as e:
    print(sys.argv[1], "failed", e); print(open(f"gpurun_out/bench_chunk{sys.argv[1]}.err").read()[-800:])
PY
}
run 148
run 296
run 592
run 148
