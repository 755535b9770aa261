#!/usr/bin/env bash
# Device-side JPEG entropy decoding: parity tests + the decode workload with either path.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jpeg.py -m gpu -q -x > gpurun_out/pytest_q.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_q.log | cut -c1-300
for M in default; do
  if [ "$M" = "default" ]; then unset CVB_JPEG_DEVICE_MIN; else export CVB_JPEG_DEVICE_MIN=$M; fi
  timeout 300 python bench.py --workload decode > gpurun_out/bench_decode_$M.json 2> gpurun_out/bench_decode_$M.err; echo "decode $M exit $?"
  python - $M <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_decode_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("decode", sys.argv[1], round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), d["clocks"])
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/bench_decode_{sys.argv[1]}.err").read()[-600:])
PY
done
