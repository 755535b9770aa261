#!/usr/bin/env bash
# Last check of the round on the final binary and tests: whole GPU suite + smoke + the driver's bench command.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_final.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_final.log | cut -c1-200
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final6.json 2> gpurun_out/bench_final6.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final6.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "boards/s e2e", round(d["e2e"]["value"], 1), "api", round(d["e2e_api"]["value"], 1), "frac", round(d["roofline"]["frac"], 3),
      "warp frac", round(d["roofline_warp_crop"]["frac"], 3), d["clocks"], (d["cpu_baseline"] or {}).get("value"))
PY
