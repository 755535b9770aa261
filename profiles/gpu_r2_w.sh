#!/usr/bin/env bash
# k_warp_board at five CTAs per SM (48 registers) + multiply-add tap addressing: byte parity, the driver's bench command, kernel time.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_w.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_w.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_w.log | cut -c1-200
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_w.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "boards/s e2e", round(d["e2e"]["value"], 1), "api", round(d["e2e_api"]["value"], 1), "frac", round(d["roofline"]["frac"], 3),
      "warp frac", round(d["roofline_warp_crop"]["frac"], 3), d["clocks"], (d["cpu_baseline"] or {}).get("value"), d["stage_ms_per_step"])
print(d["roofline"]["traffic"], d["roofline"]["algorithmic_bytes_per_launch"], d["roofline_warp_crop"]["traffic"])
PY
P="python profiles/prof_step.py --boards 148 --warmup 1 --steps 1"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_warp_board -s 1 -c 1 -f -o gpurun_out/prof_warp_w $P > gpurun_out/prof_warp_w.log 2>&1
ncu -i gpurun_out/prof_warp_w.ncu-rep --page raw --csv > gpurun_out/prof_warp_raw_w.csv 2> /dev/null
python profiles/summarize_raw.py gpurun_out/prof_warp_raw_w.csv | tail -1 | cut -c1-250
