#!/usr/bin/env bash
# Final code of the round (k_warp_board with float32 segment offsets): whole GPU suite, smoke(), the driver's bench command,
# ncu launch list + full-set capture + traffic over one 148-board pass.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_final.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_final.log | cut -c1-200
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final5.json 2> gpurun_out/bench_final5.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final5.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "boards/s e2e", round(d["e2e"]["value"], 1), "api", round(d["e2e_api"]["value"], 1), "frac", round(d["roofline"]["frac"], 3),
      "warp frac", round(d["roofline_warp_crop"]["frac"], 3), d["clocks"], (d["cpu_baseline"] or {}).get("value"), d["stage_ms_per_step"])
PY
P="python profiles/prof_step.py --boards 148 --warmup 1 --steps 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/prof_launches.log 2>&1
N=$(grep -c 'gpu__time_duration.sum' gpurun_out/launches.csv)
PASS=$((N / 2))
echo "launches per pass: $PASS"
timeout 600 ncu --set full --clock-control none --import-source on -s $PASS -c $PASS -f -o /tmp/prof_all $P > gpurun_out/prof_full.log 2>&1
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2> gpurun_out/prof_export.err
python profiles/summarize_raw.py gpurun_out/prof_all_raw.csv > gpurun_out/ncu_full_summary.md; grep -i "warp_board\|mask_to_quad_fast\|k_head\|stem" gpurun_out/ncu_full_summary.md | cut -c1-200
python profiles/traffic_from_ncu.py gpurun_out/prof_all_raw.csv 148 > gpurun_out/traffic.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_warp_board -s 1 -c 1 -f -o gpurun_out/prof_warp_final $P > gpurun_out/prof_warp_final.log 2>&1
