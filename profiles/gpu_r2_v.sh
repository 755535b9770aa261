#!/usr/bin/env bash
# k_warp_board with float32 segment offsets (anchor per 16 pixels, magic-number FMA, deferred literal pixels):
# byte parity (geometry + pipeline suites incl. the 256-quad noise fuzz), short bench, launch time and full-set capture of the kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_geometry.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/pytest_v.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_v.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --steps 8 --warmup 3 --api-steps 1 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_v.json").read().strip().splitlines()[-1])
    print(round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2),
          {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, d["roofline_warp_crop"]["frac"], d["clocks"])
except Exception as e:
    print("failed", e)
PY
P="python profiles/prof_step.py --boards 148 --warmup 1 --steps 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_warp_board --csv --log-file gpurun_out/launches_warp_v.csv $P > gpurun_out/prof_launches_v.log 2>&1
grep k_warp_board gpurun_out/launches_warp_v.csv | tail -2 | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_warp_board -s 1 -c 1 -f -o /tmp/prof_warp $P > gpurun_out/prof_full_v.log 2>&1
ncu -i /tmp/prof_warp.ncu-rep --page raw --csv > gpurun_out/prof_warp_raw_v.csv 2> gpurun_out/prof_export_v.err
python profiles/summarize_raw.py gpurun_out/prof_warp_raw_v.csv > gpurun_out/ncu_warp_summary_v.md; cat gpurun_out/ncu_warp_summary_v.md | cut -c1-250
cp /tmp/prof_warp.ncu-rep gpurun_out/prof_warp_v.ncu-rep
