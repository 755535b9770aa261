#!/usr/bin/env bash
# Round 2, call A: parity suite with programmatic dependent launch on, bench A/B (PDL off/on, chunk 128 vs 148 = the SM
# count, so every conv launch is a whole number of waves), and a source-level ncu capture of the two long non-conv kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/nvsmi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
run() { # name, env, args
  env $2 timeout 300 python bench.py --no-cpu-baseline --steps 6 --warmup 3 $3 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench $1 exit $?"
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2),
          {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, d["clocks"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run pdl0_c128 CVB_PDL=0 "--chunk 128 --boards 1024"
run pdl1_c128 CVB_PDL=1 "--chunk 128 --boards 1024"
run pdl1_c148 CVB_PDL=1 "--chunk 148 --boards 1184"
run pdl0_c148 CVB_PDL=0 "--chunk 148 --boards 1184"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_mask_to_quad_fast|k_resnet_stem_tc|k_unet_stem_tc' -s 3 -c 3 -f -o gpurun_out/aux_kernels \
    python profiles/prof_step.py --boards 128 --warmup 1 --steps 1 > gpurun_out/prof_aux.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
