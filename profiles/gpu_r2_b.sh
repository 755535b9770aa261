#!/usr/bin/env bash
# Round 2, call B: the whole GPU parity suite (new: 10k mask fuzz, 631 GT masks, large-capacity quad path, small inputs,
# generic warp, batched API) and the new bench lines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_pipeline.json 2> gpurun_out/bench_pipeline.err; echo "bench exit $?"; tail -3 gpurun_out/bench_pipeline.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_pipeline.json").read().strip().splitlines()[-1])
    print(round(d["value"], 1), "boards/s  e2e", round(d["e2e"]["value"], 1), "api", round(d["e2e_api"]["value"], 1), d["e2e_api"]["fen_equal_e2e_arm"],
          "ms/step", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, d["clocks"], d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --workload unet-sweep --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_unet_sweep.json 2> gpurun_out/bench_unet_sweep.err; echo "sweep exit $?"; tail -2 gpurun_out/bench_unet_sweep.err
timeout 600 python bench.py --workload classify --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_classify.json 2> gpurun_out/bench_classify.err; echo "classify exit $?"; tail -2 gpurun_out/bench_classify.err
python - <<'PY'
import json
for f in ("unet_sweep", "classify"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), d["roofline"]["achieved"], d["roofline"]["frac"], d.get("clocks"))
        if "sweep" in d:
            print([(s["batch"], round(s["boards_per_s"]), round(s["tflops"])) for s in d["sweep"]])
    except Exception as e:
        print(f, "parse failed", e)
PY
