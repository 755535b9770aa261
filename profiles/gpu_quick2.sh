#!/usr/bin/env bash
# smoke() + the JPEG front-end line
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --workload decode --steps 3 --warmup 1 > gpurun_out/decode_n1.json 2> gpurun_out/decode_n1.err; echo "decode exit $?"; cut -c1-700 gpurun_out/decode_n1.json
