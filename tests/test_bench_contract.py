"""bench.py's contract on the arm that runs without a GPU: ``--impl reference`` prints exactly ONE JSON line on stdout with
the keys the driver reads (everything else a library prints goes to stderr).  CPU only; the B200 arm is exercised on the
GPU box by the driver itself."""
import json
import subprocess
import sys

from conftest import ROOT

REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "cpu_baseline", "e2e")


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                        "--cpu-boards-per-step", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "boards/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and "stage_ms_per_board" in d["cpu_baseline"] and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_roofline_traffic_cites_the_committed_capture_and_scales_with_the_launch_size():
    """`roofline.traffic` is read from profiles/<round>/traffic.json (written from the ncu --set full csv of one 148-board
    pass) and scaled to the boards one launch of the bench covers; both kernel classes the pipeline line cites must be there."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for cls in ("unet_conv_tc", "warp_board"):
        per148, src = bench.ncu_traffic(cls)
        assert per148 and per148 > 0 and "traffic.json" in src and "148 boards per launch" in src, (cls, src)
        scaled, src2 = bench.ncu_traffic(cls, 296)
        assert abs(scaled - 2 * per148) < 1e-6 * per148 and "scaled to 296 boards per launch" in src2
    # algorithmic bytes of the warp stage: 512x512x3 read + 512x512 written per board
    assert bench.WARP_BYTES_PER_BOARD == 512 * 512 * 3 + 512 * 512
