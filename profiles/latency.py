#!/usr/bin/env python
"""Single-image latency of the drop-in call (ChessVision.process_image on one 512x512 board, host numpy in -> result objects
out), with the CUDA graph of the pass (default) and with direct launches (CVB_NO_GRAPH=1), plus the device-only time of the
pass.  Writes one JSON line."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "chessvision-3lc_b200")]


def measure():
    import torch
    import bench
    from chessvision import ChessVision
    w = ROOT / "weights"
    cv = ChessVision(board_extractor_weights=str(w / "best_extractor.pth"), classifier_weights=str(w / "best_classifier.pth"), max_batch=1)
    imgs = bench.synthetic_boards(16)
    for i in range(20):
        cv.process_image(imgs[i % 16])
    ts = []
    for i in range(300):
        t0 = time.perf_counter()
        cv.process_image(imgs[i % 16])
        ts.append((time.perf_counter() - t0) * 1e3)
    eng = cv._engine
    dev = torch.from_numpy(imgs[:1]).cuda()
    out = eng.alloc_outputs(1)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(10):
            eng.image_to_fen(dev, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(100):
            eng.image_to_fen(dev, out)
        e1.record(st)
    torch.cuda.synchronize()
    return {"graph": os.environ.get("CVB_NO_GRAPH") is None, "process_image_ms_median": float(np.median(ts)), "process_image_ms_p90": float(np.percentile(ts, 90)),
            "device_pass_ms": e0.elapsed_time(e1) / 100, "graph_replays": eng.graph_replays(), "launches_per_pass": eng.launch_count() // 430}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        print(json.dumps(measure()))
    else:
        res = []
        for env in ({}, {"CVB_NO_GRAPH": "1"}):
            p = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=dict(os.environ, **env))
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            res.append(json.loads(line[-1]) if line else {"error": p.stderr[-400:]})
        print(json.dumps({"single_image_latency": res}))
