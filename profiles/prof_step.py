#!/usr/bin/env python
"""Small driver for ncu: W warm-up + K timed passes of the device-resident image->FEN pipeline over one chunk of boards
(the same synthetic boards and weights as bench.py).  Used under
    ncu --metrics gpu__time_duration.sum --clock-control none ... python profiles/prof_step.py
    ncu --set full --clock-control none --import-source on -k regex:<kernel> ... python profiles/prof_step.py
Numbers printed under a profiler are never bench values."""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "chessvision-3lc_b200")]
import bench  # noqa: E402  (synthetic_boards)
from chessvision import _native, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--boards", type=int, default=128)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
eng = _native.Engine(0, max_batch=a.boards)
eng.load_unet(utils.load_state_dict(str(ROOT / "weights" / "best_extractor.pth"))[0])
eng.load_resnet18(utils.load_state_dict(str(ROOT / "weights" / "best_classifier.pth"))[0])
distinct = bench.synthetic_boards(min(a.boards, 64))
img = torch.from_numpy(np.concatenate([distinct] * ((a.boards + 63) // 64))[: a.boards]).cuda()
out = eng.alloc_outputs(a.boards)
for _ in range(a.warmup + a.steps):
    eng.image_to_fen(img, out)
torch.cuda.synchronize()
print("found rate", float(out["found"].float().mean()), "launches", eng.launch_count())
eng.close()
