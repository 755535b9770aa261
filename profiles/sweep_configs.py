#!/usr/bin/env python
"""BASELINE.json configs[1] and configs[2] as measured side lines (the bench line is configs[3]):
  configs[1]  UNet board-extractor forward alone, batch sweep 1..1024 at the reference input size (512x512x3 -> 256x256 logits)
  configs[2]  square extraction + piece classifier: warp/crop gather + ResNet-18 forward over 4096 boards (262,144 squares)
Device-resident inputs, CUDA events on the launching stream, 3 warm-up + 5 timed passes each.  One JSON line per point."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "chessvision-3lc_b200")]
import bench  # noqa: E402
from chessvision import _native, utils  # noqa: E402

UNET_GFLOP, CLS_GFLOP = 96.335, 18.127


def timed(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    eng = _native.Engine(0, max_batch=128)
    eng.load_unet(utils.load_state_dict(str(ROOT / "weights" / "best_extractor.pth"))[0])
    eng.load_resnet18(utils.load_state_dict(str(ROOT / "weights" / "best_classifier.pth"))[0])
    distinct = bench.synthetic_boards(64)
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
        img = torch.from_numpy(np.concatenate([distinct] * ((n + 63) // 64))[:n]).cuda()
        ms = timed(lambda: eng.unet_forward(img))
        print(json.dumps({"config": "configs[1]: UNet forward alone", "batch": n, "ms": round(ms, 4), "boards_per_s": round(n / ms * 1e3, 1),
                          "tflops": round(UNET_GFLOP * n / ms, 1)}), flush=True)
    n = 4096
    img = torch.from_numpy(np.concatenate([distinct] * (n // 64))).cuda()
    out = eng.image_to_fen(img[:1024], eng.alloc_outputs(1024))                      # quads of the stream for the first 1024, cycled
    quad = out["quad"].repeat(4, 1, 1).contiguous()
    found = out["found"].repeat(4).contiguous()

    def squares_and_classify():
        board = eng.warp_squares(img, quad, found)
        return eng.classify(board)

    ms = timed(squares_and_classify, warm=2, reps=3)
    print(json.dumps({"config": "configs[2]: warp/crop + ResNet-18 over 4096 boards (262,144 squares)", "boards": n, "ms": round(ms, 3),
                      "boards_per_s": round(n / ms * 1e3, 1), "squares_per_s": round(64 * n / ms * 1e3, 1), "classifier_tflops": round(CLS_GFLOP * n / ms, 1),
                      "found_rate": float(found.float().mean())}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
