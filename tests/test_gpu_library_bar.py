"""Library bar on the same B200: the two networks run by stock PyTorch eager (cuDNN) in fp16
channels_last, timed with CUDA events.  Informational — it records what the hand-written convolutions have to beat in
``gpurun_out/library_bar.json``; the only assertion is that both library runs complete."""
import json
import os

import pytest
import torch

from conftest import ROOT
from oracle import nets

pytestmark = pytest.mark.gpu


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_cudnn_eager_throughput():
    torch.backends.cudnn.benchmark = True
    res = {}
    unet = nets.BoardUNet().eval().cuda().half().to(memory_format=torch.channels_last)
    cls = nets.PieceResNet18().eval().cuda().half().to(memory_format=torch.channels_last)
    with torch.no_grad():
        for b in (32, 128):
            x = torch.rand(b, 3, 256, 256, device="cuda", dtype=torch.half).contiguous(memory_format=torch.channels_last)
            ms = timed(lambda: unet(x), 5)
            res[f"unet_fp16_b{b}"] = {"ms": ms, "boards_per_s": b / ms * 1e3, "tflops": 96.335e9 * b / ms / 1e9}
        for b in (32, 128):
            x = torch.rand(b * 64, 1, 64, 64, device="cuda", dtype=torch.half).contiguous(memory_format=torch.channels_last)
            ms = timed(lambda: cls(x), 5)
            res[f"resnet18_fp16_b{b}"] = {"ms": ms, "boards_per_s": b / ms * 1e3, "tflops": 18.127e9 * b / ms / 1e9}
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    json.dump(res, open(ROOT / "gpurun_out" / "library_bar.json", "w"), indent=1)
    print("cuDNN eager fp16 channels_last:", json.dumps(res))
    assert all(v["ms"] > 0 for v in res.values())
