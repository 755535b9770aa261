"""JPEG decode front-end — what the reference does with ``cv2.imread(path)`` (scripts/eval/evaluate.py:147) and
``cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)`` (app/computeroot/cv_endpoint.py:151-153) before it calls
``ChessVision.process_image``, moved in front of the batched path: Huffman decoding on host threads,
inverse DCT / chroma upsampling / colour conversion on the GPU, output bit-identical to OpenCV's (``cvb_decode_jpeg``).
Supported subset: baseline 4:2:0 JPEGs whose dimensions are multiples of 16 (every image under the reference's
``data/test``); other files raise ``NativeError`` — there is no CPU fallback.
"""
from __future__ import annotations

from pathlib import Path

import torch


def _engine():
    from .core import _engine_for_statics
    return _engine_for_statics()


def imdecode_batch(streams, engine=None) -> torch.Tensor:
    """list of JPEG byte strings (same dimensions) -> u8[N,H,W,3] BGR CUDA tensor."""
    return (engine or _engine()).decode_jpeg([bytes(s) for s in streams])


def imread_batch(paths, engine=None) -> torch.Tensor:
    return imdecode_batch([Path(p).read_bytes() for p in paths], engine)


def imdecode(stream, engine=None):
    """``cv2.imdecode(buf, cv2.IMREAD_COLOR)``: u8[H,W,3] BGR numpy array."""
    return imdecode_batch([stream], engine)[0].cpu().numpy()


def imread(path, engine=None):
    """``cv2.imread(path)``."""
    return imdecode(Path(path).read_bytes(), engine)
