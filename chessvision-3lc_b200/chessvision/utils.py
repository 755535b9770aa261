"""Host helpers of the image->FEN path (reference: chessvision/utils.py:20-132).  Device work goes through
``_native.Engine``; nothing here falls back to PyTorch-eager or OpenCV compute."""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch
from numpy.typing import NDArray

logger = logging.getLogger(__name__)


def get_device() -> torch.device:
    """utils.py:20-29 — this build runs on CUDA (sm_100) only and says so instead of silently picking the CPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("chessvision (B200 build): no CUDA device available and there is no CPU fallback")
    return torch.device("cuda")


def load_state_dict(checkpoint_path: str) -> tuple[dict, dict]:
    """The checkpoint layouts utils.load_model_checkpoint accepts (utils.py:55-86) -> (fp32 state_dict, metadata)."""
    assert checkpoint_path is not None and Path(checkpoint_path).exists(), f"Checkpoint not found: {checkpoint_path}"
    blob = torch.load(checkpoint_path, map_location="cpu")
    metadata: dict = {}
    if isinstance(blob, dict):
        for key in ("model_state_dict", "state_dict", "model"):
            if key in blob:
                metadata = blob.get("metadata", {}) or {}
                blob = blob[key]
                break
    if hasattr(blob, "state_dict"):
        blob = blob.state_dict()
    sd = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in blob.items()}
    return sd, metadata


def ratio(a: float, b: float) -> float:
    """utils.py:89-93."""
    if a == 0 or b == 0:
        return -1
    return min(a, b) / float(max(a, b))


def create_binary_mask(mask: NDArray[np.float32], threshold: float = 0.5) -> NDArray[np.uint8]:
    """utils.py:101-112 (pure host bookkeeping on an already computed probability map)."""
    assert isinstance(mask, np.ndarray), "Mask must be a numpy array"
    assert mask.dtype == np.float32, "Mask must be float32"
    assert 0 <= threshold <= 1, "Threshold must be between 0 and 1"
    return np.where(mask > threshold, 255, 0).astype(np.uint8)
