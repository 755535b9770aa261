"""Host helpers of the image->FEN path (reference: chessvision/utils.py:20-132), same names and signatures.  Device work
goes through ``_native.Engine``; nothing here falls back to PyTorch-eager or OpenCV compute."""
from __future__ import annotations

import logging
import os
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


def get_classifier_model(model_id: str = "resnet18") -> torch.nn.Module:
    """utils.py:32-39.  The reference asks timm for ``resnet18(num_classes=13, in_chans=1)``; this returns a module with the
    same parameter names and shapes (so timm checkpoints load with ``load_model_checkpoint``) whose forward pass runs on the
    native tcgen05 classifier."""
    from .modules import NativeResNet18
    logger.info(f"Creating classifier model: {model_id}")
    assert model_id == "resnet18", f"classifier architecture '{model_id}' has no native implementation (resnet18 only)"
    return NativeResNet18()


def load_model_checkpoint(model: torch.nn.Module, checkpoint_path: str, device: torch.device | None = None) -> torch.nn.Module:
    """utils.py:42-86: load a checkpoint in any of the reference's layouts (``model_state_dict`` / timm ``state_dict`` /
    legacy ``model`` / bare state dict / pickled module) into ``model`` and attach its ``metadata``."""
    assert isinstance(model, torch.nn.Module), "Model must be a torch.nn.Module"
    assert Path(checkpoint_path).exists(), f"Checkpoint not found: {checkpoint_path}"
    logger.info(f"Loading checkpoint from {checkpoint_path}")
    state_dict, metadata = load_state_dict(checkpoint_path)
    model.load_state_dict(state_dict)
    if metadata:
        model.metadata = metadata
        logger.debug(f"Loaded checkpoint metadata: {metadata}")
    return model


def listdir_nohidden(path: str) -> list[str]:
    """utils.py:96-98."""
    return [f for f in os.listdir(path) if not f.startswith(".")]


def extract_perspective(image: NDArray[np.uint8], approx: NDArray[np.float32], out_size: tuple[int, int]) -> NDArray[np.uint8]:
    """utils.py:115-132: ``cv2.getPerspectiveTransform`` + ``cv2.warpPerspective`` on the device (``cvb_warp_perspective``),
    bit-identical to OpenCV for 1- and 3-channel uint8 images and any ``out_size``."""
    assert isinstance(image, np.ndarray), "Image must be a numpy array"
    assert image.dtype == np.uint8, "Image must be uint8"
    assert isinstance(approx, np.ndarray), "Approx must be a numpy array"
    assert approx.dtype == np.float32, "Approx must be float32"
    assert len(approx) == 4, "Approx must contain exactly 4 points"
    assert image.ndim == 2 or (image.ndim == 3 and image.shape[2] in (1, 3)), "Image must be u8[H,W], u8[H,W,1] or u8[H,W,3]"
    from .core import _engine_for_statics
    eng = _engine_for_statics()
    dev_img = torch.from_numpy(np.ascontiguousarray(image)).to(eng.device)
    out = eng.warp_perspective(dev_img, torch.from_numpy(np.ascontiguousarray(approx, dtype=np.float32).reshape(4, 2)), out_size)
    res = out.cpu().numpy()
    return res[:, :, 0] if image.ndim == 3 and image.shape[2] == 1 else res   # cv2 drops a trailing channel axis of one


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
