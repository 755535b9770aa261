"""Checkpoint files either side of the image->FEN path.

Weights reach ``ChessVision`` as files written by the reference's training scripts and leave the B200 trainer as files those
scripts (and ``utils.load_model_checkpoint``, utils.py:42-86) read back.  Pure host bookkeeping on ``torch.save`` dictionaries:

* ``save_unet_checkpoint``        scripts/train/train_unet.py:31-40   {"model_state_dict", "optimizer_state_dict", "metadata"}
* ``save_classifier_checkpoint``  scripts/train/train_classifier.py:112-123 (same three keys)
* ``strip_optimizer``             scripts/train/strip_optimizer.py:15-47: drop the optimizer state of a checkpoint
* ``checkpoint_layout``           which of the five layouts of utils.py:55-86 a file uses
"""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Any

import torch

logger = logging.getLogger(__name__)


def _save(model_or_state, checkpoint_path: str, optimizer_state, metadata: dict[str, Any] | None) -> None:
    state = model_or_state.state_dict() if hasattr(model_or_state, "state_dict") else dict(model_or_state)
    blob = {"model_state_dict": state, "metadata": metadata or {}}
    if optimizer_state is not None:
        blob["optimizer_state_dict"] = optimizer_state.state_dict() if hasattr(optimizer_state, "state_dict") else optimizer_state
    Path(checkpoint_path).parent.mkdir(parents=True, exist_ok=True)
    torch.save(blob, checkpoint_path)


def save_unet_checkpoint(model, checkpoint_path: str, optimizer=None, metadata: dict[str, Any] | None = None) -> None:
    """train_unet.py:31-40.  ``model``: a module, a state dict, or a ``chessvision.training.UNetTrainer`` (its ``state_dict()``
    returns the reference's parameter names)."""
    _save(model, checkpoint_path, optimizer, metadata)


def save_classifier_checkpoint(model, checkpoint_path: str, optimizer=None, metadata: dict[str, Any] | None = None) -> None:
    """train_classifier.py:112-123."""
    _save(model, checkpoint_path, optimizer, metadata)


def checkpoint_layout(checkpoint_path: str) -> str:
    """One of "model_state_dict", "state_dict" (timm), "model" (legacy), "bare" (a plain state dict), "module" (a pickled
    nn.Module) -- the cases utils.load_model_checkpoint distinguishes (utils.py:55-86)."""
    blob = torch.load(checkpoint_path, map_location="cpu")
    if isinstance(blob, dict):
        for key in ("model_state_dict", "state_dict", "model"):
            if key in blob:
                return key
        return "bare"
    return "module"


def strip_optimizer(checkpoint_path: str, output_path: str | None = None) -> None:
    """strip_optimizer.py:15-47: keep ``model_state_dict`` (or timm's ``state_dict``) and ``metadata`` only; other layouts are
    left untouched with a warning, exactly like the reference's script.  Overwrites the input when ``output_path`` is None."""
    checkpoint = torch.load(checkpoint_path, map_location="cpu")
    if not isinstance(checkpoint, dict):
        logger.warning(f"Checkpoint at {checkpoint_path} is not a dictionary, skipping...")
        return
    if "model_state_dict" in checkpoint:
        stripped = {"model_state_dict": checkpoint["model_state_dict"], "metadata": checkpoint.get("metadata", {})}
    elif "state_dict" in checkpoint:
        stripped = {"state_dict": checkpoint["state_dict"], "metadata": checkpoint.get("metadata", {})}
    else:
        logger.warning(f"Checkpoint at {checkpoint_path} has unexpected format, skipping...")
        return
    output_path = output_path or checkpoint_path
    torch.save(stripped, output_path)
    logger.info(f"Saved stripped checkpoint to {output_path}")
