"""Extraction quality scores of the reference's enrichment pipeline —
``scripts/process_new_raw/process_pipeline.py:357-378,416-467`` with the same names, computed by ``cvb_quality_scores`` on
the device; ``mask_completeness`` is defined for 256x256 arrays (the shape the reference passes).
No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch


def _engine():
    from .core import _engine_for_statics
    return _engine_for_statics()


def quality_scores_batch(values, quad=None, found=None, engine=None) -> np.ndarray:
    """values f32[N,...] device tensor (the reference passes ``BoardExtractionResult.probabilities``), quad f32[N,4,2]
    -> f64[N,4] = (quadrangle_regularity, mask_completeness (NaN unless the arrays are 256x256), probability_distribution,
    probability_confidence)."""
    eng = engine or _engine()
    return eng.quality_scores(values, quad, found).cpu().numpy()


def _one(values: np.ndarray) -> np.ndarray:
    eng = _engine()
    v = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float32).reshape(1, -1)).to(eng.device)
    return quality_scores_batch(v, engine=eng)[0]


def probability_distribution(mask: np.ndarray) -> float:
    """process_pipeline.py:357-378."""
    return float(_one(mask)[2])


def probability_confidence(probabilities: np.ndarray) -> float:
    """process_pipeline.py:459-467."""
    return float(_one(probabilities)[3])


def quadrangle_regularity(quadrangle: np.ndarray | None) -> float:
    """process_pipeline.py:416-456; ``quadrangle`` is the f32[4,1,2] array of ``BoardExtractionResult``."""
    if quadrangle is None:
        return 0.0
    eng = _engine()
    q = torch.from_numpy(np.ascontiguousarray(quadrangle, dtype=np.float32).reshape(1, 4, 2)).to(eng.device)
    dummy = torch.zeros((1, 4), dtype=torch.float32, device=eng.device)
    return float(quality_scores_batch(dummy, q, engine=eng)[0, 0])


def mask_completeness(mask: np.ndarray) -> float:
    """process_pipeline.py:380-414 for a 256x256 array."""
    if np.asarray(mask).shape != (256, 256):
        raise NotImplementedError("mask_completeness on the B200 path takes the 256x256 array the reference passes")
    return float(_one(mask)[1])
