"""Evaluation metrics over image->FEN results — the reference's ``scripts/eval/evaluate.py:28-140,406-440`` with the same
names and return types, computed by ``cvb_eval_metrics`` on the device buffers the pipeline wrote.

python-chess is not needed: the only thing the reference uses it for here is reading the piece-placement field of a FEN
(``chess.BaseBoard(fen)``), which ``fen_to_labels`` does.  ``board_to_labels`` accepts a FEN string or any object with
``board_fen()`` (a ``chess.BaseBoard`` / ``chess.Board``).  There is no CPU fallback for the metric arithmetic.
"""
from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass

import numpy as np
import torch
from numpy.typing import NDArray

from . import constants
from .cv_types import PositionResult

_LABEL_INDEX = {s: i for i, s in enumerate(constants.LABEL_NAMES)}


@dataclass
class PositionAccuracy:
    """evaluate.py:28-34."""

    accuracy: float
    num_correct: int
    total_squares: int = 64


@dataclass
class TopKAccuracyResult:
    """evaluate.py:89-107."""

    k: int
    accuracies: Sequence[float]

    @property
    def top_1(self) -> float:
        return self.accuracies[0]

    @property
    def top_2(self) -> float:
        return self.accuracies[1] if len(self.accuracies) > 1 else 0.0

    @property
    def top_3(self) -> float:
        return self.accuracies[2] if len(self.accuracies) > 2 else 0.0


def fen_to_labels(fen: str) -> list[str]:
    """64 piece symbols in FEN order (a8-h8, a7-h7, ..., a1-h1), "f" = empty: ``board_to_labels(chess.BaseBoard(fen))``."""
    rows = fen.split()[0].split("/")
    if len(rows) != 8:
        raise ValueError(f"expected 8 rows in position part of fen: {fen!r}")   # python-chess raises ValueError too
    labels: list[str] = []
    for row in rows:
        n = 0
        for ch in row:
            if ch.isdigit():
                labels.extend("f" * int(ch))
                n += int(ch)
            elif ch in _LABEL_INDEX and ch != "f":
                labels.append(ch)
                n += 1
            else:
                raise ValueError(f"invalid character in position part of fen: {fen!r}")
        if n != 8:
            raise ValueError(f"expected 8 columns per row in position part of fen: {fen!r}")
    return labels


def board_to_labels(board) -> list[str]:
    """evaluate.py:61-86."""
    return fen_to_labels(board if isinstance(board, str) else board.board_fen())


def _true_indices(fens: Sequence[str]) -> torch.Tensor:
    return torch.tensor([[_LABEL_INDEX[s] for s in fen_to_labels(f)] for f in fens], dtype=torch.uint8).reshape(len(fens), 64)


def _engine():
    from .core import _engine_for_statics
    return _engine_for_statics()


def evaluate_batch(probs, labels, labels_valid, true_fens: Sequence[str], k: int = 3, flip: bool = False, engine=None):
    """Batched form on device tensors (what ``Engine.image_to_fen`` wrote): probs f32[N,64,13], labels / labels_valid
    u8[N,64] (or None) -> (topk_hits i32[N,k], correct i32[N,2]) as CPU tensors."""
    eng = engine or _engine()
    truth = _true_indices(true_fens).to(eng.device)
    hits, correct = eng.eval_metrics(probs, labels, labels_valid, truth, k=k, flip=flip)
    return hits.cpu(), correct.cpu()


def compute_position_accuracy(predicted_fen: str, true_fen: str) -> PositionAccuracy:
    """evaluate.py:37-52."""
    eng = _engine()
    pred = _true_indices([predicted_fen]).to(eng.device)
    dummy = torch.zeros((1, 64, 13), dtype=torch.float32, device=eng.device)
    _, correct = evaluate_batch(dummy, pred, None, [true_fen], k=1, engine=eng)
    c = int(correct[0, 0])
    return PositionAccuracy(accuracy=c / 64, num_correct=c)


def evaluate_position(result: PositionResult, true_fen: str) -> tuple[PositionAccuracy, PositionAccuracy]:
    """evaluate.py:54-58."""
    return compute_position_accuracy(result.original_fen, true_fen), compute_position_accuracy(result.fen, true_fen)


def compute_model_topk_accuracy(model_probabilities: NDArray[np.float32], true_fen: str, k: int = 3) -> TopKAccuracyResult:
    """evaluate.py:109-140."""
    eng = _engine()
    probs = torch.from_numpy(np.ascontiguousarray(model_probabilities, dtype=np.float32).reshape(1, 64, 13)).to(eng.device)
    hits, _ = evaluate_batch(probs, None, None, [true_fen], k=k, engine=eng)
    return TopKAccuracyResult(k=k, accuracies=[int(h) / 64 for h in hits[0]])


def get_label_indices(probabilities: NDArray[np.float32], true_fen: str) -> tuple[list[int], list[int]]:
    """evaluate.py:406-427."""
    return np.argmax(probabilities, axis=1).tolist(), get_validated_indices(true_fen)


def get_validated_indices(fen: str) -> list[int]:
    """evaluate.py:430-440."""
    return [_LABEL_INDEX[s] for s in fen_to_labels(fen)]
