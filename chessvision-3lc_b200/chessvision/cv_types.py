"""Result records returned by :class:`chessvision.ChessVision` — field-compatible with the reference's
``chessvision/cv_types.py:9-62`` so callers can switch packages without code changes."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from numpy.typing import NDArray


@dataclass
class ValidationFix:
    square_name: str       # e.g. "e8"
    original_piece: str    # label before the rule fired
    corrected_piece: str   # label after
    rule_name: str         # "no_pawns_on_ends"


@dataclass
class BoardExtractionResult:
    probabilities: NDArray[np.float32]        # f32[256,256]; NB the reference stores the *logits* here (core.py:287,306)
    binary_mask: NDArray[np.uint8]            # u8[256,256] in {0,255}
    quadrangle: NDArray[np.float32] | None    # f32[4,1,2] in image coordinates, None when no board was found
    board_image: NDArray[np.uint8] | None     # u8[512,512] gray, None when no board was found


@dataclass
class PositionResult:
    fen: str                                   # after validation
    original_fen: str                          # straight argmax
    model_probabilities: NDArray[np.float32]   # f32[64,13]
    squares: NDArray[np.uint8]                 # u8[64,64,64,1]
    square_names: list[str]
    validation_fixes: list[ValidationFix]


@dataclass
class ChessVisionResult:
    board_extraction: BoardExtractionResult
    position: PositionResult | None
    processing_time: float


@dataclass
class ValidationMetrics:
    accuracy_before: float
    accuracy_after: float
    num_fixes: int
    fixes: list[ValidationFix]

    @property
    def accuracy_delta(self) -> float:
        return self.accuracy_after - self.accuracy_before
