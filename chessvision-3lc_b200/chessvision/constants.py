"""Sizes, label tables and default paths of the image->FEN path.

Same names and values as the reference's ``chessvision/constants.py`` (sizes :18-20, NUM_CLASSES :15, LABEL_NAMES :23,
weight paths :46-50, INVALID_PAWN_SQUARES :88-105, SQUARE_NAMES_* :109-129), generated instead of spelled out.
"""
import os
from pathlib import Path

_FILES, _RANKS = "abcdefgh", "12345678"

CVROOT = os.getenv("CVROOT", Path(__file__).resolve().parent.parent.parent.as_posix())
DATA_ROOT = Path(CVROOT) / "data"
WEIGHTS_DIR = Path(CVROOT) / "weights"
BEST_EXTRACTOR_WEIGHTS = str(WEIGHTS_DIR / "best_extractor.pth")
BEST_CLASSIFIER_WEIGHTS = str(WEIGHTS_DIR / "best_classifier.pth")

INPUT_SIZE = (256, 256)   # UNet input
BOARD_SIZE = (512, 512)   # warped board
PIECE_SIZE = (64, 64)     # one square

LABEL_NAMES = list("BKNPQRbknpqr") + ["f"]   # class index -> FEN symbol, "f" = empty square
NUM_CLASSES = len(LABEL_NAMES)
LABEL_INDICES = {name: i for i, name in enumerate(LABEL_NAMES)}
LABEL_DESCRIPTIONS = [f"{colour} {piece}" for colour in ("White", "Black")
                      for piece in ("Bishop", "King", "Knight", "Pawn", "Queen", "Rook")] + ["Empty Square", "Unknown"]
SEGMENTATION_MAP = {0: "background", 255: "chessboard"}

# a1 is dark: file index + rank index even
DARK_SQUARES = {f + r for fi, f in enumerate(_FILES) for ri, r in enumerate(_RANKS) if (fi + ri) % 2 == 0}
INVALID_PAWN_SQUARES = {f + r for f in _FILES for r in "18"}

# square i of the classifier batch, rank 8 first (white at the bottom) / rotated by 180 degrees
SQUARE_NAMES_NORMAL = [f + r for r in reversed(_RANKS) for f in _FILES]
SQUARE_NAMES_FLIPPED = SQUARE_NAMES_NORMAL[::-1]
