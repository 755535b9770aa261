"""Drop-in for the image->FEN path of ``chessvision`` (reference: chessvision/__init__.py:1-3)."""
from .core import ChessVision

__all__ = ["ChessVision"]
