"""Drop-in for the image->FEN path of ``chessvision`` (reference: chessvision/__init__.py:1-3)."""

__all__ = ["ChessVision"]


def __getattr__(name):
    if name == "ChessVision":
        from .core import ChessVision
        return ChessVision
    raise AttributeError(name)
