"""ORACLE tooling (test / measurement infrastructure, never product code): import the UNMODIFIED reference package.

``load_reference()`` imports ``chessvision`` of the reference from ``/root/reference`` where that checkout exists (the
build container) and otherwise from ``oracle/_ref`` (byte-for-byte copies staged by ``oracle/build_ref.sh``; that
directory is git-ignored and travels to the GPU box with the snapshot).  The reference imports two packages that are not
installed in this image; both are replaced by minimal stand-ins registered in ``sys.modules`` before the import
(SURVEY.md Appendix D):

* ``chess``  (python-chess 1.11.2): only ``SQUARE_NAMES``, ``Piece.from_symbol``, ``BaseBoard.set_piece_at/board_fen``
  are touched (core.py:330-349);
* ``timm``   (1.0.15): only ``create_model("resnet18", num_classes=13, in_chans=1)`` (utils.py:35-39), provided by
  torchvision's resnet18 with a 1-channel conv1 (same topology and state-dict keys).

The product package in ``chessvision-3lc_b200/`` has the same import name, so a process imports one or the other:
``bench.py`` runs its reference legs in a process of their own.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CV_REFERENCE", "/root/reference")
STAGED = os.path.join(ROOT, "oracle", "_ref")


def install_standins():
    import torch
    import torchvision

    chess = types.ModuleType("chess")
    chess.SQUARE_NAMES = [f + r for r in "12345678" for f in "abcdefgh"]

    class Piece:
        def __init__(self, sym):
            self.sym = sym

        @classmethod
        def from_symbol(cls, sym):
            return cls(sym)

        def symbol(self):
            return self.sym

    class BaseBoard:
        def __init__(self, board_fen=None):
            self.sq = [None] * 64

        def set_piece_at(self, square, piece, promoted=False):
            self.sq[square] = piece

        def board_fen(self, promoted=False):
            rows = []
            for r in range(7, -1, -1):
                row, e = "", 0
                for f in range(8):
                    p = self.sq[r * 8 + f]
                    if p is None:
                        e += 1
                    else:
                        row += (str(e) if e else "") + p.symbol()
                        e = 0
                rows.append(row + (str(e) if e else ""))
            return "/".join(rows)

    chess.Piece, chess.BaseBoard = Piece, BaseBoard
    sys.modules["chess"] = chess

    timm = types.ModuleType("timm")

    def create_model(model_id, num_classes=1000, in_chans=3, **kw):
        assert model_id == "resnet18"
        m = torchvision.models.resnet18(num_classes=num_classes)
        m.conv1 = torch.nn.Conv2d(in_chans, 64, 7, 2, 3, bias=False)
        return m

    timm.create_model = create_model
    sys.modules["timm"] = timm


def reference_root() -> str | None:
    """Directory that holds the reference's ``chessvision`` package: the checkout if present, else the staged copy."""
    for root in (REF, STAGED):
        if os.path.isfile(os.path.join(root, "chessvision", "core.py")):
            return root
    return None


def load_reference():
    """-> the reference's ``chessvision`` module (unmodified).  Raises ImportError when neither location exists or when
    another ``chessvision`` (the product package) is already imported in this process."""
    root = reference_root()
    if root is None:
        raise ImportError(f"reference not found: neither {REF}/chessvision nor {STAGED}/chessvision exists (run oracle/build_ref.sh)")
    loaded = sys.modules.get("chessvision")
    if loaded is not None and not os.path.abspath(getattr(loaded, "__file__", "")).startswith(os.path.abspath(root)):
        raise ImportError(f"a different 'chessvision' is already imported from {loaded.__file__}")
    install_standins()
    sys.path.insert(0, root)
    import chessvision  # the reference package, unmodified
    assert os.path.abspath(chessvision.__file__).startswith(os.path.abspath(root)), chessvision.__file__
    return chessvision


def reference_pipeline(extractor_weights: str, classifier_weights: str):
    """``ChessVision`` of the reference with both networks on the host cores (its ``utils.get_device`` would pick CUDA on
    a GPU box; this arm is the CPU implementation), loaded eagerly."""
    import torch
    ref = load_reference()
    ref.core.utils.get_device = lambda: torch.device("cpu")
    return ref.ChessVision(board_extractor_weights=extractor_weights, classifier_weights=classifier_weights,
                           classifier_model_id="resnet18", lazy_load=False)
