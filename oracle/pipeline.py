"""ORACLE (test infrastructure, not product code): the whole image->FEN path on the CPU in fp32, composed from the
restatements in ``oracle/geometry.py`` and ``oracle/nets.py``.  Mirrors ``ChessVision.process_image``
(chessvision/core.py:152-195) stage by stage and returns every intermediate so parity can be reported per stage.

Importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and --impl reference) — never the product package.
"""
from __future__ import annotations

import numpy as np
import torch

from . import geometry as g
from .nets import BoardUNet, PieceResNet18, load_state


class OraclePipeline:
    def __init__(self, unet: BoardUNet, classifier: PieceResNet18):
        self.unet = unet.eval()
        self.classifier = classifier.eval()

    @classmethod
    def from_checkpoints(cls, extractor_path: str, classifier_path: str) -> "OraclePipeline":
        return cls(load_state(BoardUNet(), extractor_path), load_state(PieceResNet18(), classifier_path))

    @torch.no_grad()
    def logits(self, image: np.ndarray) -> np.ndarray:
        """core.py:212-220: INTER_AREA resize, /255, NHWC->NCHW (BGR kept), UNet forward."""
        small = g.resize_area_half(image) if image.shape[:2] == (512, 512) else g.resize_area(image, (256, 256))
        x = (torch.from_numpy(small[None].astype(np.float32)) / 255).permute(0, 3, 1, 2)
        return self.unet(x)[0, 0].numpy()

    @torch.no_grad()
    def probabilities(self, board: np.ndarray) -> np.ndarray:
        """core.py:232-243: squares, /255, classifier forward, softmax."""
        squares = g.extract_squares(board)
        batch = torch.from_numpy(squares.astype(np.float32)).permute(0, 3, 1, 2) / 255.0
        return torch.softmax(self.classifier(batch), dim=1).numpy()

    def process_image(self, image: np.ndarray, threshold: float = 0.5, flip: bool = False, quad_override=None) -> dict:
        assert isinstance(image, np.ndarray) and image.dtype == np.uint8 and image.ndim == 3
        out = {"logits": self.logits(image)}
        out["mask"] = g.binary_mask(out["logits"], threshold)
        quad = g.find_quadrangle(out["mask"]) if quad_override is None else quad_override
        out["quad"] = quad
        out["found"] = quad is not None
        if quad is None:
            return out
        out["board"] = g.extract_board(image, g.scale_quadrangle(quad, image.shape[:2]))
        out["probs"] = self.probabilities(out["board"])
        fen, original_fen, labels, fixed, fixes = g.position_from_probabilities(out["probs"], flip)
        out.update(fen=fen, original_fen=original_fen, labels=labels, labels_valid=fixed, fixes=fixes)
        return out
