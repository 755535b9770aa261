"""ORACLE tooling: golden vectors for the evaluation metrics and extraction quality scores (n4), taken
from the UNMODIFIED reference functions ``scripts/eval/evaluate.py`` and ``scripts/process_new_raw/process_pipeline.py``.

Runs only where /root/reference exists.  Both scripts import packages that are not installed here (``tlc``, ``cairosvg``,
``boto3``, ``chess``, ``PIL`` ...) at module level for their CLI / 3LC parts; those are registered as inert stand-ins in
``sys.modules`` — only ``chess.BaseBoard(fen)`` / ``piece_at`` / ``piece_map`` are actually exercised (a 20-line FEN
reader), exactly the calls of evaluate.py:39-47,74-76,118.  While generating, ``oracle/metrics.py`` is compared with the
reference on every vector; a mismatch aborts.

    python oracle/make_golden_metrics.py      # writes tests/golden/metrics_vectors.npz
"""
from __future__ import annotations

import glob
import os
import sys
import types
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CV_REFERENCE", "/root/reference")


def install_standins():
    chess = types.ModuleType("chess")
    chess.SQUARES = list(range(64))
    chess.SQUARE_NAMES = [f + r for r in "12345678" for f in "abcdefgh"]

    class Piece:
        def __init__(self, sym):
            self.sym = sym

        @classmethod
        def from_symbol(cls, sym):
            return cls(sym)

        def symbol(self):
            return self.sym

        def __eq__(self, other):
            return isinstance(other, Piece) and other.sym == self.sym

        def __hash__(self):
            return hash(self.sym)

    class BaseBoard:
        def __init__(self, board_fen="rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR"):
            self.sq = [None] * 64
            if board_fen is not None:
                for r, row in enumerate(board_fen.split()[0].split("/")):
                    f = 0
                    for ch in row:
                        if ch.isdigit():
                            f += int(ch)
                        else:
                            self.sq[(7 - r) * 8 + f] = Piece(ch)   # square index 0 = a1
                            f += 1

        def piece_at(self, square):
            return self.sq[square]

        def piece_map(self):
            return {i: p for i, p in enumerate(self.sq) if p is not None}

        def set_piece_at(self, square, piece, promoted=False):
            self.sq[square] = piece

    chess.Piece, chess.BaseBoard, chess.Board = Piece, BaseBoard, BaseBoard
    chess.svg = types.ModuleType("chess.svg")
    sys.modules["chess"], sys.modules["chess.svg"] = chess, chess.svg
    for name in ("tlc", "cairosvg", "boto3", "timm", "PIL", "PIL.Image", "tqdm", "torchvision.transforms.v2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock()


def load_reference():
    """Import the two reference scripts unmodified.  sys.path is restored afterwards: a process that keeps /root/reference
    at its front would hand it to spawned children, which then import the reference's ``chessvision`` instead of ours."""
    import importlib
    install_standins()
    saved = list(sys.path)
    sys.path.insert(0, REF)
    try:
        ref_eval = importlib.import_module("scripts.eval.evaluate")
        ref_pipe = importlib.import_module("scripts.process_new_raw.process_pipeline")
    finally:
        sys.path[:] = saved
    assert ref_eval.__file__.startswith(REF) and ref_pipe.__file__.startswith(REF)
    return ref_eval, ref_pipe


def make_cases(seed=20261017):
    rng = np.random.default_rng(seed)
    fens = []
    for path in sorted(glob.glob(os.path.join(REF, "data/test/*/ground_truth/*.txt")))[:12]:
        fens.append(open(path).read().strip().split()[0])
    fens += ["8/8/8/8/8/8/8/8", "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR"]
    probs = []
    for i, _ in enumerate(fens):
        z = rng.normal(size=(64, 13)).astype(np.float32) * (1.0 + i % 3)
        p = np.exp(z - z.max(1, keepdims=True))
        p = (p / p.sum(1, keepdims=True)).astype(np.float32)
        # no exact ties: among equal probabilities the reference's order is whatever np.argsort's (vectorised,
        # machine-dependent, unstable) quicksort yields, so ties cannot be pinned; oracle and CUDA path define them
        # as the stable order (oracle/metrics.py:topk_hits)
        probs.append(p)
    pred_fens = [fens[(i + 3) % len(fens)] for i in range(len(fens))]
    # logits-like arrays for the quality scores: smooth blobs, plain noise, values on bin edges, out-of-range values
    yy, xx = np.mgrid[0:256, 0:256].astype(np.float32)
    arrays = []
    for i in range(6):
        cx, cy, rad = rng.uniform(80, 176), rng.uniform(80, 176), rng.uniform(40, 110)
        d = rad - np.maximum(np.abs(xx - cx), np.abs(yy - cy))
        a = (d * rng.uniform(0.05, 0.4) + rng.normal(scale=0.5, size=d.shape)).astype(np.float32)
        if i % 2:
            a = (1.0 / (1.0 + np.exp(-a))).astype(np.float32)      # probabilities in (0,1)
        arrays.append(a)
    edge = rng.choice(np.linspace(0, 1, 11).astype(np.float32), size=(256, 256)).astype(np.float32)
    arrays.append(edge)
    arrays.append(rng.uniform(-0.5, 1.5, size=(256, 256)).astype(np.float32))
    quads = []
    for i in range(10):
        base = np.array([[400, 100], [100, 110], [90, 420], [410, 400]], np.float32) + rng.uniform(-60, 60, size=(4, 2)).astype(np.float32)
        quads.append(base.reshape(4, 1, 2).astype(np.float32))
    quads.append(np.array([[10, 10], [10, 10], [10, 10], [10, 10]], np.float32).reshape(4, 1, 2))   # degenerate
    return fens, probs, pred_fens, arrays, quads


def completeness_cases(seed=7):
    """256x256 masks for mask_completeness: holes, nested components, specks, thin lines, border contact, ties, empty, full."""
    import cv2
    rng = np.random.default_rng(seed)
    out = []
    m = np.zeros((256, 256), np.float32); cv2.circle(m, (128, 128), 90, 1, -1); cv2.circle(m, (128, 128), 50, 0, -1); cv2.circle(m, (128, 128), 20, 1, -1)
    m[rng.random(m.shape) > 0.985] = 1
    out.append(m)                                                      # ring with a nested blob + specks
    out.append((rng.random((256, 256)) > 0.6).astype(np.float32))      # dense noise: hundreds of components
    out.append((rng.random((256, 256)) > 0.3).astype(np.float32))      # one percolating component with many holes
    m = np.zeros((256, 256), np.float32)
    for _ in range(12):
        p = rng.integers(0, 256, 4); cv2.line(m, (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), 1, 1)
    out.append(m)                                                      # one-pixel-wide lines (doubly traversed borders)
    m = np.zeros((256, 256), np.float32); m[:, :40] = 1; m[200:, :] = 1; m[60:120, 100:160] = 1; m[80:100, 120:140] = 0
    out.append(m)                                                      # components touching the image border
    m = np.zeros((256, 256), np.float32); m[10:40, 10:40] = 1; m[100:130, 50:80] = 1; m[200:230, 200:230] = 1
    out.append(m)                                                      # three components of equal area: the tie rule
    m = np.zeros((256, 256), np.float32); m[5, 5] = m[9, 200] = m[250, 3] = 1
    out.append(m)                                                      # isolated pixels only (all areas 0)
    out.append(np.zeros((256, 256), np.float32))                       # empty
    out.append(np.ones((256, 256), np.float32))                        # full
    m = np.zeros((256, 256), np.float32); m[0, :] = m[-1, :] = m[:, 0] = m[:, -1] = 1; m[100:150, 100:150] = 1
    out.append(m)                                                      # a frame around everything: the inner blob is nested in its hole
    return out


def main():
    ref_eval, ref_pipe = load_reference()
    from oracle import metrics as om
    fens, probs, pred_fens, arrays, quads = make_cases()
    hits = np.zeros((len(fens), 5), np.int32)
    correct = np.zeros(len(fens), np.int32)
    for i, (fen, p, pf) in enumerate(zip(fens, probs, pred_fens)):
        r = ref_eval.compute_model_topk_accuracy(p, fen, k=5)
        hits[i] = np.round(np.array(r.accuracies) * 64).astype(np.int32)
        correct[i] = ref_eval.compute_position_accuracy(pf, fen).num_correct
        assert om.topk_hits(p, fen, 5) == hits[i].tolist(), (i, om.topk_hits(p, fen, 5), hits[i])
        assert om.position_correct(pf, fen) == correct[i]
        assert om.fen_to_labels(fen) == ref_eval.board_to_labels(sys.modules["chess"].BaseBoard(fen))
        assert om.label_indices(p, fen) == ref_eval.get_label_indices(p, fen)
        assert om.fen_to_indices(fen) == ref_eval.get_validated_indices(fen)
    import cv2
    ref_pipe.cv2 = cv2                                    # the stand-in loader may have mocked nothing here: cv2 is installed
    comp_arrays = arrays + completeness_cases()
    comp = np.array([ref_pipe.mask_completeness(a) for a in comp_arrays])
    for i, a in enumerate(comp_arrays):
        assert om.mask_completeness(a) == comp[i], (i, om.mask_completeness(a), comp[i])
    dist = np.array([ref_pipe.probability_distribution(a) for a in arrays])
    conf = np.array([ref_pipe.probability_confidence(a) for a in arrays])
    reg = np.array([ref_pipe.quadrangle_regularity(q) for q in quads])
    for i, a in enumerate(arrays):
        assert om.probability_distribution(a) == dist[i], (i, om.probability_distribution(a), dist[i])
        assert om.probability_confidence(a) == conf[i]
    for i, q in enumerate(quads):
        got = om.quadrangle_regularity(q)
        assert got == reg[i] or (np.isnan(got) and np.isnan(reg[i])), (i, got, reg[i])
    out = os.path.join(ROOT, "tests", "golden", "metrics_vectors.npz")
    np.savez_compressed(out, fens=np.array(fens), pred_fens=np.array(pred_fens), probs=np.stack(probs), topk_hits=hits,
                        correct=correct, arrays=np.stack(arrays).astype(np.float16 if False else np.float32), distribution=dist,
                        confidence=conf, quads=np.stack(quads), regularity=reg,
                        completeness_masks=np.packbits(np.stack(comp_arrays[len(arrays):]) > 0.5, axis=-1), completeness=comp)
    print("wrote", out, os.path.getsize(out), "bytes;", len(fens), "positions,", len(arrays), "arrays,", len(quads), "quadrangles")


if __name__ == "__main__":
    main()
