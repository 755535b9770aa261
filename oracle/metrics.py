"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's per-board consumers of the
image->FEN outputs — SURVEY.md §8(f) rows n1 and n4.

n1  evaluation metrics        scripts/eval/evaluate.py:37-140,406-440 (``compute_position_accuracy``, ``board_to_labels``,
                              ``compute_model_topk_accuracy``, ``get_label_indices``, ``get_validated_indices``)
n4  extraction quality scores scripts/process_new_raw/process_pipeline.py:357-467 (``probability_distribution``,
                              ``mask_completeness``, ``quadrangle_regularity``, ``probability_confidence``)

python-chess (``chess==1.11.2``, not installed) is only used by the reference to parse the piece-placement field of a
FEN; ``fen_to_labels`` restates that.  Pinned by ``tests/test_oracle_metrics.py``: the reference's own known answers
(tests/test_metrics.py:16-174), the golden vectors ``tests/golden/metrics_vectors.npz`` written by
``oracle/make_golden_metrics.py`` from the UNMODIFIED reference functions, and — where /root/reference exists — the live
reference functions on random inputs.
"""
from __future__ import annotations

import numpy as np

LABEL_NAMES = ["B", "K", "N", "P", "Q", "R", "b", "k", "n", "p", "q", "r", "f"]  # constants.py:23
LABEL_INDICES = {s: i for i, s in enumerate(LABEL_NAMES)}


def fen_to_labels(fen: str) -> list[str]:
    """``board_to_labels(chess.BaseBoard(fen))`` (evaluate.py:61-86): 64 piece symbols in FEN order a8..h8, a7..h1,
    "f" for an empty square.  Only the piece-placement field is read (anything after the first blank is ignored)."""
    rows = fen.split()[0].split("/")
    assert len(rows) == 8, f"expected 8 ranks in '{fen}'"
    labels: list[str] = []
    for row in rows:
        n = 0
        for ch in row:
            if ch.isdigit():
                labels.extend("f" * int(ch))
                n += int(ch)
            else:
                assert ch in LABEL_INDICES and ch != "f", f"invalid piece '{ch}' in '{fen}'"
                labels.append(ch)
                n += 1
        assert n == 8, f"rank '{row}' of '{fen}' does not have 8 squares"
    return labels


def fen_to_indices(fen: str) -> list[int]:
    """``get_validated_indices`` (evaluate.py:430-440)."""
    return [LABEL_INDICES[s] for s in fen_to_labels(fen)]


def position_correct(predicted_fen: str, true_fen: str) -> int:
    """``compute_position_accuracy(...).num_correct`` (evaluate.py:37-52): squares whose piece (or emptiness) agrees."""
    return sum(a == b for a, b in zip(fen_to_labels(predicted_fen), fen_to_labels(true_fen)))


def topk_hits(probabilities: np.ndarray, true_fen: str, k: int = 3) -> list[int]:
    """``compute_model_topk_accuracy`` (evaluate.py:109-140) before the division by 64: hits[i] = squares whose true label
    is among the i+1 most probable classes.  Ranking is ``np.argsort(axis=1)`` read from the end.  Among EQUAL
    probabilities the reference's order is numpy-implementation-defined (its default quicksort is vectorised per CPU
    and not stable: [probe] numpy 2.3 on this container's AVX-512 cores does not return the stable order), so ties are
    unpinned; here and in the CUDA kernel they are defined as the stable ascending order read from the end, i.e. the
    HIGHER class index ranks first."""
    true = fen_to_indices(true_fen)
    hits = [0] * k
    for sq in range(64):
        order = sorted(range(probabilities.shape[1]), key=lambda c: (probabilities[sq, c], c))  # ascending, stable
        for i in range(k):
            if order[-(i + 1)] == true[sq]:
                for j in range(i, k):
                    hits[j] += 1
                break
    return hits


def label_indices(probabilities: np.ndarray, true_fen: str) -> tuple[list[int], list[int]]:
    """``get_label_indices`` (evaluate.py:406-427): first maximum wins."""
    return np.argmax(probabilities, axis=1).tolist(), fen_to_indices(true_fen)


# ------------------------------------------------------------------------------------------------------------ n4
def probability_distribution(mask: np.ndarray) -> float:
    """process_pipeline.py:357-378: 1 - entropy(10-bin histogram over [0,1]) / log2(10).  ``np.histogram`` keeps the
    input dtype: for float32 input the bin index is int(v * 10) in float32, corrected against the float32 edges
    ``linspace(0, 1, 11)``; values outside [0,1] are dropped and the last bin is closed on the right."""
    hist, _ = np.histogram(np.asarray(mask).flatten(), bins=10, range=(0, 1))
    hist = hist / np.sum(hist)
    entropy = -np.sum(hist * np.log2(hist + 1e-10))
    return float(1.0 - (entropy / -np.log2(1 / 10)))


def external_components(binary: np.ndarray):
    """What ``cv2.findContours(binary, RETR_EXTERNAL, ...)`` returns, in ITS order, as (outer border points, first pixel):
    the outer borders of the 8-connected components that are not enclosed by a hole of another component (Suzuki-Abe
    border following; an outer border is top-level if the last border met on its row is the frame, or — transitively — a
    top-level outer border), listed in REVERSE discovery order [pinned against live cv2 in tests/test_oracle_metrics.py]."""
    from . import geometry as og
    h, w = binary.shape
    lab = np.zeros((h + 2, w + 2), np.int32)
    lab[1:-1, 1:-1] = binary != 0
    kind, out, nbd = {}, [], 1
    for y in range(1, h + 1):
        last = 0
        row = lab[y]
        for x in range(1, w + 1):
            p, prev = int(row[x]), int(row[x - 1])
            start = None
            if prev == 0 and p == 1:
                start = (x, False)
            elif p == 0 and prev >= 1:
                start = (x - 1, True)
                if prev > 1:
                    last = prev
            if start is not None:
                nbd += 1
                pts, _ = og._follow(lab, start[0], y, nbd, start[1])
                if start[1]:
                    kind[nbd] = (True, False)
                else:
                    ext = True if last == 0 else (False if kind[last][0] else kind[last][1])
                    kind[nbd] = (False, ext)
                    if ext:
                        out.append((pts, (start[0] - 1, y - 1)))
            v = int(row[x])
            if v != 0 and v != 1:
                last = abs(v)
    return out[::-1]


def mask_completeness(mask: np.ndarray) -> float:
    """process_pipeline.py:380-414: (#pixels > 0.5) / (#pixels of the filled largest external contour).  ``max(contours,
    key=cv2.contourArea)`` keeps the first of equal areas; ``drawContours(..., thickness=-1)`` of an outer border paints the
    component and everything it encloses (= the complement of the background that stays 4-connected to the frame when
    only that component blocks)."""
    from scipy import ndimage
    from . import geometry as og
    binary = np.asarray(mask) > 0.5
    ext = external_components(binary.astype(np.uint8))
    if not ext:
        return 0.0
    areas = [og.contour_area(p) for p, _ in ext]
    sx, sy = ext[int(np.argmax(areas))][1]
    lab8, _ = ndimage.label(binary, structure=np.ones((3, 3)))
    filled = ndimage.binary_fill_holes(lab8 == lab8[sy, sx])
    return float(binary.sum()) / float(filled.sum())


def probability_confidence(probabilities: np.ndarray) -> float:
    """process_pipeline.py:459-467: mean |p - 0.5| * 2 over the top quarter of the values (largest p)."""
    flat = np.asarray(probabilities).ravel()
    k = int(flat.size * 0.25)
    top = np.sort(flat)[-k:]
    return float(np.mean(np.abs(top - 0.5)) * 2)


def quadrangle_regularity(quadrangle: np.ndarray | None) -> float:
    """process_pipeline.py:416-456: 1 - (cv(side lengths) + std(corner angles)/(pi/2)) / 2."""
    if quadrangle is None:
        return 0.0
    q = np.asarray(quadrangle).copy().squeeze(1)
    sides = [np.sqrt(((q[i] - q[(i + 1) % 4]) ** 2).sum()) for i in range(4)]
    angles = []
    for i in range(4):
        v1, v2 = q[(i - 1) % 4] - q[i], q[(i + 1) % 4] - q[i]
        norm = np.linalg.norm(v1) * np.linalg.norm(v2)
        angles.append(np.arccos(np.dot(v1, v2) / norm) if norm > 0 else 0)
    side_var = np.std(sides) / np.mean(sides) if np.mean(sides) > 0 else 1.0
    angle_var = np.std(angles) / (np.pi / 2)
    return float(1.0 - (side_var * 0.5 + angle_var * 0.5))
