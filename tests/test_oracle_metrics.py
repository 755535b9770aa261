"""Pins ``oracle/metrics.py``: (1) the reference's own known answers for the metric functions
(``tests/test_metrics.py:16-174`` of the reference, restated on FEN strings), (2) the golden vectors that
``oracle/make_golden_metrics.py`` froze from the UNMODIFIED reference functions, (3) — only where the reference checkout
exists — the live reference functions.  CPU only."""
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE
from oracle import metrics as om

LI = om.LABEL_INDICES


@pytest.fixture(scope="module")
def vec():
    return np.load(GOLDEN / "metrics_vectors.npz")


def test_board_to_labels_known_answers():
    """reference tests/test_metrics.py:16-46."""
    labels = om.fen_to_labels("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1")
    assert labels[:8] == ["r", "n", "b", "q", "k", "b", "n", "r"]
    assert labels[8:16] == ["p"] * 8 and labels[16:48] == ["f"] * 32 and labels[48:56] == ["P"] * 8
    assert labels[56:] == ["R", "N", "B", "Q", "K", "B", "N", "R"]
    assert all(l == "f" for l in om.fen_to_labels("8/8/8/8/8/8/8/8"))
    labels = om.fen_to_labels("8/8/8/8/4Q3/8/8/8")                       # a queen on e4
    assert labels[4 * 8 + 4] == "Q" and sum(l != "f" for l in labels) == 1


def test_topk_known_answers():
    """reference tests/test_metrics.py:49-105."""
    p = np.zeros((64, 13), np.float32)
    p[:32, LI["f"]] = 1.0
    p[32:48, LI["p"]], p[32:48, LI["f"]] = 1.0, 0.9
    p[48:, LI["P"]], p[48:, LI["p"]], p[48:, LI["f"]] = 1.0, 0.9, 0.8
    assert om.topk_hits(p, "8/8/8/8/8/8/8/8", 3) == [32, 48, 64]
    p = np.zeros((64, 13), np.float32)
    p[48:56, LI["P"]] = 1.0
    p[list(range(48)) + list(range(56, 64)), LI["f"]] = 1.0
    assert om.topk_hits(p, "8/8/8/8/8/8/PPPPPPPP/8", 1) == [64]
    assert om.topk_hits(p, "8/8/8/8/8/8/PPPPPPPP/8", 5) == [64] * 5


def test_topk_with_errors_known_answers():
    """reference tests/test_metrics.py:137-174."""
    fen = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR"
    true = om.fen_to_labels(fen)
    p = np.zeros((64, 13), np.float32)
    for sq, lab in enumerate(true):
        if sq < 8:
            p[sq, LI["p"]], p[sq, LI["q"]], p[sq, LI[lab]] = 0.9, 0.8, 0.7
        elif sq >= 56:
            p[sq, LI["P"]], p[sq, LI[lab]], p[sq, LI["Q"]] = 0.9, 0.8, 0.7
        else:
            p[sq, LI[lab]], p[sq, LI["f"]], p[sq, LI["p"]] = 0.9, 0.8, 0.7
    assert om.topk_hits(p, fen, 3) == [40, 57, 64]          # expected_top1/2/3 of the reference test


def test_golden_vectors_from_the_reference(vec):
    for i, fen in enumerate(vec["fens"]):
        assert om.topk_hits(vec["probs"][i], str(fen), 5) == vec["topk_hits"][i].tolist()
        assert om.position_correct(str(vec["pred_fens"][i]), str(fen)) == int(vec["correct"][i])
    for i, a in enumerate(vec["arrays"]):
        assert om.probability_distribution(a) == vec["distribution"][i]
        assert om.probability_confidence(a) == vec["confidence"][i]
    masks = [a for a in vec["arrays"]] + [m.astype(np.float32) for m in np.unpackbits(vec["completeness_masks"], axis=-1)]
    for i, m in enumerate(masks):
        assert om.mask_completeness(m) == vec["completeness"][i], i
    for i, q in enumerate(vec["quads"]):
        got, want = om.quadrangle_regularity(q), vec["regularity"][i]
        assert got == want or (np.isnan(got) and np.isnan(want))
    assert om.quadrangle_regularity(None) == 0.0


@pytest.mark.reference
@pytest.mark.skipif(not (REFERENCE / "scripts" / "eval" / "evaluate.py").exists(), reason="reference checkout not present")
def test_against_the_live_reference():
    from oracle import make_golden_metrics as gen
    ref_eval, ref_pipe = gen.load_reference()
    rng = np.random.default_rng(5)
    fens, *_ = gen.make_cases(seed=99)
    for fen in fens[:6]:
        p = rng.dirichlet(np.ones(13), size=64).astype(np.float32)
        r = ref_eval.compute_model_topk_accuracy(p, fen, k=3)
        assert om.topk_hits(p, fen, 3) == [round(a * 64) for a in r.accuracies]
        a = rng.normal(size=(64, 64)).astype(np.float32)
        assert om.probability_distribution(a) == ref_pipe.probability_distribution(a)
        assert om.probability_confidence(a) == ref_pipe.probability_confidence(a)
    import cv2
    ref_pipe.cv2 = cv2
    for seed in range(12):                                  # small random masks: noise at several densities, blobs
        r = np.random.default_rng(seed)
        m = (r.random((48, 64)) > (0.3 + 0.05 * seed)).astype(np.float32)
        assert om.mask_completeness(m) == ref_pipe.mask_completeness(m), seed
