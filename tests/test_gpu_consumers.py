"""GPU parity of the output consumers (csrc/consumers.cu; SURVEY.md §8(f) n1 evaluation metrics, n4 quality scores)
against ``oracle/metrics.py`` and the golden vectors frozen from the unmodified reference functions
(``tests/golden/metrics_vectors.npz``).  Integer results bit-exact; floating-point scores within the stated tolerance."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import metrics as om

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vec():
    return np.load(GOLDEN / "metrics_vectors.npz")


def _truth(fens):
    return torch.tensor([om.fen_to_indices(str(f)) for f in fens], dtype=torch.uint8).cuda()


def test_topk_and_position_accuracy_golden(engine, vec):
    probs = torch.from_numpy(vec["probs"]).cuda()
    pred = _truth(vec["pred_fens"])
    hits, correct = engine.eval_metrics(probs, pred, pred, _truth(vec["fens"]), k=5)
    assert np.array_equal(hits.cpu().numpy(), vec["topk_hits"])
    assert np.array_equal(correct.cpu().numpy()[:, 0], vec["correct"]) and np.array_equal(correct.cpu().numpy()[:, 1], vec["correct"])


def test_topk_ties_and_flip_against_the_oracle(engine):
    rng = np.random.default_rng(3)
    fens = ["rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR", "8/8/8/8/8/8/8/8", "2b3k1/pp3pp1/8/1n1p3p/1b1P1B1P/1P3PP1/P4KN1/3B4"]
    probs = np.round(rng.dirichlet(np.ones(13), size=(3, 64)) * 6).astype(np.float32) / 6    # many exact ties
    labels = rng.integers(0, 13, (3, 64)).astype(np.uint8)
    for flip in (False, True):
        hits, correct = engine.eval_metrics(torch.from_numpy(probs).cuda(), torch.from_numpy(labels).cuda(), None, _truth(fens), k=13, flip=flip)
        for i, fen in enumerate(fens):
            assert hits[i].cpu().tolist() == om.topk_hits(probs[i], fen, 13)
            true = om.fen_to_indices(fen)
            seen = labels[i][::-1] if flip else labels[i]
            assert int(correct[i, 0]) == int(sum(int(a) == b for a, b in zip(seen, true))) and int(correct[i, 1]) == 0


def test_reference_known_answers_through_the_python_mirror(engine):
    """reference tests/test_metrics.py:49-105 through chessvision.evaluation (same names as scripts/eval/evaluate.py)."""
    from chessvision import constants, evaluation as ev
    LI = {s: i for i, s in enumerate(constants.LABEL_NAMES)}
    p = np.zeros((64, 13), np.float32)
    p[:32, LI["f"]] = 1.0
    p[32:48, LI["p"]], p[32:48, LI["f"]] = 1.0, 0.9
    p[48:, LI["P"]], p[48:, LI["p"]], p[48:, LI["f"]] = 1.0, 0.9, 0.8
    r = ev.compute_model_topk_accuracy(p, "8/8/8/8/8/8/8/8", k=3)
    assert isinstance(r, ev.TopKAccuracyResult) and r.k == 3 and len(r.accuracies) == 3
    assert (r.top_1, r.top_2, r.top_3) == (0.5, 0.75, 1.0)
    r1 = ev.compute_model_topk_accuracy(p, "8/8/8/8/8/8/8/8", k=1)
    assert r1.top_1 == 0.5 and r1.top_2 == 0.0
    acc = ev.compute_position_accuracy("8/8/8/8/4Q3/8/8/8", "8/8/8/8/4q3/8/8/8")
    assert acc.num_correct == 63 and acc.accuracy == 63 / 64 and acc.total_squares == 64
    assert ev.board_to_labels("8/8/8/8/4Q3/8/8/8")[36] == "Q"


def test_quality_scores_golden(engine, vec):
    arrays = torch.from_numpy(vec["arrays"]).cuda()
    scores = engine.quality_scores(arrays).cpu().numpy()
    # histogram counts are exact, the entropy is evaluated in float64 on both sides: 1e-12; the confidence is a float32
    # pairwise sum in numpy and a float64 sum here: 2e-6 relative
    assert np.allclose(scores[:, 2], vec["distribution"], rtol=0, atol=1e-12), (scores[:, 2], vec["distribution"])
    assert np.allclose(scores[:, 3], vec["confidence"], rtol=2e-6, atol=0), (scores[:, 3], vec["confidence"])
    masks = np.concatenate([vec["arrays"], np.unpackbits(vec["completeness_masks"], axis=-1).astype(np.float32)])
    comp = engine.quality_scores(torch.from_numpy(masks).cuda()).cpu().numpy()[:, 1]
    assert np.array_equal(comp, vec["completeness"]), (comp, vec["completeness"])     # a ratio of two pixel counts: exact
    quads = torch.from_numpy(vec["quads"].reshape(-1, 4, 2)).cuda()
    dummy = torch.zeros((quads.shape[0], 16), dtype=torch.float32, device="cuda")
    reg = engine.quality_scores(dummy, quads).cpu().numpy()[:, 0]
    want = vec["regularity"]
    ok = np.isclose(reg, want, rtol=0, atol=2e-6) | (np.isnan(reg) & np.isnan(want))   # float32 arithmetic, acosf within 2 ulp
    assert ok.all(), (reg, want)


def test_quality_scores_on_pipeline_logits(engine):
    """The shape the reference feeds (process_pipeline.py:288-291): the 256x256 logits of real boards."""
    rng = np.random.default_rng(8)
    logits = (rng.normal(size=(5, 256, 256)) * 6).astype(np.float32)
    logits[1] = np.clip(logits[1], 0, 1)            # everything inside the histogram range, many values on 0 and 1
    logits[2, :128] = 0.5                           # a large tie group straddling the top-quarter threshold
    scores = engine.quality_scores(torch.from_numpy(logits).cuda()).cpu().numpy()
    for i in range(5):
        assert scores[i, 1] == om.mask_completeness(logits[i])
        assert abs(scores[i, 2] - om.probability_distribution(logits[i])) <= 1e-12
        assert abs(scores[i, 3] - om.probability_confidence(logits[i])) <= 2e-6 * abs(scores[i, 3])


def test_probability_confidence_of_fewer_than_four_values(engine):
    """k = int(L * 0.25) is 0 for L < 4 and numpy's [-0:] is then the WHOLE array (process_pipeline.py:463-466), not an empty one."""
    for L in (1, 2, 3, 4, 5, 7, 8):
        v = np.random.default_rng(L).normal(size=(3, L)).astype(np.float32)
        got = engine.quality_scores(torch.from_numpy(v).cuda()).cpu().numpy()[:, 3]
        for i in range(3):
            want = om.probability_confidence(v[i])
            assert abs(got[i] - want) <= 2e-6 * abs(want), (L, got[i], want)


def test_data_test_accuracy_from_device_outputs_equals_the_oracle():
    """The reference's evaluation flow (scripts/eval/evaluate.py:227-330) end to end on the device — decode, image->FEN,
    metrics — against the oracle's metrics on the REFERENCE's outputs for the same 38 files (golden vectors)."""
    import json
    from conftest import WEIGHTS
    from chessvision import ChessVision, decode, evaluation
    man = json.load(open(GOLDEN / "manifest.json"))["images"]
    arr = np.load(GOLDEN / "reference_outputs.npz")
    cv = ChessVision(board_extractor_weights=str(WEIGHTS / "best_extractor.pth"), classifier_weights=str(WEIGHTS / "best_classifier.pth"),
                     classifier_model_id="resnet18", lazy_load=False, max_batch=64)
    eng = cv._engine
    files = [GOLDEN / "data_test" / e["file"] for e in man]
    fens = [e["ground_truth_fen"].split()[0] for e in man]
    out = eng.image_to_fen(decode.imread_batch(files, engine=eng), eng.alloc_outputs(len(files)))
    hits, correct = evaluation.evaluate_batch(out["probs"], out["labels"], out["labels_valid"], fens, k=3, engine=eng)
    found = out["found"].cpu().numpy().astype(bool)
    top1 = top3 = tot = 0
    for i, e in enumerate(man):
        assert found[i] == e["found"]
        if not e["found"]:
            continue
        want = om.topk_hits(arr[f"probs_{i}"].astype(np.float32), fens[i], 3)
        assert int(hits[i, 0]) == want[0], e["file"]                      # labels are bit-identical, so top-1 is too
        assert abs(int(hits[i, 1]) - want[1]) <= 1 and abs(int(hits[i, 2]) - want[2]) <= 1, e["file"]   # near-ties of 2nd/3rd rank
        assert int(correct[i, 0]) == om.position_correct(e["original_fen"], fens[i])
        assert int(correct[i, 1]) == om.position_correct(e["fen"], fens[i])
        top1 += int(hits[i, 0]); top3 += int(hits[i, 2]); tot += 64
    print(f"data/test from device outputs: top-1 {top1 / tot:.4f}, top-3 {top3 / tot:.4f} over {tot // 64} boards")


def test_mask_completeness_random_masks(engine):
    """Bit-exact against the oracle (itself equal to the reference function + live cv2) on masks built to break contour
    code: noise at every density, blobs with holes and islands, single-pixel bridges, everything on the image border."""
    from scipy import ndimage
    rng = np.random.default_rng(21)
    masks = []
    for i in range(24):
        kind = i % 4
        if kind == 0:
            m = rng.random((256, 256)) > 0.2 + 0.03 * i
        elif kind == 1:
            m = ndimage.gaussian_filter(rng.normal(size=(256, 256)), 2 + i % 5) > 0.01 * (i - 12)
        elif kind == 2:
            m = ndimage.gaussian_filter(rng.normal(size=(256, 256)), 6) > 0
            m ^= rng.random((256, 256)) > 0.995
        else:
            m = np.zeros((256, 256), bool)
            m[::2, ::2] = True                     # a lattice of isolated pixels ...
            m[100:140, 90:170] = True              # ... around one block, plus diagonal one-pixel bridges
            idx = np.arange(60)
            m[140 + idx, 170 + idx] = True
        masks.append(m.astype(np.float32))
    got = engine.quality_scores(torch.from_numpy(np.stack(masks)).cuda()).cpu().numpy()[:, 1]
    for i, m in enumerate(masks):
        assert got[i] == om.mask_completeness(m), i
