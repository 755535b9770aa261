"""Pins the oracle to the UNMODIFIED reference: (1) against the committed golden vectors that oracle/make_golden.py froze
from ``/root/reference`` running on its own ``data/test`` images, and (2) — only where the reference checkout exists,
i.e. in the build container — against the reference imported live.  CPU only."""
import json
import sys

import cv2
import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, WEIGHTS
from oracle import geometry as og


@pytest.fixture(scope="module")
def golden():
    man = json.load(open(GOLDEN / "manifest.json"))
    arr = np.load(GOLDEN / "reference_outputs.npz")
    return man, arr


def image(entry):
    return cv2.imread(str(GOLDEN / "data_test" / entry["file"]))


def test_golden_set_is_the_reference_test_set(golden):
    man, _ = golden
    assert len(man["images"]) == 38                       # data/test/initial (24) + data/test/2024-11-04-2024-11-04 (14)
    assert man["summary"]["found"] >= 30, "the locally trained extractor must find most boards or parity is vacuous"
    assert all(e["ground_truth_fen"] for e in man["images"])


def test_quads_from_reference_masks(golden):
    man, arr = golden
    for i, e in enumerate(man["images"]):
        mask = (np.unpackbits(arr[f"mask_{i}"]).reshape(256, 256) * 255).astype(np.uint8)
        quad = og.find_quadrangle(mask)
        assert (quad is not None) == e["found"], e["file"]
        if e["found"]:
            assert quad.reshape(4, 2).tolist() == e["quad"], e["file"]


def test_boards_from_reference_quads(golden):
    man, arr = golden
    for i, e in enumerate(man["images"][:12]):
        if not e["found"]:
            continue
        quad = np.array(e["quad"], np.int32).reshape(4, 1, 2)
        board = og.extract_board(image(e), og.scale_quadrangle(quad, (512, 512)))
        assert np.array_equal(board, arr[f"board_{i}"]), e["file"]


def test_fen_from_reference_probabilities(golden):
    man, arr = golden
    n_fix = 0
    for i, e in enumerate(man["images"]):
        if not e["found"]:
            continue
        fen, original_fen, labels, fixed, fixes = og.position_from_probabilities(arr[f"probs_{i}"], False)
        assert (fen, original_fen) == (e["fen"], e["original_fen"]), e["file"]
        assert [og.LABEL_NAMES.index(l) for l in labels] == arr[f"labels_{i}"].tolist()
        assert len(fixes) == e["n_fixes"]
        n_fix += len(fixes)
    print("validation fixes across data/test:", n_fix)


def test_networks_against_reference_outputs(golden):
    """fp32 oracle networks (oracle/nets.py) with the committed weights reproduce the reference's logits and
    probabilities; golden logits are stored as fp16, hence the 2^-11 relative tolerance."""
    from oracle.pipeline import OraclePipeline
    man, arr = golden
    ext, cls = WEIGHTS / "best_extractor.pth", WEIGHTS / "best_classifier.pth"
    assert ext.exists() and cls.exists(), "weights/ missing: run oracle/train_weights.py + oracle/make_golden.py"
    import hashlib
    assert hashlib.sha1(ext.read_bytes()).hexdigest() == man["weights"]["extractor_sha1"], "golden vectors were made with other weights"
    assert hashlib.sha1(cls.read_bytes()).hexdigest() == man["weights"]["classifier_sha1"]
    orc = OraclePipeline.from_checkpoints(str(ext), str(cls))
    done = 0
    for i, e in enumerate(man["images"]):
        if not e["found"] or done == 2:
            continue
        out = orc.process_image(image(e))
        ref = arr[f"logits_{i}"].astype(np.float32)
        assert np.abs(out["logits"] - ref).max() <= 2.0 ** -10 * max(1.0, np.abs(ref).max()) + 1e-4
        assert out["quad"].reshape(4, 2).tolist() == e["quad"]
        assert np.array_equal(out["board"], arr[f"board_{i}"])
        assert np.abs(out["probs"] - arr[f"probs_{i}"]).max() <= 1e-5
        assert (out["fen"], out["original_fen"]) == (e["fen"], e["original_fen"])
        done += 1
    assert done == 2


@pytest.mark.reference
def test_live_reference_equals_oracle(golden):
    """Import /root/reference/chessvision unmodified (python-chess and timm replaced by the stand-ins of
    oracle/make_golden.py) and compare one process_image call stage by stage with the oracle, bit for bit."""
    from oracle import ref_loader
    if ref_loader.reference_root() is None:
        pytest.skip("neither the reference checkout nor its staging oracle/_ref is present on this machine")
    import torch
    from oracle.pipeline import OraclePipeline
    man, _ = golden
    ext, cls = str(WEIGHTS / "best_extractor.pth"), str(WEIGHTS / "best_classifier.pth")
    saved_path, saved_mod = list(sys.path), sys.modules.pop("chessvision", None)
    saved_sub = {k: sys.modules.pop(k) for k in list(sys.modules) if k.startswith("chessvision.")}
    try:
        ref = ref_loader.load_reference()
        ref.core.utils.get_device = lambda: torch.device("cpu")
        cv = ref.ChessVision(board_extractor_weights=ext, classifier_weights=cls, classifier_model_id="resnet18", lazy_load=False)
        orc = OraclePipeline.from_checkpoints(ext, cls)
        e = next(x for x in man["images"] if x["found"])
        img = image(e)
        res, out = cv.process_image(img), orc.process_image(img)
        assert np.array_equal(res.board_extraction.probabilities, out["logits"])
        assert np.array_equal(res.board_extraction.binary_mask, out["mask"])
        assert np.array_equal((res.board_extraction.quadrangle / 2).astype(np.int32).reshape(4, 2), out["quad"].reshape(4, 2))
        assert np.array_equal(res.board_extraction.board_image, out["board"])
        assert np.array_equal(res.position.model_probabilities, out["probs"])
        assert (res.position.fen, res.position.original_fen) == (out["fen"], out["original_fen"])
        # the reference's own known-answer test (tests/test_chessvision.py:119-146) through the reference and the oracle
        board = np.repeat(np.repeat(np.arange(64, dtype=np.uint8).reshape(8, 8), 64, 0), 64, 1)
        assert np.array_equal(ref.ChessVision.extract_squares(board), og.extract_squares(board))
    finally:
        for k in [k for k in sys.modules if k == "chessvision" or k.startswith("chessvision.")]:
            del sys.modules[k]
        sys.path[:] = saved_path
        if saved_mod is not None:
            sys.modules["chessvision"] = saved_mod
        sys.modules.update(saved_sub)
        sys.modules.pop("chess", None)
        sys.modules.pop("timm", None)
