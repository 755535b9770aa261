"""Host-side logic of the drop-in ``chessvision`` package that needs no GPU: constants, result types, checkpoint loader,
and the static helpers of ``ChessVision`` that the reference also computes on the host (core.py:310-355, 382-469).
Known answers come from the reference's own tests (tests/test_chessvision.py:25-42,119-146; tests/test_metrics.py)."""
import numpy as np
import pytest
import torch

from chessvision import ChessVision, constants, cv_types, utils
from oracle import geometry as og


def test_constants_match_reference():
    assert constants.INPUT_SIZE == (256, 256) and constants.BOARD_SIZE == (512, 512) and constants.PIECE_SIZE == (64, 64)
    assert constants.NUM_CLASSES == 13
    assert constants.LABEL_NAMES == ["B", "K", "N", "P", "Q", "R", "b", "k", "n", "p", "q", "r", "f"]       # constants.py:23
    assert constants.SQUARE_NAMES_NORMAL[:8] == ["a8", "b8", "c8", "d8", "e8", "f8", "g8", "h8"]          # constants.py:109-118
    assert constants.SQUARE_NAMES_NORMAL[-1] == "h1" and constants.SQUARE_NAMES_FLIPPED[0] == "h1"
    assert constants.INVALID_PAWN_SQUARES == {f + r for f in "abcdefgh" for r in "18"}                       # constants.py:88-105
    assert constants.SQUARE_NAMES_NORMAL == og.SQUARE_NAMES_NORMAL and constants.LABEL_NAMES == og.LABEL_NAMES


def test_lazy_initialisation_attributes():
    """tests/test_chessvision.py:25-42 of the reference."""
    cvm = ChessVision()
    assert cvm._board_extractor is None and cvm._classifier is None
    assert cvm._board_extractor_weights == constants.BEST_EXTRACTOR_WEIGHTS
    assert cvm._classifier_weights == constants.BEST_CLASSIFIER_WEIGHTS
    cvm = ChessVision(board_extractor_weights="a.pth", classifier_weights="b.pth", classifier_model_id="resnet18")
    assert cvm._board_extractor_weights == "a.pth" and cvm._classifier_weights == "b.pth" and cvm._classifier_model_id == "resnet18"


def test_extract_squares_known_answer():
    board = np.zeros((512, 512), np.uint8)
    for rank in range(8):
        for file in range(8):
            board[rank * 64:(rank + 1) * 64, file * 64:(file + 1) * 64] = rank * 8 + file
    squares = ChessVision.extract_squares(board)
    assert squares.shape == (64, 64, 64, 1)
    for i in (0, 7, 8, 15, 16, 23, 56, 63):
        assert squares[i, 0, 0, 0] == i
    rng = np.random.default_rng(0)
    b = rng.integers(0, 256, (512, 512), dtype=np.uint8)
    assert np.array_equal(ChessVision.extract_squares(b), og.extract_squares(b))


def test_position_from_probabilities_equals_oracle():
    rng = np.random.default_rng(1)
    for trial in range(20):
        probs = rng.dirichlet(np.ones(13) * 0.3, 64).astype(np.float32)
        if trial % 2:
            probs[rng.integers(0, 8, 3), 3] = 2.0      # force pawns on rank 8
            probs[56 + rng.integers(0, 8, 3), 9] = 2.0  # and on rank 1
        for flip in (False, True):
            names = constants.SQUARE_NAMES_FLIPPED if flip else constants.SQUARE_NAMES_NORMAL
            res = ChessVision.process_position_probabilities(probs, names, np.zeros((64, 64, 64, 1), np.uint8))
            fen, original_fen, labels, fixed, fixes = og.position_from_probabilities(probs, flip)
            assert (res.fen, res.original_fen) == (fen, original_fen)
            assert [(f.square_name, f.original_piece, f.corrected_piece, f.rule_name) for f in res.validation_fixes] == fixes
            assert (res.original_fen != res.fen) == bool(res.validation_fixes)
            assert isinstance(res, cv_types.PositionResult) and res.model_probabilities is probs


def test_validate_position_mutates_in_place():
    probs = np.full((64, 13), 0.01, np.float32)
    probs[0, 3], probs[0, 4] = 0.9, 0.5
    labels = ["P"] + ["f"] * 63
    out, fixes = ChessVision.validate_position(labels, probs, constants.SQUARE_NAMES_NORMAL)
    assert out is labels and labels[0] == "Q"
    assert fixes == [cv_types.ValidationFix("a8", "P", "Q", "no_pawns_on_ends")]


def test_quadrangle_helpers():
    q = np.array([[[200, 50]], [[60, 52]], [[58, 210]], [[205, 208]]], np.int32)
    assert np.array_equal(ChessVision._rotate_quadrangle(q), q)
    r = q[[1, 2, 3, 0]]
    assert np.array_equal(ChessVision._rotate_quadrangle(r), r[[3, 0, 1, 2]])
    s = ChessVision._scale_quadrangle(q, (512, 768))           # the height scales both axes (core.py:416)
    assert s.dtype == np.float32 and np.array_equal(s, q * 2.0)
    assert np.array_equal(s, og.scale_quadrangle(q, (512, 768)))


def test_filter_contours_thresholds():
    big = np.array([[[10, 10]], [[10, 240]], [[240, 240]], [[240, 10]]], np.int32)       # 80% of the mask, square
    small = np.array([[[10, 10]], [[10, 60]], [[60, 60]], [[60, 10]]], np.int32)         # 3.8%
    thin = np.array([[[0, 0]], [[0, 255]], [[120, 255]], [[120, 0]]], np.int32)          # 47% but 121x256 -> ratio 0.47
    kept = ChessVision._filter_contours((256, 256), [big, small, thin])
    assert len(kept) == 1 and kept[0] is big
    assert utils.ratio(3, 4) == 0.75 and utils.ratio(4, 3) == 0.75 and utils.ratio(0, 0) == -1   # utils.py:89-93


def test_create_binary_mask_asserts_and_values():
    p = np.array([[0.2, 0.5, 0.50001, 1.0]], np.float32)
    assert np.array_equal(utils.create_binary_mask(p, 0.5), np.array([[0, 0, 255, 255]], np.uint8))
    with pytest.raises(AssertionError):
        utils.create_binary_mask(p.astype(np.float64), 0.5)
    with pytest.raises(AssertionError):
        utils.create_binary_mask(p, 1.5)


def test_checkpoint_loader_accepts_every_reference_layout(tmp_path):
    """utils.load_model_checkpoint (utils.py:55-86): model_state_dict / state_dict / model / bare dict, plus metadata."""
    sd = {"conv.weight": torch.randn(4, 3, 3, 3), "bn.num_batches_tracked": torch.tensor(7)}
    layouts = {
        "a.pth": {"model_state_dict": sd, "metadata": {"epochs": 3}},
        "b.pth": {"state_dict": sd},
        "c.pth": {"model": sd},
        "d.pth": sd,
    }
    for name, blob in layouts.items():
        torch.save(blob, tmp_path / name)
        got, meta = utils.load_state_dict(str(tmp_path / name))
        assert set(got) == set(sd) and torch.equal(got["conv.weight"], sd["conv.weight"])
        assert meta == ({"epochs": 3} if name == "a.pth" else {})
    with pytest.raises((AssertionError, FileNotFoundError)):
        utils.load_state_dict(str(tmp_path / "missing.pth"))


def test_process_image_input_asserts():
    cvm = ChessVision(board_extractor_weights="a.pth", classifier_weights="b.pth")
    with pytest.raises(AssertionError):
        cvm.process_image([[1, 2, 3]])
    with pytest.raises(AssertionError):
        cvm.process_image(np.zeros((512, 512, 3), np.float32))
    with pytest.raises(AssertionError):
        cvm.process_image(np.zeros((512, 512), np.uint8))
    with pytest.raises(AssertionError):
        ChessVision.process_board_extraction_logits(np.zeros((256, 256), np.float64), np.zeros((512, 512, 3), np.uint8), 0.5)
    with pytest.raises(AssertionError):
        ChessVision.process_board_extraction_logits(np.zeros((256, 256), np.float32), np.zeros((512, 512, 3), np.uint8), 1.5)


def test_checkpoint_writers_and_strip_optimizer(tmp_path):
    """train_unet.py:31-40 / train_classifier.py:112-123 / strip_optimizer.py:15-47: the three-key checkpoint, its stripped form,
    and utils.load_model_checkpoint reading every layout into a drop-in module (parameter names of timm's resnet18)."""
    from chessvision import checkpoints
    from chessvision.modules import NativeResNet18, NativeUNet
    model = utils.get_classifier_model("resnet18")
    assert isinstance(model, NativeResNet18) and sum(p.numel() for p in model.parameters()) == 11_176_909   # notebooks/model-summary.ipynb:233
    assert sum(p.numel() for p in NativeUNet().parameters()) == 31_037_633
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    path = tmp_path / "cls.pth"
    checkpoints.save_classifier_checkpoint(model, str(path), opt, {"epochs": 2})
    blob = torch.load(path, map_location="cpu")
    assert set(blob) == {"model_state_dict", "optimizer_state_dict", "metadata"} and checkpoints.checkpoint_layout(str(path)) == "model_state_dict"
    checkpoints.strip_optimizer(str(path), str(tmp_path / "stripped.pth"))
    stripped = torch.load(tmp_path / "stripped.pth", map_location="cpu")
    assert set(stripped) == {"model_state_dict", "metadata"} and stripped["metadata"] == {"epochs": 2}
    fresh = utils.load_model_checkpoint(utils.get_classifier_model(), str(tmp_path / "stripped.pth"), torch.device("cpu"))
    assert fresh.metadata == {"epochs": 2}
    for (ka, va), (kb, vb) in zip(model.state_dict().items(), fresh.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    torch.save({"state_dict": model.state_dict(), "optimizer": 1}, tmp_path / "timm.pth")
    checkpoints.strip_optimizer(str(tmp_path / "timm.pth"))                       # in place
    assert set(torch.load(tmp_path / "timm.pth", map_location="cpu")) == {"state_dict", "metadata"}
    torch.save({"model": model.state_dict()}, tmp_path / "legacy.pth")
    checkpoints.strip_optimizer(str(tmp_path / "legacy.pth"))                     # unexpected layout: left as it is
    assert set(torch.load(tmp_path / "legacy.pth", map_location="cpu")) == {"model"}
    with pytest.raises(RuntimeError):
        model.train()                                                            # inference modules: training goes through chessvision.training
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            model(torch.zeros(64, 1, 64, 64))                                     # no CPU fallback behind the module either
