"""Pins the integer/geometry oracle (oracle/geometry.py) against the live OpenCV of this image — the un-vendored
third-party library in which the reference's arithmetic for these stages lives (core.py:212,299,300,360,373-374,394,398;
utils.py:131-132) — and against the reference's own known-answer tests.  CPU only.
"""
import cv2
import numpy as np
import pytest

import cvb_synth as synth
from oracle import geometry as og


def ref_find_quadrangle(mask):
    """The reference's _find_quadrangle/_filter_contours/_rotate_quadrangle call sequence (core.py:358-411) on cv2."""
    contours, _ = cv2.findContours(mask, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_TC89_KCOS)
    if len(contours) > 1:
        kept = []
        area_all = float(mask.shape[0] * mask.shape[1])
        for c in contours:
            a = cv2.contourArea(c) / area_all
            if a < 0.35 or a > 1.0:
                continue
            _, _, w, h = cv2.boundingRect(c)
            if min(w, h) / max(w, h) < 0.6:
                continue
            kept.append(c)
        contours = kept
    for c in contours:
        approx = cv2.approxPolyDP(c, 0.1 * cv2.arcLength(c, True), True)
        if len(approx) == 4:
            if approx[0, 0, 0] < approx[2, 0, 0]:
                approx = approx[[3, 0, 1, 2]]
            return approx
    return None


def test_resize_area_half_equals_cv2():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    assert np.array_equal(og.resize_area_half(img), cv2.resize(img, (256, 256), interpolation=cv2.INTER_AREA))
    img[:] = 255
    assert np.array_equal(og.resize_area_half(img), cv2.resize(img, (256, 256), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("h,w", [(600, 800), (1024, 768), (768, 768), (513, 512), (256, 300), (1080, 1920), (256, 256), (1024, 1024), (257, 999)])
def test_resize_area_general_equals_cv2(h, w):
    """Every code path of cv2.resize(INTER_AREA): integer factors (1x1, 2x2, 3x3, 4x4), fractional, mixed."""
    img = np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(og.resize_area(img, (256, 256)), cv2.resize(img, (256, 256), interpolation=cv2.INTER_AREA))


def test_binary_mask_follows_fp32_sigmoid():
    import torch
    x = np.array([[-1.0, -1e-8, 0.0, 5e-8, 8.9e-8, 9.0e-8, 1.2e-7, 1.0]], np.float32)
    want = np.where(torch.sigmoid(torch.from_numpy(x)).numpy() > 0.5, 255, 0).astype(np.uint8)
    assert np.array_equal(og.binary_mask(x, 0.5), want)
    rng = np.random.default_rng(1)
    x = rng.normal(0, 3, (256, 256)).astype(np.float32)
    for thr in (0.0, 0.3, 0.5, 0.9, 1.0):
        want = np.where(torch.sigmoid(torch.from_numpy(x)).numpy() > thr, 255, 0).astype(np.uint8)
        assert np.array_equal(og.binary_mask(x, thr), want)


def test_contours_equal_cv2_kcos():
    masks = synth.mask_suite(seed=3, n=40)
    total = 0
    for i, m in enumerate(masks):
        want, _ = cv2.findContours(m, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_TC89_KCOS)
        got = og.find_contours_ccomp_kcos(m)
        assert len(got) == len(want), f"mask {i}: {len(got)} contours, cv2 finds {len(want)}"
        for a, b in zip(got, want):
            assert np.array_equal(np.asarray(a).reshape(-1, 2), b.reshape(-1, 2)), f"mask {i}: contour differs"
            total += 1
    assert total > 60


def test_contour_measures_equal_cv2():
    masks = synth.mask_suite(seed=5, n=24)
    n = 0
    for m in masks:
        for c in cv2.findContours(m, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_TC89_KCOS)[0]:
            pts = c.reshape(-1, 2)
            assert og.contour_area(pts) == cv2.contourArea(c)
            assert tuple(og.bounding_rect(pts)) == tuple(cv2.boundingRect(c))
            assert og.arc_length_closed(pts) == cv2.arcLength(c, True)
            for frac in (0.1, 0.02):
                eps = frac * cv2.arcLength(c, True)
                want = cv2.approxPolyDP(c, eps, True).reshape(-1, 2)
                assert np.array_equal(og.approx_poly_closed(pts, eps).reshape(-1, 2), want)
            n += 1
    assert n > 40


def test_find_quadrangle_equals_reference_sequence():
    masks = synth.mask_suite(seed=7, n=96)
    found = 0
    for i, m in enumerate(masks):
        want, got = ref_find_quadrangle(m), og.find_quadrangle(m)
        assert (want is None) == (got is None), f"mask {i}"
        if want is not None:
            found += 1
            assert np.array_equal(got.reshape(4, 2), want.reshape(4, 2)), f"mask {i}"
    assert found > 40


def test_edge_masks():
    empty = np.zeros((256, 256), np.uint8)
    assert og.find_quadrangle(empty) is None
    full = np.full((256, 256), 255, np.uint8)
    want = ref_find_quadrangle(full)
    got = og.find_quadrangle(full)
    assert (want is None) == (got is None)
    if want is not None:
        assert np.array_equal(got.reshape(4, 2), want.reshape(4, 2))


def test_perspective_and_warp_equal_cv2():
    rng = np.random.default_rng(11)
    for _ in range(4):
        img, q = synth.board_image(rng)
        scaled = og.scale_quadrangle(q.reshape(4, 1, 2), (512, 512))
        assert scaled.dtype == np.float32 and np.array_equal(scaled.reshape(4, 2), q * 2.0)
        dest = np.array([[0, 0], [512, 0], [512, 512], [0, 512]], np.float32)       # utils.py:127-128
        M = cv2.getPerspectiveTransform(scaled.reshape(4, 2), dest)
        assert np.allclose(og.perspective_matrix(scaled.reshape(4, 2), dest), M, rtol=0, atol=1e-9 * np.abs(M).max())
        want = cv2.warpPerspective(img, M, (512, 512))
        assert np.array_equal(og.warp_perspective_u8(img, M, (512, 512)), want)
        gray = cv2.flip(cv2.cvtColor(want, cv2.COLOR_BGR2GRAY), 1)                    # core.py:299-300
        assert np.array_equal(og.extract_board(img, scaled), gray)


def test_warp_partially_outside_image():
    rng = np.random.default_rng(13)
    img = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    quad = np.array([[270, -10], [-12, 6], [4, 262], [250, 240]], np.int32)    # corners outside the image: BORDER_CONSTANT 0
    scaled = og.scale_quadrangle(quad.reshape(4, 1, 2), (512, 512))
    M = cv2.getPerspectiveTransform(scaled.reshape(4, 2), np.array([[0, 0], [512, 0], [512, 512], [0, 512]], np.float32))
    want = cv2.flip(cv2.cvtColor(cv2.warpPerspective(img, M, (512, 512)), cv2.COLOR_BGR2GRAY), 1)
    assert np.array_equal(og.extract_board(img, scaled), want)


def test_gray_equals_cv2():
    rng = np.random.default_rng(17)
    img = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    assert np.array_equal(og.bgr_to_gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def test_extract_squares_known_answer():
    """The reference's only known-answer test on the hot path (tests/test_chessvision.py:119-146)."""
    board = np.zeros((512, 512), np.uint8)
    for rank in range(8):
        for file in range(8):
            board[rank * 64:(rank + 1) * 64, file * 64:(file + 1) * 64] = rank * 8 + file
    squares = og.extract_squares(board)
    assert squares.shape == (64, 64, 64, 1)
    for i in (0, 7, 8, 15, 16, 23, 56, 63):
        assert squares[i, 0, 0, 0] == i


START_LABELS = list("rnbqkbnr") + ["p"] * 8 + ["f"] * 32 + ["P"] * 8 + list("RNBQKBNR")


def test_board_fen_known_answers():
    """FEN <-> label order as pinned by the reference's tests/test_metrics.py:19-46 and data/test ground truth."""
    assert og.board_fen(START_LABELS, og.SQUARE_NAMES_NORMAL) == "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR"
    labels = ["f"] * 64
    labels[36] = "P"   # e4 <-> index 36 (tests/test_metrics.py:44-46)
    assert og.SQUARE_NAMES_NORMAL[36] == "e4"
    assert og.board_fen(labels, og.SQUARE_NAMES_NORMAL) == "8/8/8/8/4P3/8/8/8"
    assert og.board_fen(START_LABELS, og.SQUARE_NAMES_FLIPPED) == "RNBKQBNR/PPPPPPPP/8/8/8/8/pppppppp/rnbkqbnr"


def test_rule_one_fix():
    """validate_position rule 1 (core.py:451-469): a pawn on rank 1/8 becomes the most probable non-pawn class."""
    probs = np.full((64, 13), 0.01, np.float32)
    probs[:, 12] = 0.5                      # empty everywhere
    probs[3, 3], probs[3, 9], probs[3, 4] = 0.9, 0.8, 0.7     # d8: P, then p, then Q
    probs[60, 9], probs[60, 12] = 0.95, 0.02                  # e1: p, then a tie among the rest
    fen, original_fen, labels, fixed, fixes = og.position_from_probabilities(probs, False)
    assert original_fen == "3P4/8/8/8/8/8/8/4p3"
    assert labels[3] == "P" and fixed[3] == "Q"
    assert fixed[60] not in ("P", "p")
    assert fen.startswith("3Q4/")
    assert len(fixes) == 2 and fixes[0][0] == "d8"


@pytest.mark.parametrize("h,w", [(100, 100), (128, 128), (255, 255), (200, 300), (300, 200), (64, 512), (37, 211), (256, 100), (1, 1), (2, 3), (250, 1000)])
def test_resize_area_with_an_enlarged_axis_equals_cv2(h, w):
    """INTER_AREA below 256 px (core.py:212 resizes any input): OpenCV's fixed-point bilinear emulation, bit for bit."""
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(og.resize_area(img, (256, 256)), cv2.resize(img, (256, 256), interpolation=cv2.INTER_AREA))


def test_perspective_transform_and_warp_for_any_output_size_equal_cv2():
    """utils.extract_perspective (utils.py:115-132) with float corners and output sizes other than the 512x512 board: the
    Point2f products of getPerspectiveTransform are float32, the warp blocks are min(1024 / min(16, h), w) wide."""
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    sizes = [(512, 512), (300, 200), (64, 64), (50, 30), (100, 10), (1000, 8), (33, 77), (640, 480), (5, 5), (17, 300)]
    for t in range(40):
        ow, oh = sizes[t % len(sizes)]
        quad = np.array([[40, 30], [350, 60], [330, 260], [60, 240]], np.float32) + rng.uniform(-60, 60, (4, 2)).astype(np.float32)
        dest = np.array(((0, 0), (ow, 0), (ow, oh), (0, oh)), np.float32)
        M = cv2.getPerspectiveTransform(quad, dest)
        Mo = og.perspective_matrix(quad, dest)
        assert np.array_equal(M, Mo), (ow, oh)
        assert np.array_equal(og.warp_perspective_u8(img, Mo, (ow, oh)), cv2.warpPerspective(img, M, (ow, oh))), (ow, oh)
    gray = img[:, :, 1].copy()
    assert np.array_equal(og.warp_perspective_u8(gray, Mo, (ow, oh)), cv2.warpPerspective(gray, M, (ow, oh)))


def test_cv2_spelling_of_find_quadrangle_equals_the_cv2_free_oracle():
    import cvb_synth as synth
    masks = synth.fuzz_masks(seed=3, n=48)
    n = 0
    for m in masks:
        a, b = og.find_quadrangle(m), og.find_quadrangle_cv2(m)
        assert (a is None) == (b is None)
        if a is not None:
            n += 1
            assert np.array_equal(a, b)
    assert n >= 15


def test_reference_ground_truth_masks_golden():
    """tests/golden/gt_masks.npz (the reference's 631 GT masks + what its unmodified _find_quadrangle returns): the oracle
    reproduces every quadrangle (a sample with the cv2-free restatement, all of them with the cv2 spelling)."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "gt_masks.npz")
    masks = (np.unpackbits(g["masks"], axis=-1).reshape(-1, 256, 256) * 255).astype(np.uint8)
    assert len(masks) == 631 and int(g["found"].sum()) == 631
    for i, m in enumerate(masks):
        q = og.find_quadrangle_cv2(m)
        assert q is not None and np.array_equal(q.reshape(4, 2), g["quads"][i]), i
        if i % 16 == 0:
            o = og.find_quadrangle(m)
            assert o is not None and np.array_equal(o.reshape(4, 2), g["quads"][i]), i
