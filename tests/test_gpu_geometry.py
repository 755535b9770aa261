"""Bit-exact parity of the integer/geometric kernels (csrc/geometry.cu, csrc/stem.cu) against the oracle
(oracle/geometry.py, itself pinned to cv2 and to the reference's golden vectors) on the same seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import geometry as og
import cvb_synth as synth

pytestmark = pytest.mark.gpu


def test_resize_area_half(engine):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (3, 512, 512, 3), dtype=np.uint8)
    out = engine.resize_area_half(torch.from_numpy(img).cuda()).cpu().numpy()
    for i in range(3):
        assert np.array_equal(out[i], og.resize_area_half(img[i]))


def test_mask_from_logits(engine):
    rng = np.random.default_rng(1)
    logits = rng.normal(0, 2, (2, 256, 256)).astype(np.float32)
    logits[0, 0, :8] = [0.0, 1e-8, -1e-8, 1e-6, -1e-6, 30.0, -30.0, 0.5]
    for thr in (0.5, 0.3, 0.9):
        got = engine.mask_from_logits(torch.from_numpy(logits).cuda(), thr).cpu().numpy()
        want = np.stack([og.binary_mask(l, thr) for l in logits])
        # expf differs from the CPU libm by <= 2 ulp: only logits within 1e-6 of the decision boundary may differ
        diff = got != want
        if diff.any():
            edge = np.log(thr / (1 - thr))
            assert np.all(np.abs(logits[diff] - edge) < 1e-5)
        assert set(np.unique(got)) <= {0, 255}


def test_mask_to_quad_matches_oracle(engine):
    masks = synth.mask_suite(seed=7, n=64)
    quad, found, status = engine.mask_to_quad(torch.from_numpy(masks).cuda())
    quad, found, status = quad.cpu().numpy(), found.cpu().numpy(), status.cpu().numpy()
    n_found = 0
    for i, m in enumerate(masks):
        want = og.find_quadrangle(m)
        assert status[i] != 2, f"mask {i}: capacity overflow"
        if want is None:
            assert found[i] == 0, f"mask {i}: oracle found nothing, kernel returned {quad[i].tolist()}"
        else:
            n_found += 1
            assert found[i] == 1, f"mask {i}: kernel found nothing, oracle {want.reshape(4, 2).tolist()}"
            assert np.array_equal(quad[i], want.reshape(4, 2)), f"mask {i}: {quad[i].tolist()} != {want.reshape(4, 2).tolist()}"
    assert n_found >= 30


def _check_quads(eng, masks):
    quad, found, status = eng.mask_to_quad(torch.from_numpy(masks).cuda())
    quad, found, status = quad.cpu().numpy(), found.cpu().numpy(), status.cpu().numpy()
    n_found = 0
    for i, m in enumerate(masks):
        want = og.find_quadrangle(m)
        assert status[i] in (0, 1), f"mask {i}: status {status[i]}"
        assert bool(found[i]) == (want is not None), f"mask {i}: found flag differs from the oracle"
        if want is not None:
            n_found += 1
            assert np.array_equal(quad[i], want.reshape(4, 2)), f"mask {i}: {quad[i].tolist()} != {want.reshape(4, 2).tolist()}"
    return n_found


def test_mask_to_quad_compact_and_full_kernels_agree_with_oracle(monkeypatch):
    """The compact kernel (bit planes, 4 boards per SM) defers big holes / capacity overflows to the full-state kernel;
    both routes, and the full-state kernel alone (CVB_QUAD_FULL=1), must reproduce the oracle on a second, larger suite
    that includes frames (a hole that passes the area filter), nested shapes and very ragged masks."""
    from chessvision import _native
    masks = list(synth.mask_suite(seed=23, n=96))
    frame = np.full((256, 256), 255, np.uint8)
    frame[20:236, 24:230] = 0                       # big hole: only hole border and image-frame outer border
    ring = np.zeros((256, 256), np.uint8)
    ring[16:240, 16:240] = 255
    ring[40:216, 40:216] = 0
    ring[60:196, 60:196] = 255                      # component nested inside the hole of another
    rng = np.random.default_rng(4)
    noisy = (rng.random((256, 256)) < 0.5).astype(np.uint8) * 255   # thousands of tiny borders
    speck = synth.quad_mask(rng, specks=0)
    speck[::7, ::5] ^= 255                           # isolated pixels and pin holes everywhere
    masks = np.stack(masks + [frame, ring, noisy, speck])
    eng = _native.Engine(0, max_batch=8)
    try:
        assert _check_quads(eng, masks) >= 40
    finally:
        eng.close()
    monkeypatch.setenv("CVB_QUAD_FULL", "1")
    eng = _native.Engine(0, max_batch=8)
    try:
        assert _check_quads(eng, masks) >= 40
    finally:
        eng.close()


def test_warp_squares_bit_exact(engine):
    rng = np.random.default_rng(3)
    imgs, quads = zip(*[synth.board_image(rng) for _ in range(6)])
    imgs, quads = np.stack(imgs), np.stack(quads).astype(np.int32)
    found = np.array([1, 1, 1, 0, 1, 1], np.uint8)
    board = engine.warp_squares(torch.from_numpy(imgs).cuda(), torch.from_numpy(quads).cuda(), torch.from_numpy(found).cuda()).cpu().numpy()
    for i in range(len(imgs)):
        if not found[i]:
            assert not board[i].any()
            continue
        want = og.extract_board(imgs[i], og.scale_quadrangle(quads[i].reshape(4, 1, 2), (512, 512)))
        assert np.array_equal(board[i], want), f"board {i}: {(board[i] != want).sum()} bytes differ"
