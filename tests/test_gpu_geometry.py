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
