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


def _run_quads(eng, masks):
    quad, found, status = eng.mask_to_quad(torch.from_numpy(np.ascontiguousarray(masks)).cuda())
    return quad.cpu().numpy(), found.cpu().numpy(), status.cpu().numpy()


def test_mask_to_quad_fuzz_10k_against_cv2_and_the_oracle(engine):
    """10,240 seeded masks (blur + noise, holes, specks, frames, several blobs, salt-and-pepper fields, combs) through the
    three-kernel mask->quad path; every result must equal ChessVision._find_quadrangle spelled with the reference's own cv2
    calls (oracle.geometry.find_quadrangle_cv2), and every 64th the cv2-free oracle too.  No status other than none / found."""
    n_found = n_total = 0
    for part in range(10):
        masks = synth.fuzz_masks(seed=1000 + part, n=1024)
        quad, found, status = _run_quads(engine, masks)
        assert set(np.unique(status)) <= {0, 1}, f"part {part}: status values {np.unique(status)}"
        for i, m in enumerate(masks):
            want = og.find_quadrangle_cv2(m)
            assert bool(found[i]) == (want is not None), f"part {part} mask {i}: found flag differs from cv2"
            if want is not None:
                n_found += 1
                assert np.array_equal(quad[i], want.reshape(4, 2)), f"part {part} mask {i}: {quad[i].tolist()} != {want.reshape(4, 2).tolist()}"
            if i % 64 == 0:
                o = og.find_quadrangle(m)
                assert (o is None) == (want is None) and (o is None or np.array_equal(o, want)), f"part {part} mask {i}: oracle != cv2"
        n_total += len(masks)
    assert n_total >= 10000 and n_found >= 0.4 * n_total, (n_found, n_total)


def test_mask_to_quad_reference_ground_truth_masks_golden(engine):
    """The reference's 631 ground-truth board masks (data/board_extraction/masks) against the quadrangles its UNMODIFIED
    _find_quadrangle returned for them (tests/golden/gt_masks.npz, oracle/make_golden_masks.py)."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "gt_masks.npz")
    masks = (np.unpackbits(g["masks"], axis=-1).reshape(-1, 256, 256) * 255).astype(np.uint8)
    assert len(masks) == 631
    quad, found, status = _run_quads(engine, masks)
    assert np.array_equal(found, g["found"]) and set(np.unique(status)) <= {0, 1}
    assert np.array_equal(quad, g["quads"]), f"{int((quad != g['quads']).any(axis=(1, 2)).sum())} masks differ"


def test_mask_to_quad_contours_beyond_the_shared_memory_capacity(engine):
    """Combs (one contour of ~40,000 border points), a dense checker field (tens of thousands of borders) and a board with a
    ragged fringe: the shared-memory kernels flag them, the large-capacity kernel finishes them -- results equal to the
    oracle and to cv2, never status 2."""
    rng = np.random.default_rng(11)
    masks = [synth.comb_mask(rng) for _ in range(5)] + [synth.comb_mask(rng, teeth=2).T.copy() for _ in range(3)]
    checker = np.zeros((256, 256), np.uint8)
    checker[::2, ::2] = 255                                     # 16,384 isolated pixels: more borders than int16 labels hold
    checker[1::2, 1::2] = 255
    fringe = np.zeros((256, 256), np.uint8)
    fringe[30:226, 30:150] = 255
    fringe[30:226:2, 150:236] = 255                             # 98 one-pixel teeth of 86 px on a board-sized block
    masks = np.stack(masks + [checker, fringe, 255 - fringe])
    quad, found, status = _run_quads(engine, masks)
    for i, m in enumerate(masks):
        want = og.find_quadrangle_cv2(m)
        o = og.find_quadrangle(m)
        assert (o is None) == (want is None) and (o is None or np.array_equal(o, want)), f"mask {i}: oracle != cv2"
        assert status[i] in (0, 1), f"mask {i}: status {status[i]}"
        assert bool(found[i]) == (want is not None), f"mask {i}"
        if want is not None:
            assert np.array_equal(quad[i], want.reshape(4, 2)), f"mask {i}"


def test_warp_squares_with_corners_outside_the_image(engine):
    """Quads that leave the image (BORDER_CONSTANT: taps outside contribute 0), including corners far outside and a sliver
    (nearly collinear corners): bytes identical to the oracle."""
    rng = np.random.default_rng(5)
    imgs = np.stack([synth.board_image(rng)[0] for _ in range(6)])
    quads = np.array([
        [[270, -20], [-15, 10], [5, 250], [260, 290]],          # every corner outside
        [[255, 0], [0, 0], [0, 255], [255, 255]],               # the whole frame
        [[300, 40], [100, 30], [90, 200], [310, 220]],          # right half outside
        [[200, -60], [40, -50], [50, 120], [190, 130]],         # top outside
        [[1000, -800], [-900, -700], [-1000, 900], [1100, 1000]],   # far outside: mostly border colour
        [[12, 10], [100, 101], [203, 200], [250, 252]],         # a sliver: ill-conditioned system, identical arithmetic all the same
    ], np.int32)
    found = np.ones(6, np.uint8)
    board = engine.warp_squares(torch.from_numpy(imgs).cuda(), torch.from_numpy(quads).cuda(), torch.from_numpy(found).cuda()).cpu().numpy()
    for i in range(6):
        want = og.extract_board(imgs[i], og.scale_quadrangle(quads[i].reshape(4, 1, 2), (512, 512)))
        assert np.array_equal(board[i], want), f"board {i}: {(board[i] != want).sum()} bytes differ"


def _fuzz_quads(rng, n):
    """Board quads in the 256-px mask frame: ordinary boards, strong perspective, tiny and over-sized boards, corners outside
    the image, near-degenerate slivers -- every regime of k_warp_board (float32 offsets, per-thread and per-tile exact paths)."""
    out = []
    for i in range(n):
        kind = i % 8
        if kind <= 2:                                            # ordinary: a rotated, mildly sheared board
            c, r = rng.uniform(90, 166, 2), rng.uniform(50, 125)
            a = rng.uniform(0, 2 * np.pi) + np.array([0, 0.5, 1.0, 1.5]) * np.pi + rng.uniform(-0.2, 0.2, 4)
            q = c + (r * rng.uniform(0.8, 1.2, 4))[:, None] * np.stack([np.cos(a), np.sin(a)], 1)
        elif kind == 3:                                          # strong perspective: one edge a quarter of the opposite one
            w0, w1 = rng.uniform(20, 60), rng.uniform(160, 250)
            y0, y1 = rng.uniform(10, 80), rng.uniform(170, 250)
            q = np.array([[128 + w0 / 2, y0], [128 - w0 / 2, y0], [128 - w1 / 2, y1], [128 + w1 / 2, y1]]) + rng.uniform(-4, 4, (4, 2))
        elif kind == 4:                                          # tiny board: many destination pixels per source pixel
            c, r = rng.uniform(60, 200, 2), rng.uniform(6, 25)
            a = rng.uniform(0, 2 * np.pi) + np.array([0, 0.5, 1.0, 1.5]) * np.pi
            q = c + r * np.stack([np.cos(a), np.sin(a)], 1)
        elif kind == 5:                                          # larger than the frame: taps outside, footprints beyond the staged patch
            c, r = rng.uniform(100, 156, 2), rng.uniform(200, 700)
            a = rng.uniform(0, 2 * np.pi) + np.array([0, 0.5, 1.0, 1.5]) * np.pi + rng.uniform(-0.3, 0.3, 4)
            q = c + r * np.stack([np.cos(a), np.sin(a)], 1)
        elif kind == 6:                                          # any four points (self-intersecting, W changing sign)
            q = rng.uniform(-40, 296, (4, 2))
        else:                                                    # sliver
            t = np.sort(rng.uniform(0, 255, 4))
            q = np.stack([t, t + rng.uniform(-3, 3, 4)], 1)
        out.append(np.round(q).astype(np.int32))
    return np.stack(out)


def test_warp_squares_fuzz_noise_images_against_cv2(engine):
    """256 quads x white-noise images (every 1/32-px coordinate difference changes bytes): warp + gray + mirror must equal
    cv2.getPerspectiveTransform + cv2.warpPerspective + cvtColor + flip byte for byte (utils.py:115-132, core.py:298-300)."""
    import cv2
    rng = np.random.default_rng(20261018)
    quads = _fuzz_quads(rng, 256)
    dest = np.array(((0, 0), (512, 0), (512, 512), (0, 512)), np.float32)
    imgs = rng.integers(0, 256, (8, 512, 512, 3), dtype=np.uint8)
    dimgs = torch.from_numpy(imgs).cuda()
    found = torch.ones(8, dtype=torch.uint8).cuda()
    n_checked = 0
    for lo in range(0, 256, 8):
        q = quads[lo:lo + 8]
        board = engine.warp_squares(dimgs, torch.from_numpy(q).cuda(), found).cpu().numpy()
        for i in range(8):
            src = og.scale_quadrangle(q[i].reshape(4, 1, 2), (512, 512)).reshape(4, 2)
            try:
                M = cv2.getPerspectiveTransform(src, dest)
            except cv2.error:
                continue
            if not np.isfinite(M).all() or abs(np.linalg.det(M)) < 1e-12:
                continue                                        # singular system: cv2 returns garbage of its own LU; not a board
            want = cv2.flip(cv2.cvtColor(cv2.warpPerspective(imgs[i], M, (512, 512)), cv2.COLOR_BGR2GRAY), 1)
            assert np.array_equal(board[i], want), f"quad {lo + i} {q[i].tolist()}: {(board[i] != want).sum()} bytes differ"
            n_checked += 1
    assert n_checked >= 200, n_checked


@pytest.mark.parametrize("shape,out_size", [((300, 400, 3), (512, 512)), ((300, 400, 3), (300, 200)), ((480, 640, 3), (640, 480)),
                                            ((97, 131, 3), (50, 30)), ((256, 256), (64, 64)), ((200, 100), (1000, 8)),
                                            ((64, 64, 1), (33, 77)), ((512, 512, 3), (5, 5))])
def test_warp_perspective_any_size_equals_cv2(engine, shape, out_size):
    """cvb_warp_perspective (utils.extract_perspective, utils.py:115-132) against cv2.getPerspectiveTransform +
    cv2.warpPerspective themselves: float32 corners (partly outside the image), 1 and 3 channels, any output size."""
    import cv2
    rng = np.random.default_rng(sum(shape) + out_size[0])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    h, w = shape[:2]
    for _ in range(4):
        quad = (np.array([[0.1, 0.1], [0.9, 0.15], [0.85, 0.9], [0.12, 0.8]]) * [w, h] + rng.uniform(-0.2, 0.2, (4, 2)) * [w, h]).astype(np.float32)
        dest = np.array(((0, 0), (out_size[0], 0), (out_size[0], out_size[1]), (0, out_size[1])), np.float32)
        want = cv2.warpPerspective(img, cv2.getPerspectiveTransform(quad, dest), out_size)
        got = engine.warp_perspective(torch.from_numpy(img).cuda(), torch.from_numpy(quad), out_size).cpu().numpy()
        assert np.array_equal(got.reshape(want.shape), want), f"{(got.reshape(want.shape) != want).sum()} bytes differ"
        assert np.array_equal(og.warp_perspective_u8(img.reshape(h, w, -1) if img.ndim == 3 else img, og.perspective_matrix(quad, dest), out_size).reshape(want.shape), want)
