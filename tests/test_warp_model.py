"""The arithmetic k_warp_board's float32 coordinate path relies on, checked on the CPU (no GPU, no product code): the
integer-level emulation of the kernel (profiles/probes/warp_kernel_emulation.py -- every float32 step rounded as the device
rounds it) must give OpenCV's fixed-point coordinates (oracle.geometry.warp_coords, pinned to cv2) for every pixel outside
the guard band, and must never address a tap outside the staged patch.  The byte-level parity of the kernel itself is the
GPU suite's job (tests/test_gpu_geometry.py)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "profiles" / "probes")]
import warp_kernel_emulation as emu  # noqa: E402
from oracle import geometry as og  # noqa: E402


def _minv(quad):
    dest = np.array(((0, 0), (512, 0), (512, 512), (0, 512)), np.float32)
    M = og.perspective_matrix(og.scale_quadrangle(np.asarray(quad, np.int32).reshape(4, 1, 2), (512, 512)).reshape(4, 2), dest)
    return og.invert3(np.asarray(M, np.float64))


def test_float32_segment_offsets_round_like_opencv_outside_the_guard_band():
    quads = [
        [[158, 77], [219, 120], [158, 180], [102, 136]],      # an ordinary board
        [[147, 42], [109, 45], [10, 227], [246, 233]],        # strong perspective: some tiles leave the float32 path
        [[255, 0], [0, 0], [0, 255], [255, 255]],             # the whole frame, one source pixel per destination pixel
        [[168, 162], [179, 181], [160, 193], [149, 174]],     # tiny board: exact ties, a quarter of the pixels in the band
        [[43, 255], [219, 207], [-37, 283], [225, 151]],      # W changes sign inside the board
    ]
    fast = 0
    for i, q in enumerate(quads):
        st = emu.emulate(_minv(q))                            # asserts that no tap leaves the patch
        assert st["wrong"] == 0, (q, st)
        fast += st["fast_px"]
        if i == 0:                                            # a generic map: 0.2 % of the pixels in the band
            assert st["fast_px"] == 512 * 512 and st["redo_px"] < 0.01 * st["fast_px"], st
        if i in (2, 3):                                       # dyadic scales put whole columns on exact ties: the literal path handles them
            assert st["redo_px"] > 0.05 * st["fast_px"], st
    assert fast > 0.6 * 5 * 512 * 512                         # the float32 path is the common one


def test_guard_band_covers_the_modelled_error():
    """Worst modelled error of the float32 value (reciprocal off by a whole ulp either way) stays below half the band."""
    import warp_fast_model as model
    rng = np.random.default_rng(3)
    worst = 0.0
    for i in range(6):
        q = model.random_quad(rng, extreme=(i % 3 == 2))
        minv = _minv(q)
        if not np.isfinite(minv).all():
            continue
        for sign in (-1.0, 1.0):
            for err, _flagged, wrong, _skipped in model.model(minv, rng, sign):
                assert wrong == 0
                worst = max(worst, err)
    assert worst < model.BAND / 2, worst
