"""ORACLE tooling: freeze the reference's own ground-truth board masks and what its mask->quad step returns for them.

Runs only in the build container (needs /root/reference).  Every mask under ``data/board_extraction/masks`` (631 PNGs,
256x256, {0,255}) goes through the UNMODIFIED ``ChessVision._find_quadrangle`` (chessvision/core.py:358-379, real cv2);
masks are stored bit-packed, quads as int32[4,2] (zeros + found = 0 where the reference returns None).

    python oracle/make_golden_masks.py       # writes tests/golden/gt_masks.npz
"""
from __future__ import annotations

import glob
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import geometry as og  # noqa: E402
from oracle.ref_loader import REF, load_reference  # noqa: E402


def main():
    ref = load_reference()
    files = sorted(glob.glob(os.path.join(REF, "data/board_extraction/masks/*.png")))
    assert len(files) >= 600, len(files)
    masks, quads, found = [], [], []
    for f in files:
        m = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
        assert m.shape == (256, 256) and set(np.unique(m)) <= {0, 255}, f
        q = ref.ChessVision._find_quadrangle(m)
        o = og.find_quadrangle(m)                      # the cv2-free restatement, pinned here on all 631 masks
        assert (q is None) == (o is None) and (q is None or np.array_equal(q, o)), f
        c = og.find_quadrangle_cv2(m)                  # and the cv2 restatement the GPU fuzz test checks against
        assert (q is None) == (c is None) and (q is None or np.array_equal(q, c)), f
        masks.append(np.packbits(m > 0))
        found.append(q is not None)
        quads.append(np.zeros((4, 2), np.int32) if q is None else q.reshape(4, 2).astype(np.int32))
    out = os.path.join(ROOT, "tests", "golden", "gt_masks.npz")
    np.savez_compressed(out, masks=np.stack(masks), quads=np.stack(quads), found=np.array(found, np.uint8),
                        files=np.array([os.path.basename(f) for f in files]))
    print(f"{len(files)} masks, {int(np.sum(found))} quadrangles -> {out} ({os.path.getsize(out) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()
