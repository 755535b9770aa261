"""ORACLE (test infrastructure, not product code): cv2-free numpy restatement of the integer / geometric half of the
image->FEN path.

The arithmetic restated here lives in OpenCV, an un-vendored third-party dependency of the reference
(``opencv-python==4.11.0.86`` in the reference's uv.lock; this image has opencv-python-headless 4.13.0).  Each function
names the reference call site it stands in for.  Pinning: ``tests/test_oracle_geometry.py`` compares every function with
the live ``cv2`` of this image on real and synthetic inputs, and ``tests/golden/`` holds outputs of the unmodified
reference (``oracle/make_golden.py``).  Only tests, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs import it.
"""
from __future__ import annotations

import struct

import numpy as np

# ----------------------------------------------------------------------------------------------------------------------
# K0  cv2.resize(img, (256,256), INTER_AREA) for an exact 2x reduction        (core.py:212)
# ----------------------------------------------------------------------------------------------------------------------


def resize_area_half(img: np.ndarray) -> np.ndarray:
    """u8[2h,2w,c] -> u8[h,w,c];  (a+b+c+d+2)>>2 per channel."""
    assert img.dtype == np.uint8 and img.shape[0] % 2 == 0 and img.shape[1] % 2 == 0
    v = img.astype(np.uint16)
    s = v[0::2, 0::2] + v[0::2, 1::2] + v[1::2, 0::2] + v[1::2, 1::2] + 2
    return (s >> 2).astype(np.uint8)


def _area_tab(ssize: int, dsize: int, scale: float):
    """OpenCV's computeResizeAreaTab (imgproc/resize.cpp): per destination index the (source index, float32 weight) pairs
    of the source cells it covers — a partial cell on the left, whole cells, a partial cell on the right."""
    import math
    out = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = math.ceil(f1), math.floor(f2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        ent = []
        if s1 - f1 > 1e-3:
            ent.append((s1 - 1, np.float32((s1 - f1) / cell)))
        ent.extend((sx, np.float32(1.0 / cell)) for sx in range(s1, s2))
        if f2 - s2 > 1e-3:
            ent.append((s2, np.float32(min(min(f2 - s2, 1.0), cell) / cell)))
        out.append(ent)
    return out


def _linear_area_tab(ssize: int, dsize: int):
    """Coefficient loop of OpenCV's bilinear resizer in "area mode" (imgproc/resize.cpp, cv::resize with INTER_AREA when an
    axis is enlarged): per destination index the left source index and two weights in 1/2048 (saturate_cast<short> of
    float32 products), and the first destination index whose right tap is outside the row."""
    import math
    inv = dsize / ssize
    scale = 1.0 / inv
    ofs = np.zeros(dsize, np.int64)
    wgt = np.zeros((dsize, 2), np.int64)
    dmax = dsize
    for d in range(dsize):
        s = math.floor(d * scale)
        f = np.float32((d + 1) - (s + 1) * inv)
        f = np.float32(0.0) if f <= 0 else np.float32(f - np.float32(math.floor(f)))
        if s < 0:
            f, s = np.float32(0), 0
        if s + 1 >= ssize:
            dmax = min(dmax, d)
            if s >= ssize - 1:
                f, s = np.float32(0), ssize - 1
        ofs[d] = s
        wgt[d, 0] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(2048)))
        wgt[d, 1] = int(np.rint(np.float32(f) * np.float32(2048)))
    return ofs, wgt, dmax


def resize_area_enlarge(img: np.ndarray, dsize=(256, 256)) -> np.ndarray:
    """``cv2.resize(img, dsize, interpolation=cv2.INTER_AREA)`` when at least one axis is enlarged.  OpenCV implements true
    area interpolation for reductions only ("In other cases it is emulated using some variant of bilinear", resize.cpp):
    both axes then go through the 8-bit fixed-point bilinear resizer with area-mode weights,
    ``h = S[sx]*a0 + S[sx+1]*a1`` per source row and ``(((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2`` per pixel
    [pinned bit for bit against live cv2 in tests/test_oracle_geometry.py]."""
    dw, dh = dsize
    sh, sw, cn = img.shape
    xo, xa, xmax = _linear_area_tab(sw, dw)
    yo, ya, _ = _linear_area_tab(sh, dh)
    src = img.astype(np.int64)
    x1 = np.minimum(xo + 1, sw - 1)
    hbuf = src[:, xo] * xa[None, :, 0, None] + src[:, x1] * xa[None, :, 1, None]
    edge = np.arange(dw) >= xmax
    hbuf[:, edge] = src[:, xo[edge]] * 2048
    s0, s1 = hbuf[np.clip(yo, 0, sh - 1)], hbuf[np.clip(yo + 1, 0, sh - 1)]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_area(img: np.ndarray, dsize=(256, 256)) -> np.ndarray:
    """``cv2.resize(img, dsize, interpolation=cv2.INTER_AREA)`` (core.py:212), u8[H,W,C].  An enlarged axis sends the call to
    :func:`resize_area_enlarge`; a reduction in both axes (H >= dsize[1], W >= dsize[0]) has two code paths in OpenCV,
    both restated bit for bit [pinned against live cv2]:
    integer scale factors average whole cells in integers ((a+b+c+d+2)>>2 for 2x2, round-half-even of sum * float32(1/area)
    otherwise); everything else accumulates float32 products row by row — horizontally in table order into a row buffer,
    then ``sum = beta*buf`` for the first source row of a destination row and ``sum += beta*buf`` after it — and rounds
    half to even at the end.  No fused multiply-add anywhere."""
    dw, dh = dsize
    sh, sw, cn = img.shape
    assert img.dtype == np.uint8
    if sh < dh or sw < dw:
        return resize_area_enlarge(img, dsize)
    fx, fy = sw / dw, sh / dh
    ix, iy = int(round(fx)), int(round(fy))
    if abs(fx - ix) < 2.220446049250313e-16 and abs(fy - iy) < 2.220446049250313e-16:
        s = img[: dh * iy, : dw * ix].reshape(dh, iy, dw, ix, cn).astype(np.int64).sum((1, 3))
        if ix == 2 and iy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        v = (s.astype(np.float32) * np.float32(1.0 / (ix * iy))).astype(np.float32)
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    xt, yt = _area_tab(sw, dw, fx), _area_tab(sh, dh, fy)
    src = img.astype(np.float32)
    buf = np.zeros((sh, dw, cn), np.float32)
    for dx, ent in enumerate(xt):
        for si, a in ent:
            buf[:, dx] = (buf[:, dx] + (src[:, si] * a).astype(np.float32)).astype(np.float32)
    out = np.zeros((dh, dw, cn), np.float32)
    for dy, ent in enumerate(yt):
        for k, (si, b) in enumerate(ent):
            t = (b * buf[si]).astype(np.float32)
            out[dy] = t if k == 0 else (out[dy] + t).astype(np.float32)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------------------
# K5  sigmoid + threshold                                                      (core.py:273, utils.py:101-112)
# ----------------------------------------------------------------------------------------------------------------------


def binary_mask(logits: np.ndarray, threshold: float = 0.5) -> np.ndarray:
    """f32 logits -> u8 {0,255}.  The reference evaluates ``torch.sigmoid`` in fp32 (core.py:273) and compares
    ``> threshold`` (utils.py:109-112); the same library call is used here because ATen's vectorised fp32 sigmoid is not
    bit-identical to a naive ``1/(1+exp(-x))`` for 0 < x < 2e-7 (it yields > 0.5 from x = 8.94e-8 on)."""
    import torch
    assert logits.dtype == np.float32 and 0 <= threshold <= 1
    p = torch.sigmoid(torch.from_numpy(np.ascontiguousarray(logits))).numpy()
    return np.where(p > threshold, 255, 0).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------------------
# K6  cv2.findContours(mask, RETR_CCOMP, CHAIN_APPROX_TC89_KCOS)               (core.py:360)
# ----------------------------------------------------------------------------------------------------------------------

# Freeman codes, y pointing down
_DX = (1, 1, 0, -1, -1, -1, 0, 1)
_DY = (0, -1, -1, -1, 0, 1, 1, 1)


def trace_borders(mask: np.ndarray):
    """Suzuki-Abe border following on a binary image.

    Returns a list of ``(points[n,2] int32 (x,y), codes[n] int, is_hole, parent_outer_index)`` in *discovery* order.
    ``codes[i]`` is the Freeman step from point i to point i+1 (cyclic); a single-pixel border has n=1 and no steps.
    """
    h, w = mask.shape
    lab = np.zeros((h + 2, w + 2), np.int32)
    lab[1:-1, 1:-1] = mask != 0
    nz = lab != 0
    borders = []
    owner = {}  # border label -> index (into `borders`) of the outer border of its component
    nbd = 1
    for y in range(1, h + 1):
        row_nz = nz[y]
        # candidate columns: zero/non-zero changes between x-1 and x
        cand = np.nonzero(row_nz[1:] != row_nz[:-1])[0] + 1
        last_label = 0  # label of the most recent border pixel met in this row ("lnbd")
        prev_c = 0
        for x in cand.tolist():
            # labelled pixels passed since the previous candidate keep `last_label` current
            seg = lab[y, prev_c:x]
            marked = np.nonzero((seg != 0) & (seg != 1))[0]
            if marked.size:
                last_label = abs(int(seg[marked[-1]]))
            prev_c = x
            p, prev = int(lab[y, x]), int(lab[y, x - 1])
            if prev == 0 and p == 1:
                hole, sx = False, x
            elif p == 0 and prev >= 1:
                hole, sx = True, x - 1
                if prev > 1:
                    last_label = prev
            else:
                continue
            nbd += 1
            pts, codes = _follow(lab, sx, y, nbd, hole)
            if hole:
                parent = owner.get(last_label, -1)
            else:
                parent = len(borders)
            owner[nbd] = parent
            borders.append((pts, codes, hole, parent))
    return borders


def _follow(lab, x0, y0, nbd, hole):
    s_end = s = 0 if hole else 4
    while True:
        s = (s - 1) & 7
        if lab[y0 + _DY[s], x0 + _DX[s]] != 0 or s == s_end:
            break
    if s == s_end:  # isolated pixel
        lab[y0, x0] = -nbd
        return np.array([[x0 - 1, y0 - 1]], np.int32), np.zeros(0, np.int32)
    x1, y1 = x0 + _DX[s], y0 + _DY[s]
    x3, y3 = x0, y0
    pts, codes = [], []
    while True:
        s_end = s
        while True:
            s += 1
            x4, y4 = x3 + _DX[s & 7], y3 + _DY[s & 7]
            if lab[y4, x4] != 0:
                break
        s &= 7
        if ((s - 1) & 0xFFFFFFFF) < s_end:
            lab[y3, x3] = -nbd
        elif lab[y3, x3] == 1:
            lab[y3, x3] = nbd
        pts.append((x3 - 1, y3 - 1))
        codes.append(s)
        if x4 == x0 and y4 == y0 and x3 == x1 and y3 == y1:
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return np.array(pts, np.int32), np.array(codes, np.int32)


def order_ccomp(borders):
    """RETR_CCOMP output order: outer borders in reverse discovery order, each followed by its holes (reverse order)."""
    out = []
    for i in range(len(borders) - 1, -1, -1):
        if borders[i][2]:
            continue
        out.append(i)
        for j in range(len(borders) - 1, i, -1):
            if borders[j][2] and borders[j][3] == i:
                out.append(j)
    return out


_T = (1, 2, 3, 4, 3, 2, 1, 0, 1, 2, 3, 4, 3, 2, 1)


def _f32(v: float) -> float:
    return struct.unpack("f", struct.pack("f", v))[0]


def _f32_bits(v: float) -> int:
    return struct.unpack("i", struct.pack("f", v))[0]


def kcos_reduce(pts: np.ndarray, codes: np.ndarray) -> np.ndarray:
    """CHAIN_APPROX_TC89_KCOS vertex reduction of one closed border (SURVEY Appendix A.4 ii)."""
    n = len(pts)
    if n == 1:
        return pts.copy()
    P = pts.tolist()
    c = codes.tolist()
    s = [0] * n
    ksup = [0] * n
    for i in range(n):
        s[i] = _T[c[i] - c[i - 1] + 7]
    cand = [i for i in range(n) if s[i] != 0]
    # pass 1: region of support + k-cosine
    for i in cand:
        xi, yi = P[i]
        d_num = 0
        l = 0
        k = 1
        while True:
            x1, y1 = P[(i - k) % n]
            x2, y2 = P[(i + k) % n]
            dx, dy = x2 - x1, y2 - y1
            lk = dx * dx + dy * dy
            dk = (xi - x1) * dy - (yi - y1) * dx
            t = _f32(float(d_num) * float(lk) - float(dk) * float(l))
            if k > 1 and (l >= lk or (d_num > 0 and t <= 0) or (d_num < 0 and t >= 0)):
                break
            d_num, l = dk, lk
            k += 1
        k -= 1
        ksup[i] = k
        sv = 0
        j = k
        while j > 0:
            ax, ay = P[(i - j) % n][0] - xi, P[(i - j) % n][1] - yi
            bx, by = P[(i + j) % n][0] - xi, P[(i + j) % n][1] - yi
            if (ax == 0 and ay == 0) or (bx == 0 and by == 0):
                break
            num = float(ax * bx + ay * by)
            cs = _f32(num / float(np.sqrt(np.float64(float(ax * ax + ay * ay) * float(bx * bx + by * by)))))
            sk = _f32_bits(_f32(cs + 1.1))
            if j < k and sk <= sv:
                break
            sv = sk
            j -= 1
        s[i] = sv
    # pass 2: non-maximum suppression (sequential, sees earlier zeroing)
    alive = []
    for i in cand:
        k2 = ksup[i] >> 1
        keep = True
        for j in range(1, k2 + 1):
            if s[(i - j) % n] > s[i] or s[(i + j) % n] > s[i]:
                keep = False
                break
        if keep:
            alive.append(i)
        else:
            s[i] = 0
    # pass 3: k == 1 survivors must be strict local maxima
    out = []
    for i in alive:
        if ksup[i] == 1 and (s[i] <= s[(i - 1) % n] or s[i] <= s[(i + 1) % n]):
            s[i] = 0
            continue
        out.append(i)
    return pts[out] if out else pts[:0]


def find_contours_ccomp_kcos(mask: np.ndarray):
    b = trace_borders(mask)
    return [kcos_reduce(b[i][0], b[i][1]) for i in order_ccomp(b)]


# ----------------------------------------------------------------------------------------------------------------------
# contourArea / boundingRect / arcLength / approxPolyDP                       (core.py:373-374, 394, 398)
# ----------------------------------------------------------------------------------------------------------------------


def contour_area(c: np.ndarray) -> float:
    x = c[:, 0].astype(np.int64)
    y = c[:, 1].astype(np.int64)
    a = int(np.sum(np.roll(x, 1) * y - np.roll(y, 1) * x))
    return abs(a * 0.5)


def bounding_rect(c: np.ndarray):
    x0, y0 = int(c[:, 0].min()), int(c[:, 1].min())
    return x0, y0, int(c[:, 0].max()) - x0 + 1, int(c[:, 1].max()) - y0 + 1


def arc_length_closed(c: np.ndarray) -> float:
    """Per-segment length in float32, accumulated in float64 in blocks of 16 added back-to-front (cv::arcLength)."""
    n = len(c)
    if n <= 1:
        return 0.0
    p = c.astype(np.float32)
    d = p - np.roll(p, 1, axis=0)
    seg = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
    total = 0.0
    for b0 in range(0, n, 16):
        blk = seg[b0:b0 + 16]
        for v in blk[::-1]:
            total += float(v)
    return total


def approx_poly_closed(c: np.ndarray, epsilon: float) -> np.ndarray:
    """cv::approxPolyDP(curve, epsilon, closed=True) for integer points (SURVEY Appendix A.4 iv)."""
    src = c.tolist()
    n = len(src)
    if n == 0:
        return c[:0]
    eps = epsilon * epsilon
    dst = []
    stack = []
    pos = 0
    rstart = 0
    le_eps = False
    for _ in range(3):
        max_dist = 0.0
        pos = (pos + rstart) % n
        sx, sy = src[pos]
        pos = (pos + 1) % n
        for j in range(1, n):
            px, py = src[pos]
            pos = (pos + 1) % n
            dx, dy = px - sx, py - sy
            dist = float(dx * dx + dy * dy)
            if dist > max_dist:
                max_dist = dist
                rstart = j
        le_eps = max_dist <= eps
    if not le_eps:
        s_start = pos % n
        s_end = (rstart + s_start) % n
        stack.append((s_end, s_start))  # right slice
        stack.append((s_start, s_end))  # slice, popped first
    else:
        dst.append((sx, sy))
    while stack:
        a, b = stack.pop()
        ex, ey = src[b]
        pos = a
        sx, sy = src[pos]
        pos = (pos + 1) % n
        if pos != b:
            dx, dy = float(ex - sx), float(ey - sy)
            max_dist = 0.0
            split = 0
            while pos != b:
                px, py = src[pos]
                pos = (pos + 1) % n
                dist = abs((py - sy) * dx - (px - sx) * dy)
                if dist > max_dist:
                    max_dist = dist
                    split = (pos + n - 1) % n
            le = max_dist * max_dist <= eps * (dx * dx + dy * dy)
        else:
            le = True
        if le:
            dst.append((sx, sy))
        else:
            stack.append((split, b))
            stack.append((a, split))
    # clean-up pass (in place, as in OpenCV)
    count = len(dst)
    new_count = count
    pos = count - 1
    start = dst[pos]
    pos = 0 if pos + 1 >= count else pos + 1
    wpos = pos
    pt = dst[pos]
    pos = 0 if pos + 1 >= count else pos + 1
    i = 0
    while i < count and new_count > 2:
        end = dst[pos]
        pos = 0 if pos + 1 >= count else pos + 1
        dx, dy = float(end[0] - start[0]), float(end[1] - start[1])
        dist = abs((pt[0] - start[0]) * dy - (pt[1] - start[1]) * dx)
        sip = (pt[0] - start[0]) * (end[0] - pt[0]) + (pt[1] - start[1]) * (end[1] - pt[1])
        if dist * dist <= 0.5 * eps * (dx * dx + dy * dy) and dx != 0 and dy != 0 and sip >= 0:
            new_count -= 1
            dst[wpos] = start = end
            wpos = 0 if wpos + 1 >= count else wpos + 1
            pt = dst[pos]
            pos = 0 if pos + 1 >= count else pos + 1
            i += 2
            continue
        dst[wpos] = start = pt
        wpos = 0 if wpos + 1 >= count else wpos + 1
        pt = end
        i += 1
    return np.array(dst[:new_count], np.int32).reshape(-1, 2)


# ----------------------------------------------------------------------------------------------------------------------
# ChessVision._find_quadrangle / _filter_contours / _rotate_quadrangle / _scale_quadrangle   (core.py:358-417)
# ----------------------------------------------------------------------------------------------------------------------


def _ratio(a, b):  # utils.py:89-93
    if a == 0 or b == 0:
        return -1
    return min(a, b) / float(max(a, b))


def find_quadrangle(mask: np.ndarray):
    """u8[256,256] {0,255} -> int32[4,1,2] (x,y) or None; mirrors core.py:358-379 step by step."""
    contours = find_contours_ccomp_kcos(mask)
    if len(contours) > 1:  # the filter is skipped for a single contour (core.py:362)
        area_all = float(mask.shape[0] * mask.shape[1])
        kept = []
        for c in contours:
            a = contour_area(c) / area_all
            if a < 0.35 or a > 1.0:
                continue
            _, _, w, h = bounding_rect(c)
            if _ratio(h, w) < 0.6:
                continue
            kept.append(c)
        contours = kept
    for c in contours:
        cand = approx_poly_closed(c, 0.1 * arc_length_closed(c))
        if len(cand) == 4:
            q = cand.reshape(4, 1, 2)
            if q[0, 0, 0] < q[2, 0, 0]:  # core.py:407-411
                q = q[[3, 0, 1, 2]]
            return q.astype(np.int32)
    return None


def find_quadrangle_cv2(mask: np.ndarray):
    """``ChessVision._find_quadrangle`` (core.py:358-411) spelled with the same cv2 calls the reference makes -- the fast
    checker for large fuzz suites (``find_quadrangle`` above is the cv2-free restatement and is pinned against this one and
    against the unmodified reference in tests/test_oracle_geometry.py / oracle/make_golden_masks.py)."""
    import cv2
    contours, _ = cv2.findContours(mask, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_TC89_KCOS)
    if len(contours) > 1:
        area = float(mask.shape[0] * mask.shape[1])
        kept = []
        for c in contours:
            a = cv2.contourArea(c) / area
            if a < 0.35 or a > 1.0:
                continue
            _, _, w, h = cv2.boundingRect(c)
            if _ratio(h, w) < 0.6:
                continue
            kept.append(c)
        contours = kept
    for c in contours:
        cand = cv2.approxPolyDP(c, 0.1 * cv2.arcLength(c, True), True)
        if len(cand) == 4:
            return cand[[3, 0, 1, 2], :, :] if cand[0, 0, 0] < cand[2, 0, 0] else cand
    return None


def scale_quadrangle(q: np.ndarray, orig_hw) -> np.ndarray:
    return np.array(q * (orig_hw[0] / 256.0), dtype=np.float32)  # height for both axes (core.py:414-417)


# ----------------------------------------------------------------------------------------------------------------------
# K7  getPerspectiveTransform + warpPerspective + BGR2GRAY + flip              (utils.py:127-132, core.py:299-300)
# ----------------------------------------------------------------------------------------------------------------------


def perspective_matrix(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """cv2.getPerspectiveTransform: 8x8 system, LU with partial pivoting in float64."""
    src = np.asarray(src, np.float32).reshape(4, 2)
    dst = np.asarray(dst, np.float32).reshape(4, 2)
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        A[i, 0] = A[i + 4, 3] = src[i, 0]
        A[i, 1] = A[i + 4, 4] = src[i, 1]
        A[i, 2] = A[i + 4, 5] = 1.0
        # Point2f * Point2f: OpenCV forms these products in float32 before they enter the float64 system (exact whenever the
        # destination size is a power of two, as for the 512 x 512 board)
        A[i, 6] = np.float32(-src[i, 0] * dst[i, 0])
        A[i, 7] = np.float32(-src[i, 1] * dst[i, 0])
        A[i + 4, 6] = np.float32(-src[i, 0] * dst[i, 1])
        A[i + 4, 7] = np.float32(-src[i, 1] * dst[i, 1])
        b[i] = dst[i, 0]
        b[i + 4] = dst[i, 1]
    m = 8
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(A[j, i]) > abs(A[k, i]):
                k = j
        if abs(A[k, i]) < np.finfo(np.float64).eps * 100:
            return None
        if k != i:
            A[[i, k], i:] = A[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / A[i, i]
        for j in range(i + 1, m):
            alpha = A[j, i] * d
            for kk in range(i + 1, m):
                A[j, kk] += alpha * A[i, kk]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for kk in range(i + 1, m):
            s -= A[i, kk] * b[kk]
        b[i] = s / A[i, i]
    return np.append(b, 1.0).reshape(3, 3)


def invert3(M: np.ndarray) -> np.ndarray:
    """cv::invert for a 3x3 double matrix (adjugate times reciprocal determinant, OpenCV's operation order)."""
    a = M
    det = (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0])
           + a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))
    d = 1.0 / det
    t = np.empty(9)
    t[0] = (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) * d
    t[1] = (a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2]) * d
    t[2] = (a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]) * d
    t[3] = (a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2]) * d
    t[4] = (a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0]) * d
    t[5] = (a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]) * d
    t[6] = (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]) * d
    t[7] = (a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1]) * d
    t[8] = (a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]) * d
    return t.reshape(3, 3)


WARP_BLOCK_W = 64  # cv::WarpPerspectiveInvoker tiles the destination 64 wide x 16 high for a 512x512 output


def warp_block_width(out_w: int, out_h: int) -> int:
    """Width of the destination blocks of cv::WarpPerspectiveInvoker (imgproc/imgwarp.cpp, BLOCK_SZ = 32):
    bh0 = min(16, height); bw0 = min(1024 / bh0, width) -- 64 for every output at least 64 wide and 16 high."""
    bh0 = min(16, out_h)
    return min(1024 // bh0, out_w)


def warp_coords(Minv: np.ndarray, out_w: int, out_h: int):
    """Fixed-point source coordinates (1/32 px) for every destination pixel, OpenCV's evaluation order."""
    m = Minv.reshape(9)
    xs = np.arange(out_w)
    bw = warp_block_width(out_w, out_h)
    bx = (xs // bw) * bw
    x1 = (xs - bx).astype(np.float64)
    bx = bx.astype(np.float64)
    ys = np.arange(out_h, dtype=np.float64)[:, None]
    X0 = m[0] * bx[None, :] + m[1] * ys + m[2]
    Y0 = m[3] * bx[None, :] + m[4] * ys + m[5]
    W0 = m[6] * bx[None, :] + m[7] * ys + m[8]
    W = W0 + m[6] * x1[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        W = np.where(W != 0, 32.0 / W, 0.0)
    fX = np.clip((X0 + m[0] * x1[None, :]) * W, -2147483648.0, 2147483647.0)
    fY = np.clip((Y0 + m[3] * x1[None, :]) * W, -2147483648.0, 2147483647.0)
    return np.rint(fX).astype(np.int64), np.rint(fY).astype(np.int64)


def warp_perspective_u8(img: np.ndarray, M: np.ndarray, out_size) -> np.ndarray:
    """cv2.warpPerspective(img, M, out_size) with INTER_LINEAR / BORDER_CONSTANT(0) on u8[H,W,3]."""
    out_w, out_h = out_size
    H, Wd = img.shape[:2]
    Xi, Yi = warp_coords(invert3(np.asarray(M, np.float64)), out_w, out_h)
    # OpenCV stores the integer part as int16 (saturating)
    x0 = np.clip(Xi >> 5, -32768, 32767)
    y0 = np.clip(Yi >> 5, -32768, 32767)
    ax = (Xi & 31).astype(np.int64)
    ay = (Yi & 31).astype(np.int64)
    src = img.astype(np.int64)
    if src.ndim == 2:
        src = src[:, :, None]

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < Wd)
        v = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, Wd - 1)]
        return v * ok[:, :, None]

    w00 = ((32 - ax) * (32 - ay) * 32)[:, :, None]
    w01 = (ax * (32 - ay) * 32)[:, :, None]
    w10 = ((32 - ax) * ay * 32)[:, :, None]
    w11 = (ax * ay * 32)[:, :, None]
    acc = tap(y0, x0) * w00 + tap(y0, x0 + 1) * w01 + tap(y0 + 1, x0) * w10 + tap(y0 + 1, x0 + 1) * w11
    out = ((acc + 16384) >> 15).astype(np.uint8)
    return out if img.ndim == 3 else out[:, :, 0]


def bgr_to_gray(img: np.ndarray) -> np.ndarray:
    v = img.astype(np.int64)
    return ((3735 * v[:, :, 0] + 19235 * v[:, :, 1] + 9798 * v[:, :, 2] + 16384) >> 15).astype(np.uint8)


def extract_board(image: np.ndarray, scaled_quad: np.ndarray, out_size=(512, 512)) -> np.ndarray:
    """utils.extract_perspective + cvtColor + flip (utils.py:115-132, core.py:298-300) -> u8[512,512]."""
    w, h = out_size
    dest = np.array(((0, 0), (w, 0), (w, h), (0, h)), np.float32)
    M = perspective_matrix(np.asarray(scaled_quad, np.float32).reshape(4, 2), dest)
    board = warp_perspective_u8(image, M, out_size)
    return np.ascontiguousarray(bgr_to_gray(board)[:, ::-1])


def extract_squares(board: np.ndarray) -> np.ndarray:
    """core.py:420-439 -> u8[64,64,64,1], index = 8*row + col, row 0 = rank 8."""
    h, w = board.shape
    sh, sw = h // 8, w // 8
    return board.reshape(8, sh, 8, sw).transpose(0, 2, 1, 3).reshape(64, sh, sw, 1)


# ----------------------------------------------------------------------------------------------------------------------
# K9/K10  argmax, rule 1, FEN                                                  (core.py:310-355, 442-469)
# ----------------------------------------------------------------------------------------------------------------------

LABEL_NAMES = ["B", "K", "N", "P", "Q", "R", "b", "k", "n", "p", "q", "r", "f"]  # constants.py:23
SQUARE_NAMES_NORMAL = [f + r for r in "87654321" for f in "abcdefgh"]  # constants.py:109-118
SQUARE_NAMES_FLIPPED = SQUARE_NAMES_NORMAL[::-1]  # constants.py:120-129


def board_fen(labels, square_names) -> str:
    """python-chess ``BaseBoard.board_fen()`` for 64 labels ('f' = empty) placed on ``square_names``."""
    grid = {}
    for lab, sq in zip(labels, square_names):
        grid[sq] = None if lab == "f" else lab
    rows = []
    for r in "87654321":
        row, empty = "", 0
        for f in "abcdefgh":
            p = grid.get(f + r)
            if p is None:
                empty += 1
            else:
                row += (str(empty) if empty else "") + p
                empty = 0
        rows.append(row + (str(empty) if empty else ""))
    return "/".join(rows)


def validate_labels(labels, probs, square_names):
    """Rule 1 of core.py:442-469: pawns on ranks 1/8 -> best non-pawn class.  Returns (labels, fixes)."""
    labels = list(labels)
    fixes = []
    order = np.argsort(probs)  # ascending, same tie behaviour as the reference call
    for i, (lab, name) in enumerate(zip(labels, square_names)):
        if name[1] in "18" and lab in ("P", "p"):
            for alt in order[i][::-1]:
                alt_piece = LABEL_NAMES[alt]
                if alt_piece not in ("P", "p"):
                    fixes.append((name, lab, alt_piece, "no_pawns_on_ends"))
                    labels[i] = alt_piece
                    break
    return labels, fixes


def position_from_probabilities(probs: np.ndarray, flip: bool = False):
    names = SQUARE_NAMES_FLIPPED if flip else SQUARE_NAMES_NORMAL
    labels = [LABEL_NAMES[i] for i in np.argmax(probs, axis=1)]
    original_fen = board_fen(labels, names)
    fixed, fixes = validate_labels(labels, probs, names)
    return board_fen(fixed, names), original_fen, labels, fixed, fixes
