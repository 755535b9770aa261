"""ORACLE (test infrastructure, never on the product path): CPU restatement of the JPEG decode in front of the image->FEN
path — SURVEY.md §8(f) row n2.  The reference decodes with ``cv2.imread`` / ``cv2.imdecode(..., IMREAD_COLOR)``
(scripts/eval/evaluate.py:147, app/computeroot/cv_endpoint.py:151-153), i.e. OpenCV's bundled libjpeg-turbo
(``opencv-python==4.11.0.86`` in the reference's uv.lock; this image: OpenCV 4.13.0 with libjpeg-turbo 3.1.2) with its
default settings: ``JDCT_ISLOW``, fancy (triangle) chroma upsampling, separate YCbCr->RGB conversion, output in BGR order.
That library is not under /root/reference, so its published algorithm is restated here (numpy + plain Python for the
sequential Huffman part) and pinned against the live ``cv2.imdecode`` of this image on the reference's own 38
``data/test`` JPEGs (tests/test_oracle_jpeg.py), bit for bit.

Scope (what the reference's data needs): baseline sequential DCT (SOF0), 8-bit, three components, 4:2:0 sampling (luma
2x2, chroma 1x1), dimensions that are multiples of 16, optional restart intervals, EXIF orientation absent or 1.  Anything
else raises ``ValueError`` (the product path returns an error code for the same inputs).

Algorithm sources (libjpeg-turbo 3.1.2): jdhuff.c (entropy decoding), jidctint.c ``jpeg_idct_islow`` (dequantisation +
inverse DCT, CONST_BITS 13 / PASS1_BITS 2, range limiting through the masked table), jdsample.c ``h2v2_fancy_upsample``
with the edge-row replication of jdmainct.c, jdcolor.c ``ycc_rgb_convert`` (16-bit fixed-point tables).
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
                   62, 63])


class _Bits:
    """MSB-first bit reader over the entropy-coded segment (0xFF00 unstuffed, stops at markers)."""

    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _fill(self):
        while self.n <= 24:
            b = 0
            if self.p < len(self.d):
                b = self.d[self.p]
                if b == 0xFF:
                    nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                    if nxt == 0:
                        self.p += 2
                    else:
                        b = 0          # a marker: feed zeros (jdhuff.c does the same until the MCU ends)
                else:
                    self.p += 1
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFF
            self.n += 8

    def get(self, k: int) -> int:
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def restart(self):
        """Byte-align and step over the RSTn marker."""
        self.acc = self.n = 0
        while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _huff_table(counts, symbols):
    """code length -> (mincode, maxcode, first symbol index), canonical JPEG codes (jdhuff.c jpeg_make_d_derived_tbl)."""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        n = counts[length - 1]
        if n:
            table[length] = (code, code + n - 1, k)
        code = (code + n) << 1
        k += n
    return table, symbols


def _decode_symbol(bits: _Bits, tab) -> int:
    table, symbols = tab
    code = 0
    for length in range(1, 17):
        code = (code << 1) | bits.get(1)
        t = table.get(length)
        if t and t[0] <= code <= t[1]:
            return symbols[t[2] + code - t[0]]
    raise ValueError("corrupt JPEG: bad Huffman code")


def _extend(v: int, s: int) -> int:
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def parse(data: bytes):
    """Headers -> dict(h, w, qt[comp] (natural order), dc/ac table per component, restart interval, scan start)."""
    if data[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG")
    i, q, dc, ac, out = 2, {}, {}, {}, {"ri": 0}
    while True:
        if data[i] != 0xFF:
            raise ValueError("corrupt JPEG: marker expected")
        m = data[i + 1]
        if m == 0xFF:
            i += 1
            continue
        ln = (data[i + 2] << 8) | data[i + 3]
        seg = data[i + 4:i + 2 + ln]
        if m == 0xDB:
            j = 0
            while j < len(seg):
                pq, tq = seg[j] >> 4, seg[j] & 15
                if pq:
                    raise ValueError("16-bit quantisation tables are not supported")
                t = np.zeros(64, np.int32)
                t[ZIGZAG] = np.frombuffer(seg[j + 1:j + 65], np.uint8)
                q[tq] = t
                j += 65
        elif m == 0xC4:
            j = 0
            while j < len(seg):
                tc, th = seg[j] >> 4, seg[j] & 15
                counts = list(seg[j + 1:j + 17])
                n = sum(counts)
                (ac if tc else dc)[th] = _huff_table(counts, list(seg[j + 17:j + 17 + n]))
                j += 17 + n
        elif m == 0xC0:
            if seg[0] != 8 or seg[5] != 3:
                raise ValueError("only 8-bit three-component baseline JPEGs are supported")
            out["h"], out["w"] = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            comps = [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(3)]
            if [(c[1], c[2]) for c in comps] != [(2, 2), (1, 1), (1, 1)] or out["h"] % 16 or out["w"] % 16:
                raise ValueError("only 4:2:0 JPEGs with dimensions that are multiples of 16 are supported")
            out["comps"] = comps
        elif m in (0xC1, 0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("only baseline (SOF0) JPEGs are supported")
        elif m == 0xDD:
            out["ri"] = (seg[0] << 8) | seg[1]
        elif m == 0xE1 and seg[:6] == b"Exif\0\0":
            if exif_orientation(seg[6:]) not in (0, 1):
                raise ValueError("EXIF orientation other than 1 is not supported")
        elif m == 0xDA:
            sel = {seg[1 + 2 * k]: (seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(seg[0])}
            out["tabs"] = [(dc[sel[c[0]][0]], ac[sel[c[0]][1]]) for c in out["comps"]]
            out["qt"] = [q[c[3]] for c in out["comps"]]
            out["scan"] = i + 2 + ln
            return out
        i += 2 + ln


def exif_orientation(tiff: bytes) -> int:
    """Orientation tag (0x0112) of IFD0, 0 if absent."""
    if len(tiff) < 8:
        return 0
    le = tiff[:2] == b"II"
    u16 = lambda o: int.from_bytes(tiff[o:o + 2], "little" if le else "big")
    u32 = lambda o: int.from_bytes(tiff[o:o + 4], "little" if le else "big")
    ifd = u32(4)
    if ifd + 2 > len(tiff):
        return 0
    for k in range(u16(ifd)):
        e = ifd + 2 + 12 * k
        if e + 12 <= len(tiff) and u16(e) == 0x0112:
            return u16(e + 8)
    return 0


def coefficients(data: bytes):
    """Entropy decoding (jdhuff.c): quantised coefficients in natural order, int16 [3][rows of blocks][cols][64]."""
    hd = parse(data)
    mh, mw = hd["h"] // 16, hd["w"] // 16
    coef = [np.zeros((2 * mh, 2 * mw, 64), np.int16), np.zeros((mh, mw, 64), np.int16), np.zeros((mh, mw, 64), np.int16)]
    bits = _Bits(data, hd["scan"])
    pred = [0, 0, 0]
    layout = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (2, 0, 0)]   # component, block row, block col inside the MCU
    for mcu in range(mh * mw):
        if hd["ri"] and mcu and mcu % hd["ri"] == 0:
            bits.restart()
            pred = [0, 0, 0]
        my, mx = divmod(mcu, mw)
        for c, by, bx in layout:
            dct, act = hd["tabs"][c]
            blk = coef[c][my * (2 if c == 0 else 1) + by, mx * (2 if c == 0 else 1) + bx]
            s = _decode_symbol(bits, dct)
            pred[c] += _extend(bits.get(s), s) if s else 0
            blk[0] = pred[c]
            k = 1
            while k < 64:
                rs = _decode_symbol(bits, act)
                r, s = rs >> 4, rs & 15
                if s == 0:
                    if r != 15:
                        break
                    k += 16
                    continue
                k += r
                blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                k += 1
    return hd, coef


# ---------------------------------------------------------------------------------------------------- jidctint.c (islow)
def _fix(x):
    return int(x * (1 << 13) + 0.5)


F_0_298, F_0_390, F_0_541, F_0_765, F_0_899, F_1_175 = _fix(0.298631336), _fix(0.390180644), _fix(0.541196100), _fix(0.765366865), _fix(0.899976223), _fix(1.175875602)
F_1_501, F_1_847, F_1_961, F_2_053, F_2_562, F_3_072 = _fix(1.501321110), _fix(1.847759065), _fix(1.961570560), _fix(2.053119869), _fix(2.562915447), _fix(3.072711026)


def _idct_1d(v, shift):
    """One pass of jpeg_idct_islow over the LAST axis of v (int64 [...,8]); DESCALE by `shift` bits."""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * F_0_541
    tmp2 = z1 + z3 * (-F_1_847)
    tmp3 = z1 + z2 * F_0_765
    z2, z3 = v[..., 0], v[..., 4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F_1_175
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F_0_298, tmp1 * F_2_053, tmp2 * F_3_072, tmp3 * F_1_501
    z1, z2, z3, z4 = z1 * (-F_0_899), z2 * (-F_2_562), z3 * (-F_1_961) + z5, z4 * (-F_0_390) + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    r = (1 << (shift - 1))
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], -1)
    return (out + r) >> shift


def idct_blocks(coef: np.ndarray, qt: np.ndarray) -> np.ndarray:
    """int16 [R,C,64] quantised coefficients -> uint8 plane [8R, 8C] (dequantise, 2-D islow IDCT, masked range limit)."""
    R, C, _ = coef.shape
    w = (coef.astype(np.int64) * qt.astype(np.int64)).reshape(R, C, 8, 8)
    w = _idct_1d(w.swapaxes(-1, -2), 13 - 2).swapaxes(-1, -2)    # pass 1: columns, keeps PASS1_BITS extra bits
    w = _idct_1d(w, 13 + 2 + 3)                                  # pass 2: rows
    idx = w & 1023                                               # range_limit[(x) & RANGE_MASK], table centred on 128
    px = np.where(idx < 128, idx + 128, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896)))
    return px.astype(np.uint8).transpose(0, 2, 1, 3).reshape(8 * R, 8 * C)


# ---------------------------------------------------------------------------------------------------- jdsample.c
def upsample_h2v2_fancy(p: np.ndarray) -> np.ndarray:
    """uint8 [h,w] -> [2h,2w]: triangle filter, 3/4 nearer + 1/4 further in each axis, edge rows / columns replicated,
    rounding constants 8 and 7 alternating by output column (h2v2_fancy_upsample)."""
    h, w = p.shape
    q = p.astype(np.int32)
    above, below = np.vstack([q[:1], q[:-1]]), np.vstack([q[1:], q[-1:]])
    out = np.empty((2 * h, 2 * w), np.int32)
    for v, other in ((0, above), (1, below)):
        col = 3 * q + other                                      # thiscolsum for every column
        last = np.hstack([col[:, :1], col[:, :-1]])
        nxt = np.hstack([col[:, 1:], col[:, -1:]])
        out[v::2, 0::2] = (3 * col + last + 8) >> 4              # at the left edge last == this: (4*this + 8) >> 4
        out[v::2, 1::2] = (3 * col + nxt + 7) >> 4               # at the right edge next == this: (4*this + 7) >> 4
    return out.astype(np.uint8)


# ---------------------------------------------------------------------------------------------------- jdcolor.c
def ycc_to_bgr(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    fix = lambda x: int(x * 65536 + 0.5)
    xb, xr = cb.astype(np.int32) - 128, cr.astype(np.int32) - 128
    r = y.astype(np.int32) + ((fix(1.40200) * xr + 32768) >> 16)
    b = y.astype(np.int32) + ((fix(1.77200) * xb + 32768) >> 16)
    g = y.astype(np.int32) + ((-fix(0.34414) * xb + 32768 - fix(0.71414) * xr) >> 16)
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def imdecode(data: bytes) -> np.ndarray:
    """``cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)`` for the supported subset: u8 [H,W,3] BGR."""
    hd, coef = coefficients(data)
    y, cb, cr = (idct_blocks(coef[c], hd["qt"][c]) for c in range(3))
    return ycc_to_bgr(y, upsample_h2v2_fancy(cb), upsample_h2v2_fancy(cr))
