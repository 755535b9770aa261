"""Integer-level emulation of k_warp_board's coordinate path on the CPU (numpy; float32 steps rounded as the device rounds
them): per 64x64 tile the footprint gate, per 16-pixel segment the float64 anchor, the float32 offsets through the magic-number
FMA, the guard-band flags -- and a check that every pixel the kernel would NOT hand to the literal arithmetic carries exactly
OpenCV's fixed-point coordinates.   python profiles/probes/warp_kernel_emulation.py [n_quads]"""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import geometry as og

FRAC, BAND, MAGIC_BITS = 13, 8, 0x4B400000
ROWS, COLS16 = 88, 7


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32)


def fmaf(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def emulate(minv, H=512, W=512):
    m = minv.reshape(9)
    Xi, Yi = og.warp_coords(minv, 512, 512)                     # the literal arithmetic (pinned to cv2 by the oracle tests)
    stats = dict(px=0, fast_px=0, redo_px=0, wrong=0, tiles_fit=0)
    for ty in range(8):
        for tx in range(8):
            bx, by = tx * 64, ty * 64
            stats["px"] += 4096
            row = np.arange(64)[:, None]
            seg = np.arange(4)[None, :]
            xg, yg = (bx + 16 * seg + 8).astype(np.float64), (by + row).astype(np.float64)
            Xc, Yc, Wc = m[0] * xg + (m[1] * yg + m[2]), m[3] * xg + (m[4] * yg + m[5]), m[6] * xg + (m[7] * yg + m[8])
            with np.errstate(all="ignore"):
                q = 1.0 / Wc
                Uc, Vc = Xc * q * 262144.0, Yc * q * 262144.0
                Bx, By = f32(-Uc * m[6] + m[0] * 262144.0), f32(-Vc * m[6] + m[3] * 262144.0)
                Wcf, m6f = f32(Wc), f32(np.full_like(Wc, m[6]))
                ra, rb = f32(1.0 / fmaf(m6f, f32(-8.0), Wcf).astype(np.float64)), f32(1.0 / fmaf(m6f, f32(7.0), Wcf).astype(np.float64))
                oxa, oxb = Bx * (f32(-8.0) * ra), Bx * (f32(7.0) * rb)
                oya, oyb = By * (f32(-8.0) * ra), By * (f32(7.0) * rb)
                Uf, Vf = f32(Uc), f32(Vc)
                fast = (np.abs(oxa) < 4e6) & (np.abs(oxb) < 4e6) & (np.abs(oya) < 4e6) & (np.abs(oyb) < 4e6)
                fast &= (np.abs(Wcf) > 1e-30) & (np.abs(Wcf) < 1e30) & (np.abs(m6f * f32(8.0)) <= f32(0.25) * np.abs(Wcf))
                fast &= (np.abs(Uf) < 5e8) & (np.abs(Vf) < 5e8)
                kpx = np.float32(1.0 / 262144.0)
                z = np.float32(0)
                ex0 = np.floor((Uf + np.minimum(np.minimum(oxa, oxb), z)) * kpx)
                ex1 = np.floor((Uf + np.maximum(np.maximum(oxa, oxb), z)) * kpx)
                ey0 = np.floor((Vf + np.minimum(np.minimum(oya, oyb), z)) * kpx)
                ey1 = np.floor((Vf + np.maximum(np.maximum(oya, oyb), z)) * kpx)
            FAR = 1 << 20
            x_min = int(np.where(fast, ex0, -FAR).min()); x_max = int(np.where(fast, ex1, FAR).max())
            y_min = int(np.where(fast, ey0, -FAR).min()); y_max = int(np.where(fast, ey1, FAR).max())
            x_lo, y_lo = (x_min - 1) & ~15, y_min - 1
            ncol16, nrows = (x_max + 3 - x_lo + 15) >> 4, y_max + 3 - y_lo
            fits = 1 <= ncol16 <= COLS16 and 1 <= nrows <= ROWS
            if not fits:
                continue
            stats["tiles_fit"] += 1
            lim_x, lim_y = ncol16 * 16 - 1, nrows - 1
            Ul, Vl = Uc - float(x_lo * 262144), Vc - float(y_lo * 262144)
            k_round = (1 << (FRAC - 1)) + BAND // 2 - MAGIC_BITS
            Cx = np.where(fast, np.rint(np.where(fast, Ul, 0.0)), 0).astype(np.int64) + k_round
            Cy = np.where(fast, np.rint(np.where(fast, Vl, 0.0)), 0).astype(np.int64) + k_round
            fBx, fBy = np.where(fast, Bx, np.float32(0)), np.where(fast, By, np.float32(0))
            fW, fm6 = np.where(fast, Wcf, np.float32(1)), np.where(fast, m6f, np.float32(0))
            for k in range(16):
                d = np.float32(k - 8)
                r = f32(1.0 / fmaf(fm6, d, fW).astype(np.float64))
                s = d * r
                ix = fmaf(fBx, s, np.float32(12582912.0)).view(np.int32).astype(np.int64) + Cx
                iy = fmaf(fBy, s, np.float32(12582912.0)).view(np.int32).astype(np.int64) + Cy
                redo = ~fast | ((ix & 8191) < BAND) | ((iy & 8191) < BAND)
                lx, ly = ix >> (FRAC + 5), iy >> (FRAC + 5)
                assert ((lx >= 0) & (lx < lim_x) & (ly >= 0) & (ly < lim_y)).all(), "tap outside the staged patch"
                gx = (ix >> FRAC) + 32 * x_lo                    # back to image coordinates
                gy = (iy >> FRAC) + 32 * y_lo
                ex = Xi[by:by + 64, bx:bx + 64].reshape(64, 4, 16)[:, :, k]
                ey = Yi[by:by + 64, bx:bx + 64].reshape(64, 4, 16)[:, :, k]
                bad = ~redo & ((gx != ex) | (gy != ey))
                stats["wrong"] += int(bad.sum())
                stats["redo_px"] += int((redo & fast).sum())
                stats["fast_px"] += int(fast.sum())
    return stats


def main():
    import test_gpu_geometry as tg
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rng = np.random.default_rng(20261018)
    quads = tg._fuzz_quads(rng, n)
    dest = np.array(((0, 0), (512, 0), (512, 512), (0, 512)), np.float32)
    tot = dict(px=0, fast_px=0, redo_px=0, wrong=0, tiles_fit=0)
    for i, qd in enumerate(quads):
        try:
            M = og.perspective_matrix(og.scale_quadrangle(qd.reshape(4, 1, 2), (512, 512)).reshape(4, 2), dest)
            minv = og.invert3(np.asarray(M, np.float64))
        except Exception:
            continue
        if not np.isfinite(minv).all():
            continue
        st = emulate(minv)
        for k in tot:
            tot[k] += st[k]
        if st["wrong"]:
            print("quad", i, qd.tolist(), st)
    print(f"{n} quads: {tot['px']} px, float32 path {100 * tot['fast_px'] / tot['px']:.1f} %, guard-band pixels {100 * tot['redo_px'] / max(1, tot['fast_px']):.3f} % of those, "
          f"tiles staged {tot['tiles_fit']}, pixels with wrong coordinates outside the band: {tot['wrong']}")


if __name__ == "__main__":
    main()
