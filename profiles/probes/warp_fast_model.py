"""Error budget of k_warp_board's float32 coordinate path, modelled on the CPU (numpy float64 carrying every float32 rounding).

Per 16-pixel row segment the kernel anchors U = 32*X/W at the segment centre in float64 and adds the float32 offset
    off(d) = B * d / W(d),  B = 32*m0 - U_c*m6,  d = -8 .. 7
scaled by 2^13 (magic-number rounding inside the FMA).  The script measures max |fast - true| in 2^-13 units of 1/32 px over
random and extreme board quads, with the reciprocal perturbed by +-1 ulp (MUFU.RCP's stated bound), and checks that every pixel
outside the guard band rounds to OpenCV's integer.   python profiles/probes/warp_fast_model.py [n_quads]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import geometry as og

K = 13
BAND = 8          # kernel: ((total + 4096 + 4) & 8191) < 8 -> exact path (covers |error| < 4)


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32).astype(np.float64)


def model(minv, rng, ulp_sign):
    m = minv.reshape(9)
    ys = np.arange(512, dtype=np.float64)[:, None]
    xs = np.arange(512)
    bx = ((xs // 64) * 64).astype(np.float64)[None, :]
    x1 = (xs % 64).astype(np.float64)[None, :]
    X0 = m[0] * bx + m[1] * ys + m[2]
    Y0 = m[3] * bx + m[4] * ys + m[5]
    W0 = m[6] * bx + m[7] * ys + m[8]
    Xi, Yi = og.warp_coords(minv, 512, 512)
    x1c = ((xs % 64) // 16 * 16 + 8).astype(np.float64)[None, :]
    d = x1 - x1c
    Xc, Yc, Wc = X0 + m[0] * x1c, Y0 + m[3] * x1c, W0 + m[6] * x1c
    out = []
    true_w = W0 + m[6] * x1
    for num_c, num0, mm, exact in ((Xc, X0, m[0], Xi), (Yc, Y0, m[3], Yi)):
        Uc = 32.0 * num_c / Wc
        B = 32.0 * mm - Uc * m[6]
        B13 = f32(B * 2.0 ** K)
        Wf = f32(f32(m[6]) * d + f32(Wc))
        r = f32(1.0 / Wf)
        r = f32(r * (1.0 + ulp_sign * 2.0 ** -23))           # +-1 ulp of the reciprocal
        s = f32(d * r)
        magic = 1.5 * 2.0 ** 23
        t = f32(B13 * s + magic) - magic
        total = t + np.rint(Uc * 2.0 ** K)
        true = (num0 + mm * x1) * (32.0 / true_w) * 2.0 ** K
        # the kernel's per-thread gate: offsets at both segment ends inside the magic range, W of one sign over the tile
        seg_ok = (np.abs(B13 * s) < 2.0 ** 22 * 0.97).reshape(512, 32, 16).all(axis=2).repeat(16, axis=1).reshape(512, 512)
        wt = true_w.reshape(8, 64, 8, 64)
        tile_ok = ((wt > 0).all(axis=(1, 3)) | (wt < 0).all(axis=(1, 3)))[:, None, :, None].repeat(64, 1).repeat(64, 3).reshape(512, 512)
        ok_range = seg_ok & tile_ok & np.isfinite(total)
        err = np.where(ok_range, total - true, 0.0)
        ti = total.astype(np.int64)
        flagged = ((ti + 4096 + BAND // 2) & 8191) < BAND
        fast_int = (ti + 4096) >> K
        wrong = ok_range & ~flagged & (fast_int != exact)
        out.append((np.abs(err).max(), flagged[ok_range].mean(), int(wrong.sum()), float((~ok_range).mean())))
    return out


def random_quad(rng, extreme):
    if extreme:
        c = rng.uniform(-40, 296, (4, 2))
        c = c[np.argsort(np.arctan2(c[:, 1] - c[:, 1].mean(), c[:, 0] - c[:, 0].mean()))]
        return np.round(c).astype(np.int32)
    cx, cy = rng.uniform(90, 166, 2)
    r = rng.uniform(50, 120)
    a0 = rng.uniform(0, 2 * np.pi)
    ang = a0 + np.array([0, 0.5, 1.0, 1.5]) * np.pi + rng.uniform(-0.25, 0.25, 4)
    rad = r * rng.uniform(0.75, 1.25, 4)
    return np.round(np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)).astype(np.int32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(7)
    worst, frac, wrong, skipped = 0.0, [], 0, []
    dest = np.array(((0, 0), (512, 0), (512, 512), (0, 512)), np.float32)
    for i in range(n):
        q = random_quad(rng, extreme=(i % 4 == 3))
        try:
            M = og.perspective_matrix(og.scale_quadrangle(q.reshape(4, 1, 2), (512, 512)).reshape(4, 2), dest)
            minv = og.invert3(np.asarray(M, np.float64))
        except Exception:
            continue
        if not np.isfinite(minv).all():
            continue
        for sign in (-1.0, 1.0):
            for e, f, w, sk in model(minv, rng, sign):
                worst = max(worst, e)
                frac.append(f)
                wrong += w
                skipped.append(sk)
    print(f"quads {n}: max |fast - true| = {worst:.3f} (2^-{K} units), flagged {np.mean(frac) * 100:.3f} % per coordinate, "
          f"wrong integers outside the band: {wrong}, pixels outside the magic range: {np.mean(skipped) * 100:.2f} %")


if __name__ == "__main__":
    main()
