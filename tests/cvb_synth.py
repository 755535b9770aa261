"""Deterministic synthetic inputs shared by CPU and GPU tests (no reference data needed)."""
import cv2
import numpy as np


def quad_mask(rng, noise=0.0, sigma=0.0, holes=0, specks=0, size=256):
    """A filled convex quadrilateral (board-like) in a 256x256 mask, optionally made ragged."""
    c = size / 2 + rng.uniform(-20, 20, 2)
    half = rng.uniform(60, 105)
    base = np.array([[half, -half], [-half, -half], [-half, half], [half, half]], np.float64)
    ang = rng.uniform(-0.5, 0.5)
    rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    pts = base @ rot.T + c + rng.uniform(-14, 14, (4, 2))
    m = np.zeros((size, size), np.uint8)
    cv2.fillConvexPoly(m, np.round(pts).astype(np.int32), 255)
    for _ in range(holes):
        p = c + rng.uniform(-40, 40, 2)
        cv2.circle(m, (int(p[0]), int(p[1])), int(rng.integers(2, 9)), 0, -1)
    for _ in range(specks):
        p = rng.uniform(4, size - 4, 2)
        cv2.circle(m, (int(p[0]), int(p[1])), int(rng.integers(1, 6)), 255, -1)
    if sigma > 0 or noise > 0:
        f = cv2.GaussianBlur(m.astype(np.float32) / 255, (0, 0), max(sigma, 0.5))
        f = f + rng.normal(0, noise, f.shape).astype(np.float32)
        f = cv2.GaussianBlur(f, (0, 0), 1.0)
        m = np.where(f > 0.5, 255, 0).astype(np.uint8)
    return m


def mask_suite(seed=7, n=64):
    rng = np.random.default_rng(seed)
    out = [np.zeros((256, 256), np.uint8), np.full((256, 256), 255, np.uint8)]
    one = np.zeros((256, 256), np.uint8)
    one[100, 100] = 255
    out.append(one)
    frame = np.full((256, 256), 255, np.uint8)
    frame[30:220, 40:210] = 0   # big hole touching nothing: outer border + hole border
    out.append(frame)
    two = np.zeros((256, 256), np.uint8)
    two[10:200, 10:120] = 255
    two[20:250, 130:250] = 255  # two large blobs
    out.append(two)
    while len(out) < n:
        k = len(out) % 4
        if k == 0:
            out.append(quad_mask(rng))
        elif k == 1:
            out.append(quad_mask(rng, noise=rng.uniform(0.1, 0.3), sigma=rng.uniform(1.5, 4)))
        elif k == 2:
            out.append(quad_mask(rng, holes=int(rng.integers(1, 5)), specks=int(rng.integers(0, 12))))
        else:
            out.append(quad_mask(rng, noise=rng.uniform(0.2, 0.45), sigma=rng.uniform(2, 5), holes=2, specks=6))
    return np.stack(out)


def board_image(rng, size=512):
    """A textured BGR image with a checkerboard inside a random quadrilateral; returns (img, quad_256) with the quad in
    the reference's corner order TR, TL, BL, BR and mask-frame (256) integer coordinates."""
    img = rng.integers(0, 256, (size // 8, size // 8, 3), dtype=np.uint8)
    img = cv2.resize(img, (size, size), interpolation=cv2.INTER_CUBIC)
    q = np.array([[200, 56], [56, 56], [56, 200], [200, 200]], np.int32) + rng.integers(-24, 25, (4, 2)).astype(np.int32)
    src = np.array([[0, 0], [8, 0], [8, 8], [0, 8]], np.float32)
    dst = (q[[1, 0, 3, 2]] * 2).astype(np.float32)
    M = cv2.getPerspectiveTransform(src, dst)
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float32)
    Minv = np.linalg.inv(M)
    den = Minv[2, 0] * xs + Minv[2, 1] * ys + Minv[2, 2]
    u = (Minv[0, 0] * xs + Minv[0, 1] * ys + Minv[0, 2]) / den
    v = (Minv[1, 0] * xs + Minv[1, 1] * ys + Minv[1, 2]) / den
    inside = (u >= 0) & (u < 8) & (v >= 0) & (v < 8)
    check = ((np.floor(u) + np.floor(v)) % 2 == 0)
    a, b = rng.integers(150, 256, 3), rng.integers(0, 100, 3)
    img[inside & check] = a
    img[inside & ~check] = b
    return img, q


def comb_mask(rng, teeth=None):
    """A board-sized blob whose border is a one-pixel comb: tens of thousands of border points in ONE contour, far beyond the
    shared-memory capacity of the contour kernels (8,192 points) -- the large-capacity path must take it."""
    m = np.zeros((256, 256), np.uint8)
    x0, x1 = int(rng.integers(8, 40)), int(rng.integers(216, 248))
    y0, y1 = int(rng.integers(8, 40)), int(rng.integers(216, 248))
    step = teeth or int(rng.integers(2, 4))
    m[y1 - 5:y1, x0:x1] = 255                            # the spine ...
    m[y0:y1 - 5, x0:x1:step] = 255                       # ... and one-pixel teeth standing on it: a single component
    return m


def fuzz_masks(seed, n):
    """Masks for the large mask->quad fuzz: board-like quadrilaterals made ragged by blur + noise, holes and specks, thin
    frames (hole contours that pass the area filter), several blobs, salt-and-pepper fields with thousands of borders, and
    combs whose single contour exceeds every shared-memory capacity."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        k = len(out) % 16
        if k < 5:
            out.append(quad_mask(rng, noise=rng.uniform(0.05, 0.5), sigma=rng.uniform(0.8, 5), holes=int(rng.integers(0, 4)),
                                 specks=int(rng.integers(0, 10))))
        elif k < 9:
            out.append(quad_mask(rng, holes=int(rng.integers(0, 6)), specks=int(rng.integers(0, 16))))
        elif k < 11:
            out.append(quad_mask(rng))
        elif k == 11:                                    # a frame: the hole border is the board candidate
            m = quad_mask(rng)
            inner = cv2.erode(m, np.ones((3, 3), np.uint8), iterations=int(rng.integers(2, 12)))
            out.append(np.where(inner > 0, 0, m).astype(np.uint8))
        elif k == 12:                                    # two blobs
            m = quad_mask(rng)
            m[:, :int(rng.integers(100, 156))] = 0
            out.append(np.maximum(m, np.roll(m, int(rng.integers(-120, -90)), 1)))
        elif k == 13:                                    # salt and pepper on a board
            m = quad_mask(rng)
            flip = rng.random(m.shape) < rng.uniform(0.01, 0.3)
            out.append(np.where(flip, 255 - m, m).astype(np.uint8))
        elif k == 14:
            out.append(comb_mask(rng))
        else:                                            # a board with a comb-like ragged edge from heavy noise
            out.append(quad_mask(rng, noise=rng.uniform(0.5, 1.2), sigma=0.6))
    return np.stack(out)
