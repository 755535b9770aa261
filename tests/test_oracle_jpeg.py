"""Pins ``oracle/jpeg.py`` to the third-party decoder the reference calls — the live ``cv2.imdecode`` of
this image (OpenCV with libjpeg-turbo) — on the reference's own ``data/test`` JPEGs, bit for bit, and checks the host half
of the product decoder (marker parsing + Huffman decoding in ``csrc/jpeg.cu``, no GPU involved) against the oracle's
coefficients.  CPU only."""
import hashlib
import json

import cv2
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import jpeg as oj

FILES = sorted((GOLDEN / "data_test").glob("*/*"))
RESTART = [f for f in FILES if b"\xff\xdd" in f.read_bytes()[:1024]]      # the two files with restart intervals + EXIF
SAMPLE = FILES[:3] + RESTART + FILES[-2:]


def test_sample_covers_restart_intervals_and_exif():
    assert len(FILES) == 38 and len(RESTART) == 2


@pytest.mark.parametrize("path", SAMPLE, ids=lambda p: p.name[:8])
def test_oracle_equals_cv2(path):
    data = path.read_bytes()
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(oj.imdecode(data), want)


def test_golden_hashes_of_all_38_images():
    """sha1 of cv2.imdecode's pixels for every data/test image, frozen by the same generator that made the images' other
    golden vectors; a different OpenCV / libjpeg-turbo build that decodes differently shows up here first."""
    man = json.load(open(GOLDEN / "manifest.json"))
    for e in man["images"]:
        img = cv2.imread(str(GOLDEN / "data_test" / e["file"]))
        assert hashlib.sha1(np.ascontiguousarray(img).tobytes()).hexdigest() == e["image_sha1"], e["file"]


@pytest.mark.parametrize("path", [FILES[0], RESTART[0]], ids=lambda p: p.name[:8])
def test_host_huffman_decoder_equals_the_oracle(path):
    from chessvision import _native
    data = path.read_bytes()
    coef, qt = _native.jpeg_coefficients(data)
    hd, want = oj.coefficients(data)
    assert (hd["h"], hd["w"]) == _native.jpeg_info(data)
    flat = np.concatenate([want[c].reshape(-1) for c in range(3)])
    assert np.array_equal(coef, flat)
    assert np.array_equal(qt, np.stack(hd["qt"]).astype(np.uint16))


def test_unsupported_streams_are_rejected():
    from chessvision import _native
    ok, buf = cv2.imencode(".jpg", np.zeros((40, 48, 3), np.uint8))                       # not a multiple of 16
    with pytest.raises(_native.NativeError):
        _native.jpeg_info(buf.tobytes())
    ok, buf = cv2.imencode(".jpg", np.zeros((64, 64, 3), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(_native.NativeError):
        _native.jpeg_info(buf.tobytes())
    with pytest.raises(ValueError):
        oj.parse(buf.tobytes())
    with pytest.raises(_native.NativeError):
        _native.jpeg_info(b"not a jpeg at all")


def test_corrupted_streams_never_crash_the_host_decoder():
    """Memory safety of the host half (csrc/jpeg.cu): random byte damage and truncation anywhere in a valid file must end in
    an error code or a decoded image, never in an out-of-bounds access (run under the normal allocator; a crash kills pytest)."""
    import ctypes as C
    from chessvision import _native
    lib = _native.load_library()
    rng = np.random.default_rng(3)
    coef = np.empty(512 * 512 * 3 // 2 + 64, np.int16)
    qt = np.empty((3, 64), np.uint16)
    ok = err = 0
    for base in (FILES[0].read_bytes(), RESTART[0].read_bytes()):
        for trial in range(300):
            d = bytearray(base)
            if trial % 3 == 0:
                d = d[: int(rng.integers(2, len(d)))]                          # truncation, headers included
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(0, min(len(d), 700 if trial % 2 else len(d))))   # every other trial aims at the headers
                d[pos] = int(rng.integers(0, 256))
            buf = np.frombuffer(bytes(d), np.uint8)
            h, w = C.c_int32(), C.c_int32()
            if lib.cvb_jpeg_info(C.c_void_p(buf.ctypes.data), buf.size, C.byref(h), C.byref(w)) != 0 or (h.value, w.value) != (512, 512):
                err += 1
                continue                                                        # rejected (or a different size: the caller's buffer would not fit)
            rc = lib.cvb_jpeg_coefficients(C.c_void_p(buf.ctypes.data), buf.size, C.c_void_p(coef.ctypes.data), C.c_void_p(qt.ctypes.data))
            ok += rc == 0
            err += rc != 0
    assert ok + err == 600 and ok > 0 and err > 0
