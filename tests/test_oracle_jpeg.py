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
