"""GPU parity of the JPEG decode front-end (csrc/jpeg.cu): byte-identical to the decoder the reference
calls (``cv2.imdecode`` / ``cv2.imread``, scripts/eval/evaluate.py:147, app/computeroot/cv_endpoint.py:151-153) on the
reference's ``data/test`` JPEGs (also against their frozen sha1), on re-encoded synthetic images at several sizes,
qualities and restart intervals, and through the whole image->FEN pipeline."""
import hashlib
import json

import cv2
import numpy as np
import pytest
import torch

from conftest import GOLDEN, WEIGHTS

pytestmark = pytest.mark.gpu
FILES = sorted((GOLDEN / "data_test").glob("*/*"))


@pytest.fixture(params=["host", "device"])
def huffman_path(request, monkeypatch):
    """Entropy decoding on the host threads (small batches) or on the device (one warp per image; batches >= 1024 by default):
    CVB_JPEG_DEVICE_MIN forces either for any batch size."""
    monkeypatch.setenv("CVB_JPEG_DEVICE_MIN", "1" if request.param == "device" else "0")
    return request.param


def test_data_test_images_byte_identical(engine, huffman_path):
    man = {e["file"]: e["image_sha1"] for e in json.load(open(GOLDEN / "manifest.json"))["images"]}
    streams = [f.read_bytes() for f in FILES]
    got = engine.decode_jpeg(streams).cpu().numpy()
    assert got.shape == (38, 512, 512, 3)
    for f, s, g in zip(FILES, streams, got):
        assert np.array_equal(g, cv2.imdecode(np.frombuffer(s, np.uint8), cv2.IMREAD_COLOR)), f.name
        assert hashlib.sha1(g.tobytes()).hexdigest() == man[f"{f.parent.name}/{f.name}"], f.name


@pytest.mark.parametrize("h,w,quality,rst,n", [(16, 16, 90, 0, 3), (48, 32, 50, 0, 5), (512, 512, 100, 0, 2), (768, 1024, 75, 7, 2),
                                               (64, 64, 10, 1, 70), (256, 256, 95, 16, 9)])
def test_reencoded_synthetic_images(engine, huffman_path, h, w, quality, rst, n):
    rng = np.random.default_rng(h * 7 + w + quality)
    streams, want = [], []
    for i in range(n):
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([(xx * (i + 1) + yy) % 256, (yy * 3 + i * 40) % 256, (xx + yy * 2) % 256], -1).astype(np.float32)
        img += rng.normal(scale=40 if i % 2 else 4, size=img.shape)      # noise drives coefficients to the range-limit table
        ok, buf = cv2.imencode(".jpg", np.clip(img, 0, 255).astype(np.uint8),
                               [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
        assert ok
        streams.append(buf.tobytes())
        want.append(cv2.imdecode(buf, cv2.IMREAD_COLOR))
    got = engine.decode_jpeg(streams).cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], want[i]), f"image {i}: {np.abs(got[i].astype(int) - want[i].astype(int)).max()} levels off"


def test_unsupported_stream_fails_loudly(engine):
    from chessvision import _native
    ok, buf = cv2.imencode(".jpg", np.zeros((64, 64, 3), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(_native.NativeError):
        engine.decode_jpeg([buf.tobytes()])
    good = FILES[0].read_bytes()
    ok, small = cv2.imencode(".jpg", np.zeros((64, 64, 3), np.uint8))
    with pytest.raises(_native.NativeError):
        engine.decode_jpeg([good, small.tobytes()])                      # every stream of a batch must have the batch's H x W
    with pytest.raises(_native.NativeError):
        engine.decode_jpeg([good, good[:300]])                           # headers cut off


def test_decode_then_image_to_fen_equals_the_cv2_front_end():
    """files -> FEN with the decode on the device == cv2.imread + ChessVision.process_images."""
    from chessvision import ChessVision, decode
    cv = ChessVision(board_extractor_weights=str(WEIGHTS / "best_extractor.pth"), classifier_weights=str(WEIGHTS / "best_classifier.pth"),
                     classifier_model_id="resnet18", max_batch=8)
    files = FILES[:6]
    dev = decode.imread_batch(files, engine=cv._engine)
    via_device = cv.process_images(dev.cpu().numpy())
    via_cv2 = cv.process_images(np.stack([cv2.imread(str(f)) for f in files]))
    for a, b in zip(via_device, via_cv2):
        assert (a.position is None) == (b.position is None)
        if a.position is not None:
            assert a.position.fen == b.position.fen and a.position.original_fen == b.position.original_fen
        assert np.array_equal(a.board_extraction.binary_mask, b.board_extraction.binary_mask)


def test_device_huffman_equals_host_huffman_on_damaged_streams(engine, monkeypatch):
    """Truncated and bit-flipped entropy-coded segments: the device decoder takes the same decisions as the host decoder
    (same pixels, or both refuse the batch), and a large batch takes the device path by default."""
    from chessvision import _native
    rng = np.random.default_rng(3)
    base = [f.read_bytes() for f in FILES[:8]]
    streams = []
    for k in range(48):
        b = bytearray(base[k % len(base)])
        if k % 3 == 0:
            b = b[: len(b) - int(rng.integers(1, len(b) // 2))]                  # cut inside the scan
        else:
            start = len(b) // 3
            for _ in range(int(rng.integers(1, 6))):
                i = int(rng.integers(start, len(b) - 2))
                b[i] ^= 1 << int(rng.integers(0, 8))
        streams.append(bytes(b))

    def run(mode, items):
        monkeypatch.setenv("CVB_JPEG_DEVICE_MIN", mode)
        try:
            return engine.decode_jpeg(items).cpu().numpy()
        except _native.NativeError:
            return None

    agree = 0
    for s_ in streams:
        host, dev = run("0", [s_]), run("1", [s_])
        assert (host is None) == (dev is None)
        if host is not None:
            assert np.array_equal(host, dev)
            agree += 1
    assert agree > 0
    monkeypatch.delenv("CVB_JPEG_DEVICE_MIN")
    many = [base[i % len(base)] for i in range(1100)]                            # >= 1024: device path by default
    got = engine.decode_jpeg(many).cpu().numpy()
    for i in (0, 7, 1099):
        assert np.array_equal(got[i], cv2.imdecode(np.frombuffer(many[i], np.uint8), cv2.IMREAD_COLOR))
