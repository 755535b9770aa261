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


def test_data_test_images_byte_identical(engine):
    man = {e["file"]: e["image_sha1"] for e in json.load(open(GOLDEN / "manifest.json"))["images"]}
    streams = [f.read_bytes() for f in FILES]
    got = engine.decode_jpeg(streams).cpu().numpy()
    assert got.shape == (38, 512, 512, 3)
    for f, s, g in zip(FILES, streams, got):
        assert np.array_equal(g, cv2.imdecode(np.frombuffer(s, np.uint8), cv2.IMREAD_COLOR)), f.name
        assert hashlib.sha1(g.tobytes()).hexdigest() == man[f"{f.parent.name}/{f.name}"], f.name


@pytest.mark.parametrize("h,w,quality,rst,n", [(16, 16, 90, 0, 3), (48, 32, 50, 0, 5), (512, 512, 100, 0, 2), (768, 1024, 75, 7, 2),
                                               (64, 64, 10, 1, 70), (256, 256, 95, 16, 9)])
def test_reencoded_synthetic_images(engine, h, w, quality, rst, n):
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
