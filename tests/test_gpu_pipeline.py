"""End-to-end parity on the reference's own ``data/test`` images against golden outputs of the UNMODIFIED reference
(tests/golden/, written by oracle/make_golden.py with the same weights).

Contract (BASELINE.json north_star / SURVEY.md §8c):
  * board bytes identical given the reference's quads; labels and both FENs identical given the reference's boards
  * end to end (fp16 UNet vs fp32 reference): found flags identical, corners within +-1 px in the 256x256 mask frame,
    mask IoU >= 0.99, logits max-abs <= 0.15; FEN/label agreement is reported and must be >= 99% of squares
"""
import json
import os
from pathlib import Path

import cv2
import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, WEIGHTS, load_checkpoint

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    man = json.load(open(GOLDEN / "manifest.json"))
    arr = np.load(GOLDEN / "reference_outputs.npz")
    imgs = np.stack([cv2.imread(str(GOLDEN / "data_test" / e["file"])) for e in man["images"]])
    return man, arr, imgs


@pytest.fixture(scope="module")
def trained_engine():
    from chessvision import _native
    ext, cls = WEIGHTS / "best_extractor.pth", WEIGHTS / "best_classifier.pth"
    if not ext.exists() or not cls.exists():
        pytest.fail("weights/best_extractor.pth / best_classifier.pth missing (oracle/train_weights.py writes them)")
    eng = _native.Engine(0, max_batch=16)
    eng.load_unet(load_checkpoint(ext))
    eng.load_resnet18(load_checkpoint(cls))
    yield eng
    eng.close()


def test_images_decode_identically(golden):
    import hashlib
    man, _, imgs = golden
    for e, im in zip(man["images"], imgs):
        assert im.shape == (512, 512, 3)
        assert hashlib.sha1(np.ascontiguousarray(im).tobytes()).hexdigest() == e["image_sha1"], "JPEG decode differs from the build container"


def test_boards_bit_exact_given_reference_quads(trained_engine, golden):
    man, arr, imgs = golden
    idx = [i for i, e in enumerate(man["images"]) if e["found"]]
    quads = np.array([man["images"][i]["quad"] for i in idx], np.int32)
    board = trained_engine.warp_squares(torch.from_numpy(imgs[idx]).cuda(), torch.from_numpy(quads).cuda(),
                                        torch.ones(len(idx), dtype=torch.uint8, device="cuda")).cpu().numpy()
    for k, i in enumerate(idx):
        assert np.array_equal(board[k], arr[f"board_{i}"]), f"{man['images'][i]['file']}: {(board[k] != arr[f'board_{i}']).sum()} bytes differ"


def test_labels_and_fen_bit_exact_given_reference_boards(trained_engine, golden):
    from chessvision._native import fen_strings
    man, arr, _ = golden
    idx = [i for i, e in enumerate(man["images"]) if e["found"]]
    boards = np.stack([arr[f"board_{i}"] for i in idx])
    probs, labels, _, fen = trained_engine.classify(torch.from_numpy(boards).cuda(), False)
    probs, labels = probs.cpu().numpy(), labels.cpu().numpy()
    fens = fen_strings(fen)
    perr = max(np.abs(probs[k] - arr[f"probs_{i}"]).max() for k, i in enumerate(idx))
    flips = sum(int((labels[k] != arr[f"labels_{i}"]).sum()) for k, i in enumerate(idx))
    print(f"classifier given reference boards: probabilities max-abs err {perr:.4f}, label flips {flips}/{64 * len(idx)}")
    assert perr <= 0.05
    assert flips == 0
    for k, i in enumerate(idx):
        e = man["images"][i]
        assert fens[k] == (e["original_fen"], e["fen"]), e["file"]


def test_end_to_end_data_test(trained_engine, golden):
    from chessvision._native import fen_strings
    man, arr, imgs = golden
    n = len(imgs)
    out = trained_engine.image_to_fen(torch.from_numpy(imgs).cuda(), trained_engine.alloc_outputs(n, full=True))
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in out.items()}
    fens = fen_strings(torch.from_numpy(out["fen"]))
    stats = {"images": n, "found_ref": 0, "found_equal": 0, "corner_hist": {}, "fen_identical": 0, "label_flips": 0, "squares": 0,
             "logits_maxabs": 0.0, "mask_iou_min": 1.0}
    for i, e in enumerate(man["images"]):
        ref_logits = arr[f"logits_{i}"].astype(np.float32)
        stats["logits_maxabs"] = max(stats["logits_maxabs"], float(np.abs(out["logits"][i] - ref_logits).max()))
        ref_mask = np.unpackbits(arr[f"mask_{i}"]).reshape(256, 256).astype(bool)
        got_mask = out["mask"][i] > 0
        union = (ref_mask | got_mask).sum()
        if union:
            stats["mask_iou_min"] = min(stats["mask_iou_min"], float((ref_mask & got_mask).sum() / union))
        stats["found_ref"] += int(e["found"])
        stats["found_equal"] += int(bool(out["found"][i]) == e["found"])
        if e["found"] and out["found"][i]:
            d = int(np.abs(out["quad"][i] - np.array(e["quad"])).max())
            stats["corner_hist"][str(d)] = stats["corner_hist"].get(str(d), 0) + 1
            stats["squares"] += 64
            stats["label_flips"] += int((out["labels"][i] != arr[f"labels_{i}"]).sum())
            stats["fen_identical"] += int(fens[i] == (e["original_fen"], e["fen"]))
    print("data/test end-to-end parity:", json.dumps(stats))
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    json.dump(stats, open(ROOT / "gpurun_out" / "parity_data_test.json", "w"), indent=1)
    assert stats["found_equal"] == n
    assert all(int(k) <= 1 for k in stats["corner_hist"]), stats["corner_hist"]
    assert stats["mask_iou_min"] >= 0.99
    assert stats["logits_maxabs"] <= 0.15 + 0.01  # golden logits are stored as fp16 (<= 0.01 quantisation at |x| < 16)
    # the contract is bit-exact labels and FEN strings on every board the reference finds (BASELINE.json north_star)
    assert stats["label_flips"] == 0, stats
    assert stats["fen_identical"] == stats["found_ref"], stats


def test_host_path_equals_device_path(trained_engine, golden):
    _, _, imgs = golden
    eng = trained_engine
    sel = imgs[:37]  # not a multiple of max_batch: exercises the ragged last chunk and both slots
    dev = eng.image_to_fen(torch.from_numpy(sel).cuda(), eng.alloc_outputs(len(sel), full=True))
    torch.cuda.synchronize()
    host_in = torch.from_numpy(sel).pin_memory()
    host = eng.image_to_fen_host(host_in, eng.alloc_outputs(len(sel), full=True, pinned_host=True))
    for k in dev:
        assert torch.equal(dev[k].cpu(), host[k]), f"output '{k}' differs between device and host entry points"


def test_host_path_with_small_groups(golden, monkeypatch):
    """Same with one chunk per geometry group (CVB_GROUP_CHUNKS=1): several groups, both staging slots and a ragged tail."""
    from chessvision import _native
    _, _, imgs = golden
    monkeypatch.setenv("CVB_GROUP_CHUNKS", "1")
    eng = _native.Engine(0, max_batch=8)
    try:
        eng.load_unet(load_checkpoint(WEIGHTS / "best_extractor.pth"))
        eng.load_resnet18(load_checkpoint(WEIGHTS / "best_classifier.pth"))
        sel = imgs[:29]
        dev = eng.image_to_fen(torch.from_numpy(sel).cuda(), eng.alloc_outputs(len(sel), full=True))
        torch.cuda.synchronize()
        host = eng.image_to_fen_host(torch.from_numpy(sel).pin_memory(), eng.alloc_outputs(len(sel), full=True, pinned_host=True))
        for k in dev:
            assert torch.equal(dev[k].cpu(), host[k]), f"output '{k}' differs between device and host entry points"
    finally:
        eng.close()


def test_results_do_not_depend_on_batch_size_or_position(trained_engine, golden):
    """Size-independent property at the bench's chunk size (configs[3] shards 65,536 boards into chunks of 296): 1,184 boards
    drawn with repetition from data/test, in random order, through a context with 296-board chunks -- every board's outputs
    must equal what the 16-board context returns for that image alone-ish (integers and FEN bytes exactly, probabilities and
    logits to fp16 noise at most; measured: identical)."""
    from chessvision import _native
    _, _, imgs = golden
    base = trained_engine.image_to_fen(torch.from_numpy(imgs).cuda(), trained_engine.alloc_outputs(len(imgs), full=True))
    base = {k: v.cpu().numpy() for k, v in base.items()}
    rng = np.random.default_rng(20261018)
    idx = rng.integers(0, len(imgs), 1184)
    eng = _native.Engine(0, max_batch=296)
    try:
        eng.load_unet(load_checkpoint(WEIGHTS / "best_extractor.pth"))
        eng.load_resnet18(load_checkpoint(WEIGHTS / "best_classifier.pth"))
        big_in = torch.from_numpy(imgs[idx]).cuda()
        out = eng.image_to_fen(big_in, eng.alloc_outputs(len(idx), full=True))
        torch.cuda.synchronize()
        again = eng.image_to_fen(big_in, eng.alloc_outputs(len(idx), full=True))     # idempotence of the context (workspaces reused)
        torch.cuda.synchronize()
        for k, v in out.items():
            got, want = v.cpu().numpy(), base[k][idx]
            assert np.array_equal(got, again[k].cpu().numpy()), f"output '{k}' differs between two passes over the same batch"
            if got.dtype.kind == "f":
                assert np.allclose(got, want, atol=2e-3, rtol=0), f"output '{k}': max abs difference {np.abs(got - want).max()}"
            else:
                assert np.array_equal(got, want), f"output '{k}' depends on the batch: {(got != want).reshape(len(idx), -1).any(axis=1).sum()} boards differ"
    finally:
        eng.close()


def test_python_api_process_image(golden):
    """The drop-in class, as the reference's tests use it (tests/test_chessvision.py:45-116): structural checks plus
    agreement with the golden FEN for the reference's own fixture image."""
    import chessvision
    from chessvision import constants
    man, _, imgs = golden
    cvm = chessvision.ChessVision(board_extractor_weights=str(WEIGHTS / "best_extractor.pth"),
                                  classifier_weights=str(WEIGHTS / "best_classifier.pth"), classifier_model_id="resnet18", max_batch=4)
    i = next(k for k, e in enumerate(man["images"]) if e["file"].endswith("1bf29f73-bc30-448b-a894-bd6428754a0c.JPG"))
    res = cvm.process_image(imgs[i])
    be = res.board_extraction
    assert be.binary_mask.dtype == np.uint8 and be.binary_mask.shape == (256, 256)
    assert be.probabilities.dtype == np.float32
    if man["images"][i]["found"]:
        assert be.board_image.shape == constants.BOARD_SIZE and be.quadrangle.shape == (4, 1, 2) and be.quadrangle.dtype == np.float32
        pos = res.position
        assert pos.squares.shape == (64, 64, 64, 1) and pos.model_probabilities.shape == (64, 13)
        assert (pos.original_fen != pos.fen) == bool(pos.validation_fixes)
        for fix in pos.validation_fixes:
            assert fix.square_name in pos.square_names
        sep = cvm.classify_position(be.board_image)
        assert sep.fen == pos.fen
    eb = cvm.extract_board(imgs[i])
    assert np.array_equal(eb.binary_mask, be.binary_mask)
    with pytest.raises(AssertionError):
        cvm.process_image(imgs[i].astype(np.float32))


@pytest.mark.parametrize("h,w", [(600, 800), (1024, 768), (768, 768), (513, 512), (256, 300), (1080, 1920), (256, 256), (1024, 1024), (257, 999),
                                 (100, 100), (128, 128), (255, 255), (200, 300), (300, 200), (64, 512), (37, 211), (255, 256), (1, 1), (250, 1000)])
def test_resize_area_any_size_equals_cv2(trained_engine, h, w):
    """cvb_resize_area against the third-party call of core.py:212 itself, bit for bit, in every code path of INTER_AREA."""
    rng = np.random.default_rng(h * 3 + w)
    imgs = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    got = trained_engine.resize_area(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], cv2.resize(imgs[i], (256, 256), interpolation=cv2.INTER_AREA)), (h, w, i)


@pytest.mark.parametrize("h,w", [(768, 768), (600, 800), (1024, 1024)])
def test_other_input_sizes_equal_the_oracle(trained_engine, golden, h, w):
    """process_image on inputs that are not 512x512 (core.py:168-170 accepts any u8[H,W,3]): the data/test boards enlarged
    to h x w; found flags, quads (+-1 px in the mask frame), boards given identical quads, labels and FEN against the fp32
    oracle pipeline (oracle/pipeline.py, the reference path with its general INTER_AREA resize)."""
    from chessvision._native import fen_strings
    from oracle.pipeline import OraclePipeline
    man, _, imgs = golden
    pick = [i for i, e in enumerate(man["images"]) if e["found"]][:4]
    big = np.stack([cv2.resize(imgs[i], (w, h), interpolation=cv2.INTER_CUBIC) for i in pick])
    out = trained_engine.image_to_fen(torch.from_numpy(big).cuda(), trained_engine.alloc_outputs(len(pick), full=True))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    fens = fen_strings(torch.from_numpy(out["fen"]))
    oracle = OraclePipeline.from_checkpoints(str(WEIGHTS / "best_extractor.pth"), str(WEIGHTS / "best_classifier.pth"))
    for k in range(len(pick)):
        ref = oracle.process_image(big[k])
        assert bool(out["found"][k]) == (ref["quad"] is not None)
        assert np.abs(out["logits"][k] - ref["logits"]).max() <= 0.15
        if ref["quad"] is None:
            continue
        d = int(np.abs(out["quad"][k] - ref["quad"].reshape(4, 2)).max())
        assert d <= 1, d
        if d == 0:
            assert np.array_equal(out["board"][k], ref["board"])
            assert fens[k] == (ref["original_fen"], ref["fen"])


def test_python_api_mixed_sizes(golden):
    """ChessVision.process_images with a list of differently sized images: grouped by size, results in input order."""
    from chessvision import ChessVision
    man, _, imgs = golden
    cv = ChessVision(board_extractor_weights=str(WEIGHTS / "best_extractor.pth"), classifier_weights=str(WEIGHTS / "best_classifier.pth"),
                     classifier_model_id="resnet18", max_batch=8)
    a, b = imgs[0], imgs[1]
    mixed = [a, cv2.resize(b, (640, 480), interpolation=cv2.INTER_AREA), b, cv2.resize(a, (700, 700), interpolation=cv2.INTER_LINEAR)]
    res = cv.process_images(mixed)
    alone = [cv.process_image(m) for m in mixed]
    assert len(res) == 4
    for r, s in zip(res, alone):
        assert (r.position is None) == (s.position is None)
        assert np.array_equal(r.board_extraction.binary_mask, s.board_extraction.binary_mask)
        if r.position is not None:
            assert r.position.fen == s.position.fen
    assert cv.process_images([]) == [] and cv.process_images(np.zeros((0, 512, 512, 3), np.uint8)) == []


@pytest.mark.parametrize("h,w", [(200, 300), (128, 128), (255, 400), (180, 180)])
def test_inputs_smaller_than_256_equal_the_oracle(trained_engine, golden, h, w):
    """core.py:212 resizes ANY input to 256x256 with INTER_AREA; below 256 px OpenCV switches to its bilinear emulation.  The
    data/test boards reduced to h x w, through the device path, against the fp32 oracle pipeline on the same small images."""
    from chessvision._native import fen_strings
    from oracle.pipeline import OraclePipeline
    man, _, imgs = golden
    pick = [i for i, e in enumerate(man["images"]) if e["found"]][:3]
    small = np.stack([cv2.resize(imgs[i], (w, h), interpolation=cv2.INTER_AREA) for i in pick])
    dev_small = torch.from_numpy(small).cuda()
    resized = trained_engine.resize_area(dev_small).cpu().numpy()
    out = trained_engine.image_to_fen(dev_small, trained_engine.alloc_outputs(len(pick), full=True))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    fens = fen_strings(torch.from_numpy(out["fen"]))
    oracle = OraclePipeline.from_checkpoints(str(WEIGHTS / "best_extractor.pth"), str(WEIGHTS / "best_classifier.pth"))
    for k in range(len(pick)):
        assert np.array_equal(resized[k], cv2.resize(small[k], (256, 256), interpolation=cv2.INTER_AREA))
        ref = oracle.process_image(small[k])
        assert bool(out["found"][k]) == (ref["quad"] is not None)
        assert np.abs(out["logits"][k] - ref["logits"]).max() <= 0.15
        if ref["quad"] is None:
            continue
        d = int(np.abs(out["quad"][k] - ref["quad"].reshape(4, 2)).max())
        assert d <= 1, d
        if d == 0:
            assert np.array_equal(out["board"][k], ref["board"])
            # boards recovered from 128..255 px inputs are blurred and the classifier is no longer confident on every square:
            # labels must agree wherever the fp32 oracle's top-2 margin is clear, and the FEN strings whenever all of them are
            top2 = np.sort(ref["probs"], axis=1)[:, -2:]
            clear = (top2[:, 1] - top2[:, 0]) > 0.2
            flips = out["labels"][k] != ref["probs"].argmax(1)
            assert not (flips & clear).any(), (np.flatnonzero(flips), top2[flips])
            assert np.abs(out["probs"][k] - ref["probs"]).max() <= 0.1
            if not flips.any():
                assert fens[k] == (ref["original_fen"], ref["fen"])


def test_process_images_batch_views_and_squares(golden):
    """The batched drop-in call: results equal to one process_image call per board, `squares` equal to extract_squares of the
    board image (core.py:420-439) although it arrives as its own device output, several geometry groups in one call."""
    from chessvision import ChessVision
    man, _, imgs = golden
    cv = ChessVision(board_extractor_weights=str(WEIGHTS / "best_extractor.pth"), classifier_weights=str(WEIGHTS / "best_classifier.pth"),
                     classifier_model_id="resnet18", max_batch=2)          # group = 16 boards: 38 images = three groups
    res = cv.process_images(imgs)
    assert len(res) == len(imgs)
    for i, (r, e) in enumerate(zip(res, man["images"])):
        assert (r.position is not None) == e["found"]
        assert r.board_extraction.probabilities.shape == (256, 256) and r.board_extraction.binary_mask.shape == (256, 256)
        if r.position is None:
            assert r.board_extraction.board_image is None and r.board_extraction.quadrangle is None
            continue
        assert (r.position.original_fen, r.position.fen) == (e["original_fen"], e["fen"]), e["file"]
        assert r.position.squares.shape == (64, 64, 64, 1)
        assert np.array_equal(r.position.squares, ChessVision.extract_squares(r.board_extraction.board_image))
        assert r.board_extraction.quadrangle.shape == (4, 1, 2) and r.board_extraction.quadrangle.dtype == np.float32
        assert (r.position.original_fen != r.position.fen) == bool(r.position.validation_fixes)
    one = cv.process_image(imgs[3])
    assert np.array_equal(one.board_extraction.probabilities, res[3].board_extraction.probabilities)
    assert np.array_equal(one.board_extraction.board_image, res[3].board_extraction.board_image)


def test_utils_drop_in_functions(golden):
    """chessvision.utils of the reference (utils.py:32-132): get_classifier_model + load_model_checkpoint give a module that
    runs the native classifier; extract_perspective equals the cv2 pair it replaces."""
    from chessvision import ChessVision, utils
    from chessvision.modules import NativeUNet
    man, arr, imgs = golden
    model = utils.load_model_checkpoint(utils.get_classifier_model("resnet18"), str(WEIGHTS / "best_classifier.pth"))
    i = next(k for k, e in enumerate(man["images"]) if e["found"])
    board = arr[f"board_{i}"]
    x = torch.from_numpy(ChessVision.extract_squares(board).astype(np.float32)).permute(0, 3, 1, 2) / 255.0
    logits = model.eval()(x)
    assert logits.shape == (64, 13)
    assert np.array_equal(logits.argmax(1).cpu().numpy(), arr[f"labels_{i}"])
    assert np.abs(torch.softmax(logits, 1).cpu().numpy() - arr[f"probs_{i}"]).max() <= 0.05
    unet = utils.load_model_checkpoint(NativeUNet(n_channels=3, n_classes=1), str(WEIGHTS / "best_extractor.pth"))
    xin = torch.from_numpy(cv2.resize(imgs[i], (256, 256), interpolation=cv2.INTER_AREA)[None].astype(np.float32)).permute(0, 3, 1, 2) / 255
    ul = unet(xin)[0, 0].cpu().numpy()
    assert np.abs(ul - arr[f"logits_{i}"].astype(np.float32)).max() <= 0.16
    with pytest.raises(ValueError):
        model(x * 0.37)                                   # not u8/255: refused instead of silently quantised
    quad = np.array(man["images"][i]["quad"], np.float32).reshape(4, 1, 2) * 2
    got = utils.extract_perspective(imgs[i], quad, (512, 512))
    dest = np.array(((0, 0), (512, 0), (512, 512), (0, 512)), np.float32)
    assert np.array_equal(got, cv2.warpPerspective(imgs[i], cv2.getPerspectiveTransform(quad.reshape(4, 2), dest), (512, 512)))



def test_single_board_pass_replays_a_cuda_graph(golden, monkeypatch):
    """process_image on one board: the ~50 launches of the pass are captured into a CUDA graph on the first call and replayed
    afterwards; results are identical to direct launches (CVB_NO_GRAPH=1), for changing inputs, thresholds and orientations."""
    from chessvision import ChessVision
    man, _, imgs = golden
    kw = dict(board_extractor_weights=str(WEIGHTS / "best_extractor.pth"), classifier_weights=str(WEIGHTS / "best_classifier.pth"),
              classifier_model_id="resnet18", max_batch=4)
    cv = ChessVision(**kw)
    got = [cv.process_image(imgs[i], threshold=t, flip=f) for i, t, f in [(0, 0.5, False), (1, 0.5, False), (2, 0.5, True), (3, 0.3, False), (0, 0.5, False)]]
    assert cv._engine.graph_replays() >= 5
    batch = cv.process_images(imgs[:3])          # three boards: still one chunk, another graph
    monkeypatch.setenv("CVB_NO_GRAPH", "1")
    ref = ChessVision(**kw)
    want = [ref.process_image(imgs[i], threshold=t, flip=f) for i, t, f in [(0, 0.5, False), (1, 0.5, False), (2, 0.5, True), (3, 0.3, False), (0, 0.5, False)]]
    assert ref._engine.graph_replays() == 0
    for a, b in zip(got + batch, want + [ref.process_image(imgs[i]) for i in range(3)]):
        assert np.array_equal(a.board_extraction.probabilities, b.board_extraction.probabilities)
        assert np.array_equal(a.board_extraction.binary_mask, b.board_extraction.binary_mask)
        assert (a.position is None) == (b.position is None)
        if a.position is not None:
            assert (a.position.fen, a.position.original_fen) == (b.position.fen, b.position.original_fen)
            assert np.array_equal(a.position.model_probabilities, b.position.model_probabilities)
            assert np.array_equal(a.board_extraction.board_image, b.board_extraction.board_image)
