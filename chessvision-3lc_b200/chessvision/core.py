"""``ChessVision`` — the reference's Python API for image -> FEN (chessvision/core.py:22-567), backed by hand-written
sm_100a CUDA through the C ABI in ``include/chessvision_b200.h``.

Same constructor, methods, static helpers, attributes and error behaviour (``AssertionError`` on bad inputs, ``None``
fields when no board is found).  Added: :meth:`ChessVision.process_images` for batches.  Everything the reference
computes with PyTorch-eager or OpenCV runs on the GPU here; there is no CPU fallback (a missing library or GPU raises).

Input sizes: 512x512x3 (every image under the reference's ``data/test``) takes the fused path; any other size goes through
``cv2.resize(..., INTER_AREA)`` restated bit for bit on the device (true area interpolation when both axes shrink, OpenCV's
fixed-point bilinear emulation as soon as one axis is smaller than 256).  ``process_images`` groups a list of differently
sized images by size.
"""
from __future__ import annotations

import atexit
import logging
import threading
import time
import weakref

import numpy as np
import torch
from numpy.typing import NDArray

from . import _native, constants, utils
from .cv_types import BoardExtractionResult, ChessVisionResult, PositionResult, ValidationFix

logger = logging.getLogger(__name__)

_shared_engine: _native.Engine | None = None
_live_engines: "weakref.WeakSet[_native.Engine]" = weakref.WeakSet()


class QuadCapacityError(RuntimeError):
    """mask->quad met a mask whose contours exceed every capacity of the contour kernels; ``indices`` are the
    positions of those boards in the batch.  The reference's cv2 path has no such limit, so this is raised rather than
    reported as "no board found"."""

    def __init__(self, indices):
        self.indices = list(indices)
        super().__init__(f"mask->quad: contour capacity exceeded for board(s) {self.indices} (QUAD_OVERFLOW, DESIGN.md §4a)")


def _close_shared_engine() -> None:
    global _shared_engine
    if _shared_engine is not None:
        try:
            _shared_engine.close()
        finally:
            _shared_engine = None


def _engine_for_statics() -> _native.Engine:
    """Context used by the static helpers (mask->quad, warp) and by evaluation / quality / decode: the context of a live
    ``ChessVision`` on the current device when there is one, otherwise one small weight-less context shared by all of them
    (closed at interpreter exit, before PyTorch tears CUDA down)."""
    global _shared_engine
    dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
    for eng in list(_live_engines):
        if getattr(eng, "h", None) and eng.device.index == dev:
            return eng
    if _shared_engine is None or _shared_engine.device.index != dev:
        _close_shared_engine()
        _shared_engine = _native.Engine(dev, max_batch=1)
        atexit.register(_close_shared_engine)
    return _shared_engine


class _PinnedOutputPool:
    """Pinned host buffers for the outputs of ``process_images``.

    The library copies each group's results straight into pinned memory (asynchronous DMA at PCIe rate; pageable memory would
    be staged by the driver at a fraction of it) and the result objects hold VIEWS of those buffers, so a buffer set must stay
    untouched for as long as any result that views it is alive.  A set is therefore leased per call: ``weakref.finalize`` on
    the base arrays hands it back only when the last view of the last output has been garbage-collected.  Page-locking is
    slow (hundreds of milliseconds per gigabyte), hence the reuse; at most ``max_cached_bytes`` stay cached."""

    def __init__(self, max_cached_bytes: int = 12 << 30):
        self._free: dict[int, list] = {}
        self._lock = threading.Lock()
        self._cached = 0
        self.max_cached_bytes = max_cached_bytes

    @staticmethod
    def _nbytes(out: dict) -> int:
        return sum(t.numel() * t.element_size() for t in out.values())

    def lease(self, eng: _native.Engine, n: int):
        """-> (dict of pinned torch tensors for the native call, dict of numpy base arrays the results will view)."""
        with self._lock:
            sets = self._free.get(n)
            out = sets.pop() if sets else None
            if out is not None:
                self._cached -= self._nbytes(out)
        if out is None:
            out = eng.alloc_outputs(n, full=True, squares=True, pinned_host=True)
        arrays = {k: v.numpy() for k, v in out.items()}
        pending = [len(arrays)]

        def one_released() -> None:
            with self._lock:
                pending[0] -= 1
                if pending[0] == 0 and self._cached + self._nbytes(out) <= self.max_cached_bytes:
                    self._free.setdefault(n, []).append(out)
                    self._cached += self._nbytes(out)

        for a in arrays.values():
            weakref.finalize(a, one_released)
        return out, arrays


_pinned_pool = _PinnedOutputPool()


def _results_from_outputs(out: dict, lo: int, hi: int, names: list, scale: float, elapsed: float) -> list:
    """``ChessVisionResult`` objects of boards [lo, hi) from batch-sized host arrays (the cvb_outputs layout): every array
    field of a result is a VIEW of row i of those arrays -- no per-board copies."""
    _check_status(out["status"][lo:hi])
    found = out["found"][lo:hi].astype(bool)
    quads = np.array(out["quad"][lo:hi].reshape(-1, 4, 1, 2) * scale, dtype=np.float32)   # _scale_quadrangle, vectorised
    fen_rows = out["fen"][lo:hi]
    labels, fixed = out["labels"][lo:hi], out["labels_valid"][lo:hi]
    changed = (labels != fixed).any(axis=1)
    logits, masks, boards, squares, probs = out["logits"], out["mask"], out["board"], out["squares"], out["probs"]
    results = []
    for k in range(hi - lo):
        i = lo + k
        if found[k]:
            fixes = []
            if changed[k]:
                fixes = [ValidationFix(names[j], constants.LABEL_NAMES[labels[k, j]], constants.LABEL_NAMES[fixed[k, j]], "no_pawns_on_ends")
                         for j in np.flatnonzero(labels[k] != fixed[k])]
            row = fen_rows[k]
            position = PositionResult(fen=row[1].tobytes().split(b"\0", 1)[0].decode(), original_fen=row[0].tobytes().split(b"\0", 1)[0].decode(),
                                      model_probabilities=probs[i], squares=squares[i], square_names=names, validation_fixes=fixes)
            extraction = BoardExtractionResult(probabilities=logits[i], binary_mask=masks[i], quadrangle=quads[k], board_image=boards[i])
        else:
            position = None
            extraction = BoardExtractionResult(probabilities=logits[i], binary_mask=masks[i], quadrangle=None, board_image=None)
        results.append(ChessVisionResult(board_extraction=extraction, position=position, processing_time=elapsed))
    return results


def _check_status(status) -> None:
    bad = np.flatnonzero(np.asarray(status) == 2)
    if bad.size:
        raise QuadCapacityError(bad.tolist())


class _NetHandle:
    """What ``ChessVision.board_extractor`` / ``.classifier`` return: a callable, ``nn.Module``-like view of a network that
    lives inside the native context (core.py:66-82 returns the torch module)."""

    def __init__(self, owner: "ChessVision", kind: str, metadata: dict):
        self._owner, self._kind = owner, kind
        if metadata:
            self.metadata = metadata

    def eval(self):
        return self

    @staticmethod
    def _as_u8(x: torch.Tensor) -> torch.Tensor:
        """The native stems consume 8-bit pixels (the reference feeds ``u8 / 255``, core.py:215-216, 236-237): anything else
        (normalised, augmented, out-of-range data) would be quantised silently, so it is rejected."""
        v = x.detach().float() * 255.0
        r = v.round()
        if not bool(((v - r).abs() <= 1e-3).all()) or float(r.min()) < 0 or float(r.max()) > 255:
            raise ValueError("native network handles take u8/255 images only (the reference's own preprocessing); got values that "
                             "are not multiples of 1/255 in [0, 1]")
        return r.to(torch.uint8)

    def to(self, *_, **__):
        return self

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        eng = self._owner._engine
        if self._kind == "unet":
            # x: f32[N,3,256,256] = u8/255 (core.py:215-216).  The native stem fuses the 2x INTER_AREA reduction, and a
            # 2x pixel replication is its exact inverse ((4a+2)>>2 == a), so the u8 image is recovered and replicated.
            assert x.dim() == 4 and tuple(x.shape[1:]) == (3, 256, 256), "expected f32[N,3,256,256]"
            u8 = self._as_u8(x).permute(0, 2, 3, 1)
            u8 = u8.repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous().to(eng.device)
            logits, _ = eng.unet_forward(u8, 0.5)
            return logits.unsqueeze(1)
        # classifier: x f32[N,1,64,64] = u8/255 (core.py:236-237); N must be a multiple of 64 (whole boards)
        assert x.dim() == 4 and tuple(x.shape[1:]) == (1, 64, 64) and x.shape[0] % 64 == 0, "expected f32[64k,1,64,64]"
        u8 = self._as_u8(x)
        n = x.shape[0] // 64
        board = u8.reshape(n, 8, 8, 64, 64).permute(0, 1, 3, 2, 4).reshape(n, 512, 512).contiguous().to(eng.device)
        probs, _, _, _ = eng.classify(board, False)
        return torch.log(probs.reshape(-1, 13).clamp_min(1e-38))  # logits up to the softmax shift


class ChessVision:
    """Chess position detection from images (drop-in for the reference class)."""

    def __init__(
        self,
        board_extractor_weights: str | None = None,
        board_extractor_model_id: str | None = None,
        classifier_weights: str | None = None,
        classifier_model_id: str | None = None,
        lazy_load: bool = True,
        device_index: int | None = None,
        max_batch: int = 16,
    ):
        logger.info("Initializing ChessVision instance...")
        # construction is lazy like the reference's (core.py:52-64): the GPU is first touched when a model is needed,
        # and that step raises without an sm_100 device (no CPU fallback)
        if device_index is None:
            device_index = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self._device_index = device_index
        self.device = torch.device("cuda", device_index)
        self._max_batch = max_batch
        self._engine_obj: _native.Engine | None = None
        self._board_extractor: _NetHandle | None = None
        self._classifier: _NetHandle | None = None
        # the reference stores None and then crashes in Path(None) (core.py:55, utils.py:49); its own test expects the
        # default path (tests/test_chessvision.py:31) -> resolve None to the default weights
        self._board_extractor_weights = board_extractor_weights or constants.BEST_EXTRACTOR_WEIGHTS
        self._board_extractor_model_id = board_extractor_model_id
        self._classifier_weights = classifier_weights or constants.BEST_CLASSIFIER_WEIGHTS
        self._classifier_model_id = classifier_model_id
        if not lazy_load:
            logger.info("Eager loading models...")
            self._initialize_board_extractor()
            self._initialize_classifier()
            logger.info("Models loaded successfully")

    # ------------------------------------------------------------------------------------------------ models
    @property
    def _engine(self) -> _native.Engine:
        if self._engine_obj is None:
            self._engine_obj = _native.Engine(self._device_index, self._max_batch)
            _live_engines.add(self._engine_obj)
        return self._engine_obj

    @classmethod
    def from_engine(cls, engine: _native.Engine) -> "ChessVision":
        """A ``ChessVision`` on top of an existing native context whose weights are already loaded (no second copy of the
        activation workspaces); used by callers that drive the C ABI directly as well (bench.py)."""
        self = cls(device_index=engine.device.index, max_batch=engine.max_batch)
        self._engine_obj = engine
        _live_engines.add(engine)
        self._board_extractor = _NetHandle(self, "unet", {})
        self._classifier = _NetHandle(self, "resnet18", {})
        self._classifier_model_id = "resnet18"
        return self

    @property
    def board_extractor(self) -> _NetHandle:
        if self._board_extractor is None:
            self._initialize_board_extractor()
        assert self._board_extractor is not None
        return self._board_extractor

    @property
    def classifier(self) -> _NetHandle:
        if self._classifier is None:
            self._initialize_classifier()
        assert self._classifier is not None
        return self._classifier

    def _initialize_board_extractor(self) -> None:
        """core.py:84-106 — only the UNet extractor is part of this path (YOLO variants are out of scope)."""
        logger.info("Initializing board extraction model...")
        assert self._board_extractor_model_id is None, f"Invalid board extractor model ID: {self._board_extractor_model_id}"
        sd, meta = utils.load_state_dict(self._board_extractor_weights)
        self._engine.load_unet(sd)
        self._board_extractor = _NetHandle(self, "unet", meta)

    def _initialize_classifier(self) -> None:
        """core.py:108-150 — ``None`` falls back to resnet18 exactly like the reference does without YOLO installed."""
        logger.info("Initializing piece classifier model...")
        model_id = self._classifier_model_id or "resnet18"
        if model_id == "yolo":
            raise ImportError("YOLO classifiers are not part of the B200 image->FEN path")
        assert model_id == "resnet18", f"classifier architecture '{model_id}' has no native implementation (resnet18 only)"
        sd, meta = utils.load_state_dict(self._classifier_weights)
        self._engine.load_resnet18(sd)
        self._classifier_model_id = model_id
        self._classifier = _NetHandle(self, "resnet18", meta)

    # ------------------------------------------------------------------------------------------------ pipeline
    def process_image(self, image: NDArray[np.uint8], threshold: float = 0.5, flip: bool = False) -> ChessVisionResult:
        """core.py:152-195."""
        assert isinstance(image, np.ndarray), "Image must be a numpy array"
        assert image.dtype == np.uint8, "Image must be uint8"
        assert len(image.shape) == 3, "Image must be 3-dimensional (H,W,C)"
        return self.process_images([image], threshold, flip)[0]

    def process_images(self, images, threshold: float = 0.5, flip: bool = False) -> list[ChessVisionResult]:
        """Batched ``process_image``: u8[N,H,W,3] (or a list of u8[H,W,3], sizes may differ) -> N results, each with every
        field the reference's ``process_image`` returns (core.py:191-195).

        512x512 batches stream through ``cvb_image_to_fen_host_progress``: the library copies chunk after chunk in, runs the
        pipeline and copies each group's results straight into the arrays the result objects will view (no per-board
        copies: logits, mask, board image, squares and probabilities of board i are views of row i of batch-sized arrays).
        The native call runs in a worker thread (it releases the GIL) while this thread turns the groups that have already
        landed into result objects, so the Python half is hidden behind the GPU."""
        start = time.time()
        if not isinstance(images, np.ndarray):
            images = list(images)
            if not images:
                return []
            sizes = {im.shape[:2] for im in images}
            if len(sizes) > 1:   # one native call per image size, results back in the caller's order
                results: list = [None] * len(images)
                for hw in sizes:
                    idx = [i for i, im in enumerate(images) if im.shape[:2] == hw]
                    for i, r in zip(idx, self.process_images(np.stack([images[i] for i in idx]), threshold, flip)):
                        results[i] = r
                return results
            images = np.stack(images)
        if images.shape[0] == 0:
            return []
        batch = images if images.flags["C_CONTIGUOUS"] else np.ascontiguousarray(images)
        assert batch.dtype == np.uint8 and batch.ndim == 4 and batch.shape[3] == 3, "Images must be uint8 [N,H,W,3]"
        self.board_extractor, self.classifier  # noqa: B018  (lazy initialisation)
        eng, n = self._engine, batch.shape[0]
        host_in = torch.from_numpy(batch)
        names = constants.SQUARE_NAMES_FLIPPED if flip else constants.SQUARE_NAMES_NORMAL
        results: list = []
        scale = batch.shape[1] / 256.0

        def build(lo: int, hi: int, arrays) -> None:
            results.extend(_results_from_outputs(arrays, lo, hi, names, scale, (time.time() - start) / n))

        if batch.shape[1:3] == (512, 512):
            out, arrays = _pinned_pool.lease(eng, n)
            if n <= eng.max_batch * 8:
                # one geometry group (the single-image / small-batch latency path): nothing to overlap, no helper thread
                eng.image_to_fen_host(host_in, out, threshold, flip)
                build(0, n, arrays)
                return results
            progress = np.zeros(1, np.int32)
            err: list = []

            def run() -> None:
                try:
                    eng.image_to_fen_host(host_in, out, threshold, flip, progress=progress)
                except BaseException as e:   # noqa: BLE001  (re-raised in the calling thread)
                    err.append(e)

            worker = threading.Thread(target=run, name="cvb-image-to-fen", daemon=True)
            worker.start()
            done = 0
            while done < n:
                ready = int(progress[0])
                if ready > done:
                    build(done, ready, arrays)
                    done = ready
                elif not worker.is_alive():
                    break
                else:
                    time.sleep(0.0005)
            worker.join()
            if err:
                raise err[0]
            ready = int(progress[0])
            if ready > done:
                build(done, ready, arrays)
            del out, arrays   # from here on only the results keep the pinned set alive
        else:
            dev = eng.image_to_fen(host_in.to(eng.device), eng.alloc_outputs(n, full=True, squares=True), threshold, flip)
            build(0, n, {k: v.cpu().numpy() for k, v in dev.items()})
        return results

    def extract_board(self, image: NDArray[np.uint8], threshold: float = 0.5) -> BoardExtractionResult:
        """core.py:197-223."""
        assert isinstance(image, np.ndarray) and image.dtype == np.uint8 and image.ndim == 3
        self.board_extractor  # noqa: B018
        eng = self._engine
        dev_img = torch.from_numpy(np.ascontiguousarray(image[None])).to(eng.device)
        logits, _ = eng.unet_forward(dev_img, threshold)
        return self._logits_to_board(eng, logits, dev_img, image.shape[:2], threshold)

    def classify_position(self, board_image: NDArray[np.uint8], flip: bool = False) -> PositionResult:
        """core.py:225-249."""
        assert isinstance(board_image, np.ndarray) and board_image.dtype == np.uint8 and board_image.shape == (512, 512)
        self.classifier  # noqa: B018
        eng = self._engine
        squares = self.extract_squares(board_image)
        names = constants.SQUARE_NAMES_FLIPPED if flip else constants.SQUARE_NAMES_NORMAL
        probs, _, _, _ = eng.classify(torch.from_numpy(np.ascontiguousarray(board_image[None])).to(eng.device), flip)
        return self.process_position_probabilities(probs[0].cpu().numpy(), names, squares)

    # ------------------------------------------------------------------------------------------------ static helpers
    @staticmethod
    def _logits_to_board(eng, logits_dev, img_dev, hw, threshold) -> BoardExtractionResult:
        mask = eng.mask_from_logits(logits_dev, threshold)
        quad, found, status = eng.mask_to_quad(mask)
        _check_status(status.cpu().numpy())
        logits = logits_dev[0].cpu().numpy()
        if not bool(found[0]):
            logger.info("Failed to extract board from image")
            return BoardExtractionResult(board_image=None, binary_mask=mask[0].cpu().numpy(), quadrangle=None, probabilities=logits)
        board = eng.warp_squares(img_dev, quad, found)
        scaled = ChessVision._scale_quadrangle(quad[0].cpu().numpy().reshape(4, 1, 2), hw)
        return BoardExtractionResult(board_image=board[0].cpu().numpy(), binary_mask=mask[0].cpu().numpy(), quadrangle=scaled,
                                     probabilities=logits)

    @staticmethod
    def process_board_extraction_logits(logits: NDArray[np.float32], orig_image: NDArray[np.uint8], threshold: float) -> BoardExtractionResult:
        """core.py:252-307 — logits may come from anywhere (e.g. scripts/process_new_raw/process_pipeline.py:249)."""
        assert isinstance(logits, np.ndarray), "Logits must be a numpy array"
        assert logits.dtype == np.float32, "Logits must be float32"
        assert isinstance(orig_image, np.ndarray), "Original image must be a numpy array"
        assert orig_image.dtype == np.uint8, "Original image must be uint8"
        assert 0 <= threshold <= 1, "Threshold must be between 0 and 1"
        eng = _engine_for_statics()
        logits_dev = torch.from_numpy(np.ascontiguousarray(logits.reshape(1, 256, 256))).to(eng.device)
        img_dev = torch.from_numpy(np.ascontiguousarray(orig_image[None])).to(eng.device)
        res = ChessVision._logits_to_board(eng, logits_dev, img_dev, orig_image.shape[:2], threshold)
        res.probabilities = logits
        return res

    @staticmethod
    def process_position_probabilities(probabilities: NDArray[np.float32], square_names: list[str], square_crops: NDArray[np.uint8]) -> PositionResult:
        """core.py:310-355 — argmax, FEN, rule 1, FEN again (host strings; the batched path does this on the device)."""
        picks = np.argmax(probabilities, axis=1)
        labels = [constants.LABEL_NAMES[p] for p in picks]
        original_fen = _board_fen(labels, square_names)
        labels, fixes = ChessVision.validate_position(labels, probabilities, square_names)
        return PositionResult(fen=_board_fen(labels, square_names), original_fen=original_fen, model_probabilities=probabilities,
                              squares=square_crops, square_names=square_names, validation_fixes=fixes)

    @staticmethod
    def _find_quadrangle(mask: NDArray[np.uint8]) -> NDArray[np.int32] | None:
        """core.py:358-379 on the device (contours, filter, approxPolyDP, rotation in one kernel)."""
        assert isinstance(mask, np.ndarray) and mask.dtype == np.uint8 and mask.shape == (256, 256)
        eng = _engine_for_statics()
        quad, found, status = eng.mask_to_quad(torch.from_numpy(np.ascontiguousarray(mask[None])).to(eng.device))
        _check_status(status.cpu().numpy())
        return quad[0].cpu().numpy().reshape(4, 1, 2) if bool(found[0]) else None

    @staticmethod
    def _filter_contours(img_shape, contours, min_ratio_bounding: float = 0.6, min_area_percentage: float = 0.35,
                         max_area_percentage: float = 1.0):
        """core.py:382-404 for caller-supplied contours (int32[n,1,2]); shoelace area and bounding box on the host."""
        kept = []
        mask_area = float(img_shape[0] * img_shape[1])
        for contour in contours:
            pts = np.asarray(contour).reshape(-1, 2).astype(np.int64)
            x, y = pts[:, 0], pts[:, 1]
            area = abs(float(np.sum(np.roll(x, 1) * y - np.roll(y, 1) * x)) * 0.5) / mask_area
            if area < min_area_percentage or area > max_area_percentage:
                continue
            w, h = int(x.max() - x.min()) + 1, int(y.max() - y.min()) + 1
            if utils.ratio(h, w) < min_ratio_bounding:
                continue
            kept.append(contour)
        return kept

    @staticmethod
    def _rotate_quadrangle(approx: NDArray[np.int32]) -> NDArray[np.int32]:
        """core.py:407-411."""
        return approx[[3, 0, 1, 2], :, :] if approx[0, 0, 0] < approx[2, 0, 0] else approx

    @staticmethod
    def _scale_quadrangle(approx: NDArray[np.int32], orig_size: tuple[int, int]) -> NDArray[np.float32]:
        """core.py:414-417 — the height scales both axes."""
        return np.array(approx * (orig_size[0] / 256.0), dtype=np.float32)

    @staticmethod
    def extract_squares(board: NDArray[np.uint8]) -> NDArray[np.uint8]:
        """core.py:420-439: u8[512,512] -> u8[64,64,64,1], square i = 8*row + col with row 0 = rank 8."""
        h, w = board.shape
        sh, sw = h // 8, w // 8
        return board.reshape(8, sh, 8, sw).swapaxes(1, 2).reshape(64, sh, sw, 1)

    @staticmethod
    def validate_position(pred_labels: list[str], probabilities: NDArray[np.float32], square_names: list[str]):
        """core.py:442-469 — rule 1 (no pawns on ranks 1/8); mutates and returns ``pred_labels`` like the reference."""
        fixes: list[ValidationFix] = []
        order = np.argsort(probabilities)
        for i, (label, name) in enumerate(zip(pred_labels, square_names)):
            if name in constants.INVALID_PAWN_SQUARES and label in ("P", "p"):
                for alt in order[i][::-1]:
                    alt_piece = constants.LABEL_NAMES[alt]
                    if alt_piece not in ("P", "p"):
                        fixes.append(ValidationFix(square_name=name, original_piece=label, corrected_piece=alt_piece,
                                                   rule_name="no_pawns_on_ends"))
                        pred_labels[i] = alt_piece
                        break
        return pred_labels, fixes


def _board_fen(labels, square_names) -> str:
    """Piece-placement field of a FEN for 64 labels on the given squares ("f" = empty); what python-chess's
    ``BaseBoard.board_fen()`` prints for the reference (core.py:330-336)."""
    at = dict(zip(square_names, labels))
    ranks = []
    for r in "87654321":
        text, gap = "", 0
        for f in "abcdefgh":
            piece = at.get(f + r, "f")
            if piece == "f":
                gap += 1
                continue
            text += (str(gap) if gap else "") + piece
            gap = 0
        ranks.append(text + (str(gap) if gap else ""))
    return "/".join(ranks)
