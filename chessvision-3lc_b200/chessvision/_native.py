"""ctypes binding of ``libchessvision_b200.so`` (C ABI in ``include/chessvision_b200.h``).

PyTorch is used here only as the owner of device memory and streams: tensors are allocated with torch and their
``data_ptr()`` is handed to the library.  There is no CPU or PyTorch-eager fallback: if the shared library is missing or
the device is not sm_100, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np
import torch

_LIB_PATH = Path(__file__).resolve().parent.parent / "libchessvision_b200.so"


class NativeError(RuntimeError):
    pass


class _Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class _Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("logits", "mask", "quad", "found", "status", "board", "probs", "labels", "labels_valid", "fen", "squares")]


OUTPUT_FIELDS = tuple(n for n, _ in _Outputs._fields_)


class TrainConfig(C.Structure):
    """cvb_train_config (include/chessvision_b200.h); defaults follow scripts/train/train_unet.py of the reference."""
    _fields_ = [("batch", C.c_int32), ("loss_scale", C.c_float), ("momentum", C.c_float), ("alpha", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("max_grad_norm", C.c_float), ("bn_momentum", C.c_float), ("bn_eps", C.c_float)]

class ClsTrainConfig(C.Structure):
    """cvb_cls_train_config (include/chessvision_b200.h); defaults = torch.optim.Adam / nn.BatchNorm2d defaults
    (scripts/train/train_classifier.py:218-221 of the reference)."""
    _fields_ = [("batch", C.c_int32), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("bn_momentum", C.c_float), ("bn_eps", C.c_float)]


# every symbol include/chessvision_b200.h declares: name -> (restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int, C.c_float
SYMBOLS = {
    "cvb_version": (_I, []),
    "cvb_create": (_P, [_I, _I]),
    "cvb_destroy": (None, [_P]),
    "cvb_last_error": (C.c_char_p, [_P]),
    "cvb_max_batch": (_I, [_P]),
    "cvb_load_unet": (_I, [_P, C.POINTER(_Tensor), _I]),
    "cvb_load_resnet18": (_I, [_P, C.POINTER(_Tensor), _I]),
    "cvb_resize_area_half": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "cvb_unet_forward": (_I, [_P, _P, _I, _F, _P, _P, _P]),
    "cvb_mask_from_logits": (_I, [_P, _P, _I, _F, _P, _P]),
    "cvb_mask_to_quad": (_I, [_P, _P, _I, _P, _P, _P, _P]),
    "cvb_warp_squares": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "cvb_warp_perspective": (_I, [_P, _P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "cvb_classify": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "cvb_image_to_fen": (_I, [_P, _P, _I, _F, _I, C.POINTER(_Outputs), _P]),
    "cvb_image_to_fen_hw": (_I, [_P, _P, _I, _I, _I, _F, _I, C.POINTER(_Outputs), _P]),
    "cvb_unet_forward_hw": (_I, [_P, _P, _I, _I, _I, _F, _P, _P, _P]),
    "cvb_resize_area": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "cvb_image_to_fen_host": (_I, [_P, _P, _I, _F, _I, C.POINTER(_Outputs)]),
    "cvb_image_to_fen_host_progress": (_I, [_P, _P, _I, _F, _I, C.POINTER(_Outputs), _P]),
    "cvb_conv2d_f16": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "cvb_convt2x2_f16": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _I, _I, _P]),
    "cvb_conv3x3_convt2x2_f16": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _I, _I, _P]),
    "cvb_unet_stem": (_I, [_P, _P, _I, _P, _P]),
    "cvb_resnet_stem": (_I, [_P, _P, _I, _P, _P]),
    "cvb_train_default_config": (_I, [C.POINTER(TrainConfig)]),
    "cvb_train_create": (_I, [_P, C.POINTER(_Tensor), _I, C.POINTER(TrainConfig)]),
    "cvb_train_forward_backward": (_I, [_P, _P, _P, _P, _P]),
    "cvb_train_grads": (_I, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "cvb_train_buckets": (_I, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _I]),
    "cvb_train_bucket_wait": (_I, [_P, _I, _P]),
    "cvb_train_optimizer_step": (_I, [_P, _F, _F, _P]),
    "cvb_train_step": (_I, [_P, _P, _P, _F, _P, _P]),
    "cvb_train_export": (_I, [_P, _I, C.POINTER(_Tensor), _I]),
    "cvb_cls_train_default_config": (_I, [C.POINTER(ClsTrainConfig)]),
    "cvb_cls_train_create": (_I, [_P, C.POINTER(_Tensor), _I, C.POINTER(ClsTrainConfig)]),
    "cvb_cls_train_forward": (_I, [_P, _P, _P, _I, _P, _P, _P, _P]),
    "cvb_cls_train_forward_backward": (_I, [_P, _P, _P, _P, _P, _P]),
    "cvb_cls_train_grads": (_I, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "cvb_cls_train_optimizer_step": (_I, [_P, _F, _F, _P]),
    "cvb_cls_train_step": (_I, [_P, _P, _P, _F, _P, _P, _P]),
    "cvb_cls_train_export": (_I, [_P, _I, C.POINTER(_Tensor), _I]),
    "cvb_cls_train_steps": (C.c_int64, [_P]),
    "cvb_wgrad3x3_f16": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P]),
    "cvb_eval_metrics": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "cvb_quality_scores": (_I, [_P, _P, _P, _P, _I, _I, _P, _P]),
    "cvb_jpeg_info": (_I, [_P, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "cvb_jpeg_coefficients": (_I, [_P, C.c_int64, _P, _P]),
    "cvb_decode_jpeg": (_I, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), _I, _I, _I, _P, _P]),
    "cvb_launch_count": (C.c_int64, [_P]),
    "cvb_graph_replays": (C.c_int64, [_P]),
    "cvb_profile": (_I, [_P, _I]),
    "cvb_profile_read": (_I, [_P, C.POINTER(C.c_float), _I]),
}

_lib = None


def load_library() -> C.CDLL:
    """Load the shared library and bind every declared symbol (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise NativeError(f"{_LIB_PATH} is missing: build it with chessvision-3lc_b200/build.sh "
                              "(python -c 'import __graft_entry__ as g; g.build()'); there is no fallback path")
        lib = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _state_dict_array(sd):
    """torch state_dict -> (ctypes array of cvb_tensor, keep-alive list).  Only floating tensors are passed."""
    keep, items = [], []
    for k, v in sd.items():
        if not torch.is_tensor(v) or not v.is_floating_point():
            continue
        a = np.ascontiguousarray(v.detach().cpu().float().numpy())
        keep.append(a)
        t = _Tensor()
        t.name = k.encode()
        t.data = a.ctypes.data
        t.ndim = a.ndim
        for i in range(4):
            t.shape[i] = a.shape[i] if i < a.ndim else 1
        items.append(t)
    arr = (_Tensor * len(items))(*items)
    return arr, keep


class Engine:
    """One context on one GPU.  All tensor arguments are torch CUDA tensors on that GPU (or host arrays for *_host)."""

    def __init__(self, device: int = 0, max_batch: int = 64):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the B200 path has no CPU fallback")
        self.device = torch.device("cuda", device)
        with torch.cuda.device(self.device):   # the caller's current device is left as it was
            torch.zeros(1, device=self.device)  # make sure the primary context exists
        self.h = self.lib.cvb_create(device, max_batch)
        if not self.h:
            raise NativeError("cvb_create failed (see stderr): the library needs an sm_100 GPU and enough free memory")
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "h", None):
            self.lib.cvb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise NativeError(f"{what} failed ({rc}): {self.lib.cvb_last_error(self.h).decode()}")

    # ---- weights
    def load_unet(self, state_dict):
        arr, keep = _state_dict_array(state_dict)
        self._ck(self.lib.cvb_load_unet(self.h, arr, len(arr)), "cvb_load_unet")

    def load_resnet18(self, state_dict):
        arr, keep = _state_dict_array(state_dict)
        self._ck(self.lib.cvb_load_resnet18(self.h, arr, len(arr)), "cvb_load_resnet18")

    # ---- stages (device tensors)
    def resize_area_half(self, img):
        n, h2, w2, _ = img.shape
        out = torch.empty((n, h2 // 2, w2 // 2, 3), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.cvb_resize_area_half(self.h, _ptr(img), n, h2 // 2, w2 // 2, _ptr(out), _stream()), "cvb_resize_area_half")
        return out

    def unet_forward(self, img, threshold=0.5):
        """img u8[N,H,W,3] (any H, W >= 256) -> logits f32[N,256,256], mask u8[N,256,256]."""
        n, h, w = img.shape[:3]
        assert img.dtype == torch.uint8 and img.dim() == 4 and img.shape[3] == 3 and img.is_contiguous()
        logits = torch.empty((n, 256, 256), dtype=torch.float32, device=self.device)
        mask = torch.empty((n, 256, 256), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.cvb_unet_forward_hw(self.h, _ptr(img), n, h, w, threshold, _ptr(logits), _ptr(mask), _stream()), "cvb_unet_forward_hw")
        return logits, mask

    def resize_area(self, img):
        """cv2.resize(img, (256, 256), interpolation=cv2.INTER_AREA) for u8[N,H,W,3], H, W >= 256."""
        n, h, w = img.shape[:3]
        assert img.dtype == torch.uint8 and img.dim() == 4 and img.shape[3] == 3 and img.is_contiguous()
        out = torch.empty((n, 256, 256, 3), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.cvb_resize_area(self.h, _ptr(img), n, h, w, _ptr(out), _stream()), "cvb_resize_area")
        return out

    def mask_from_logits(self, logits, threshold=0.5):
        n = logits.shape[0]
        mask = torch.empty((n, 256, 256), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.cvb_mask_from_logits(self.h, _ptr(logits), n, threshold, _ptr(mask), _stream()), "cvb_mask_from_logits")
        return mask

    def mask_to_quad(self, mask):
        n = mask.shape[0]
        assert mask.dtype == torch.uint8 and tuple(mask.shape[1:]) == (256, 256) and mask.is_contiguous()
        quad = torch.empty((n, 4, 2), dtype=torch.int32, device=self.device)
        found = torch.empty((n,), dtype=torch.uint8, device=self.device)
        status = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._ck(self.lib.cvb_mask_to_quad(self.h, _ptr(mask), n, _ptr(quad), _ptr(found), _ptr(status), _stream()), "cvb_mask_to_quad")
        return quad, found, status

    def warp_squares(self, img, quad, found):
        n, H, W, _ = img.shape
        board = torch.empty((n, 512, 512), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.cvb_warp_squares(self.h, _ptr(img), _ptr(quad), _ptr(found), n, H, W, _ptr(board), _stream()), "cvb_warp_squares")
        return board

    def warp_perspective(self, img, corners, out_size):
        """utils.extract_perspective on the device: img u8[H,W] or u8[H,W,C] (C = 1 or 3), corners f32[4,2] -> u8[h,w(,C)]."""
        assert img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous() and img.dim() in (2, 3)
        H, W = img.shape[:2]
        ch = 1 if img.dim() == 2 else img.shape[2]
        ow, oh = int(out_size[0]), int(out_size[1])
        c = corners.to(device=self.device, dtype=torch.float32).reshape(4, 2).contiguous()
        out = torch.empty((oh, ow) if img.dim() == 2 else (oh, ow, ch), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.cvb_warp_perspective(self.h, _ptr(img), H, W, ch, _ptr(c), ow, oh, _ptr(out), _stream()), "cvb_warp_perspective")
        return out

    def classify(self, board, flip=False):
        n = board.shape[0]
        probs = torch.empty((n, 64, 13), dtype=torch.float32, device=self.device)
        labels = torch.empty((n, 64), dtype=torch.uint8, device=self.device)
        labels_valid = torch.empty((n, 64), dtype=torch.uint8, device=self.device)
        fen = torch.empty((n, 2, 72), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.cvb_classify(self.h, _ptr(board), n, int(flip), _ptr(probs), _ptr(labels), _ptr(labels_valid), _ptr(fen),
                                       _stream()), "cvb_classify")
        return probs, labels, labels_valid, fen

    def alloc_outputs(self, n, full=False, pinned_host=False, squares=False, host=False):
        """Allocate the cvb_outputs buffers (device; pinned host when ``pinned_host``; ordinary host memory when ``host``)."""
        kw = dict(device="cpu", pin_memory=True) if pinned_host else (dict(device="cpu") if host else dict(device=self.device))
        out = {
            "quad": torch.empty((n, 4, 2), dtype=torch.int32, **kw),
            "found": torch.empty((n,), dtype=torch.uint8, **kw),
            "status": torch.empty((n,), dtype=torch.int32, **kw),
            "probs": torch.empty((n, 64, 13), dtype=torch.float32, **kw),
            "labels": torch.empty((n, 64), dtype=torch.uint8, **kw),
            "labels_valid": torch.empty((n, 64), dtype=torch.uint8, **kw),
            "fen": torch.empty((n, 2, 72), dtype=torch.uint8, **kw),
        }
        if full:
            out["logits"] = torch.empty((n, 256, 256), dtype=torch.float32, **kw)
            out["mask"] = torch.empty((n, 256, 256), dtype=torch.uint8, **kw)
            out["board"] = torch.empty((n, 512, 512), dtype=torch.uint8, **kw)
        if squares:
            out["squares"] = torch.empty((n, 64, 64, 64, 1), dtype=torch.uint8, **kw)
        return out

    @staticmethod
    def _outputs_struct(out):
        o = _Outputs()
        for name in OUTPUT_FIELDS:
            t = out.get(name)
            setattr(o, name, None if t is None else t.data_ptr())
        return o

    def image_to_fen(self, img, out, threshold=0.5, flip=False):
        """img u8[N,H,W,3] on the device; 512 x 512 takes the fused path, any other size >= 256 x 256 the general one."""
        n, h, w = img.shape[:3]
        assert img.is_cuda and img.dtype == torch.uint8 and img.dim() == 4 and img.shape[3] == 3 and img.is_contiguous()
        o = self._outputs_struct(out)
        if (h, w) == (512, 512):
            self._ck(self.lib.cvb_image_to_fen(self.h, _ptr(img), n, threshold, int(flip), C.byref(o), _stream()), "cvb_image_to_fen")
        else:
            self._ck(self.lib.cvb_image_to_fen_hw(self.h, _ptr(img), n, h, w, threshold, int(flip), C.byref(o), _stream()), "cvb_image_to_fen_hw")
        return out

    def image_to_fen_host(self, img_host, out_host, threshold=0.5, flip=False, progress=None):
        """img_host: torch CPU uint8 [N,512,512,3] (pinned for overlap) ; out_host: dict of CPU tensors.  ``progress``: optional
        int32 numpy array of one element that the library advances to the number of leading boards whose results have
        landed (read by another thread while this call, which releases the GIL, is still running)."""
        n = img_host.shape[0]
        assert not img_host.is_cuda and img_host.dtype == torch.uint8 and img_host.is_contiguous()
        o = self._outputs_struct(out_host)
        if progress is None:
            self._ck(self.lib.cvb_image_to_fen_host(self.h, C.c_void_p(img_host.data_ptr()), n, threshold, int(flip), C.byref(o)),
                     "cvb_image_to_fen_host")
        else:
            assert progress.dtype == np.int32 and progress.size == 1
            self._ck(self.lib.cvb_image_to_fen_host_progress(self.h, C.c_void_p(img_host.data_ptr()), n, threshold, int(flip), C.byref(o),
                                                             C.c_void_p(progress.ctypes.data)), "cvb_image_to_fen_host_progress")
        return out_host

    # ---- building blocks for parity tests
    def conv2d_f16(self, x, w_packed, bias, ksize, stride, relu, residual=None):
        n, h, w, cin = x.shape
        cout = w_packed.shape[0]
        out = torch.empty((n, h // stride, w // stride, cout), dtype=torch.float16, device=self.device)
        self._ck(self.lib.cvb_conv2d_f16(self.h, _ptr(x), n, h, w, cin, _ptr(w_packed), _ptr(bias), cout, ksize, stride, int(relu),
                                         _ptr(residual), _ptr(out), _stream()), "cvb_conv2d_f16")
        return out

    def convt2x2_f16(self, x, w_packed, bias4, cout, out, out_c_off):
        n, h, w, cin = x.shape
        self._ck(self.lib.cvb_convt2x2_f16(self.h, _ptr(x), n, h, w, cin, _ptr(w_packed), _ptr(bias4), cout, _ptr(out), out.shape[3],
                                           out_c_off, _stream()), "cvb_convt2x2_f16")
        return out

    def conv3x3_convt2x2_f16(self, x, w_packed, bias, w2_packed, bias2, cout2, out, out_c_off):
        n, h, w, cin = x.shape
        self._ck(self.lib.cvb_conv3x3_convt2x2_f16(self.h, _ptr(x), n, h, w, cin, _ptr(w_packed), _ptr(bias), _ptr(w2_packed), _ptr(bias2), cout2,
                                                   _ptr(out), out.shape[3], out_c_off, _stream()), "cvb_conv3x3_convt2x2_f16")
        return out

    def unet_stem(self, img):
        n = img.shape[0]
        assert img.dtype == torch.uint8 and tuple(img.shape[1:]) == (512, 512, 3) and img.is_contiguous()
        out = torch.empty((n, 256, 256, 64), dtype=torch.float16, device=self.device)
        self._ck(self.lib.cvb_unet_stem(self.h, _ptr(img), n, _ptr(out), _stream()), "cvb_unet_stem")
        return out

    def resnet_stem(self, board):
        n = board.shape[0]
        assert board.dtype == torch.uint8 and tuple(board.shape[1:]) == (512, 512) and board.is_contiguous()
        out = torch.empty((n * 64, 16, 16, 64), dtype=torch.float16, device=self.device)
        self._ck(self.lib.cvb_resnet_stem(self.h, _ptr(board), n, _ptr(out), _stream()), "cvb_resnet_stem")
        return out

    # ---- UNet training step (cvb_train_*)
    def train_create(self, state_dict, batch=2, **overrides):
        """model.train() state of UNet(3,1) from a state_dict; overrides: any field of TrainConfig."""
        cfg = TrainConfig()
        self._ck(self.lib.cvb_train_default_config(C.byref(cfg)), "cvb_train_default_config")
        cfg.batch = batch
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown training option '{k}'")
            setattr(cfg, k, v)
        arr, keep = _state_dict_array(state_dict)
        self._ck(self.lib.cvb_train_create(self.h, arr, len(arr), C.byref(cfg)), "cvb_train_create")
        self.train_cfg = cfg
        self._train_shapes = {k: tuple(v.shape) for k, v in state_dict.items() if torch.is_tensor(v) and v.is_floating_point()}
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        return cfg

    def _check_batch(self, images, masks):
        b = self.train_cfg.batch
        assert images.is_cuda and images.dtype == torch.float32 and tuple(images.shape) == (b, 3, 256, 256) and images.is_contiguous()
        assert masks.is_cuda and masks.dtype == torch.float32 and masks.numel() == b * 65536 and masks.is_contiguous()

    def train_forward_backward(self, images, masks):
        """images fp32 [B,3,256,256], masks fp32 [B,1,256,256]; returns the loss as a 1-element device tensor (no sync)."""
        self._check_batch(images, masks)
        self._ck(self.lib.cvb_train_forward_backward(self.h, _ptr(images), _ptr(masks), _ptr(self._loss), _stream()),
                 "cvb_train_forward_backward")
        return self._loss

    def train_grads(self):
        """The library's flat fp32 gradient buffer as a torch tensor aliasing the same device memory (for the all-reduce)."""
        ptr, cnt = C.c_void_p(), C.c_int64()
        self._ck(self.lib.cvb_train_grads(self.h, C.byref(ptr), C.byref(cnt)), "cvb_train_grads")

        class _Alias:
            __cuda_array_interface__ = {"shape": (cnt.value,), "typestr": "<f4", "data": (ptr.value, False), "version": 3, "strides": None}

        return torch.as_tensor(_Alias(), device=self.device)

    def train_buckets(self):
        """[(lo, hi)] ranges of the flat gradient buffer in the order the backward pass completes them."""
        lo, hi = (C.c_int64 * 16)(), (C.c_int64 * 16)()
        n = self.lib.cvb_train_buckets(self.h, lo, hi, 16)
        if n < 0:
            raise NativeError("cvb_train_buckets failed: no trainer")
        return [(int(lo[i]), int(hi[i])) for i in range(n)]

    def train_bucket_wait(self, bucket: int, stream: "torch.cuda.Stream"):
        """Make ``stream`` wait until bucket ``bucket`` of the most recent forward_backward is final."""
        self._ck(self.lib.cvb_train_bucket_wait(self.h, bucket, C.c_void_p(stream.cuda_stream)), "cvb_train_bucket_wait")

    def train_optimizer_step(self, lr, grad_scale=1.0):
        self._ck(self.lib.cvb_train_optimizer_step(self.h, lr, grad_scale, _stream()), "cvb_train_optimizer_step")

    def train_step(self, images, masks, lr):
        self._check_batch(images, masks)
        self._ck(self.lib.cvb_train_step(self.h, _ptr(images), _ptr(masks), lr, _ptr(self._loss), _stream()), "cvb_train_step")
        return self._loss

    def train_export(self, grads=False):
        """Parameters + BatchNorm running statistics (or the gradients) as a CPU state_dict in torch layout."""
        out, items = {}, []
        for k, shape in self._train_shapes.items():
            if grads and ("running_" in k):
                continue
            a = np.zeros(shape, np.float32)
            out[k] = a
            t = _Tensor()
            t.name = k.encode()
            t.data = a.ctypes.data
            t.ndim = a.ndim
            for i in range(4):
                t.shape[i] = a.shape[i] if i < a.ndim else 1
            items.append(t)
        arr = (_Tensor * len(items))(*items)
        self._ck(self.lib.cvb_train_export(self.h, 1 if grads else 0, arr, len(items)), "cvb_train_export")
        return {k: torch.from_numpy(v) for k, v in out.items()}

    # ---- piece-classifier training step (cvb_cls_train_*)
    def cls_train_create(self, state_dict, batch=64, **overrides):
        """model.train() state of resnet18(num_classes=13, in_chans=1) + Adam from a state_dict; overrides: fields of ClsTrainConfig."""
        cfg = ClsTrainConfig()
        self._ck(self.lib.cvb_cls_train_default_config(C.byref(cfg)), "cvb_cls_train_default_config")
        cfg.batch = batch
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown training option '{k}'")
            setattr(cfg, k, v)
        arr, keep = _state_dict_array(state_dict)
        self._ck(self.lib.cvb_cls_train_create(self.h, arr, len(arr), C.byref(cfg)), "cvb_cls_train_create")
        self.cls_cfg = cfg
        self._cls_shapes = {k: tuple(v.shape) for k, v in state_dict.items() if torch.is_tensor(v) and v.is_floating_point()}
        self._cls_loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._cls_correct = torch.zeros(1, dtype=torch.int32, device=self.device)
        return cfg

    def _cls_check(self, data, target):
        b = self.cls_cfg.batch
        assert data.is_cuda and data.dtype == torch.float32 and data.numel() == b * 4096 and data.is_contiguous(), \
            f"data must be a contiguous fp32 CUDA tensor [{b},1,64,64]"
        if target is not None:
            assert target.is_cuda and target.dtype == torch.int32 and target.numel() == b and target.is_contiguous()

    def cls_train_forward(self, data, target=None, training=False):
        """logits fp32 [B,13] (+ loss / correct when target is given) in model.train() or model.eval() state."""
        self._cls_check(data, target)
        logits = torch.empty((self.cls_cfg.batch, 13), dtype=torch.float32, device=self.device)
        self._ck(self.lib.cvb_cls_train_forward(self.h, _ptr(data), _ptr(target) if target is not None else None, 1 if training else 0,
                                                _ptr(self._cls_loss), _ptr(self._cls_correct), _ptr(logits), _stream()), "cvb_cls_train_forward")
        return logits, self._cls_loss, self._cls_correct

    def cls_train_forward_backward(self, data, target):
        self._cls_check(data, target)
        self._ck(self.lib.cvb_cls_train_forward_backward(self.h, _ptr(data), _ptr(target), _ptr(self._cls_loss), _ptr(self._cls_correct), _stream()),
                 "cvb_cls_train_forward_backward")
        return self._cls_loss, self._cls_correct

    def cls_train_grads(self):
        ptr, cnt = C.c_void_p(), C.c_int64()
        self._ck(self.lib.cvb_cls_train_grads(self.h, C.byref(ptr), C.byref(cnt)), "cvb_cls_train_grads")

        class _Alias:
            __cuda_array_interface__ = {"shape": (cnt.value,), "typestr": "<f4", "data": (ptr.value, False), "version": 3, "strides": None}

        return torch.as_tensor(_Alias(), device=self.device)

    def cls_train_optimizer_step(self, lr, grad_scale=1.0):
        self._ck(self.lib.cvb_cls_train_optimizer_step(self.h, lr, grad_scale, _stream()), "cvb_cls_train_optimizer_step")

    def cls_train_step(self, data, target, lr):
        self._cls_check(data, target)
        self._ck(self.lib.cvb_cls_train_step(self.h, _ptr(data), _ptr(target), lr, _ptr(self._cls_loss), _ptr(self._cls_correct), _stream()),
                 "cvb_cls_train_step")
        return self._cls_loss, self._cls_correct

    def cls_train_export(self, what=0):
        """what = 0: parameters + running statistics, 1: gradients, 2 / 3: Adam exp_avg / exp_avg_sq; CPU tensors, torch layout."""
        out, items = {}, []
        for k, shape in self._cls_shapes.items():
            if what != 0 and "running_" in k:
                continue
            a = np.zeros(shape, np.float32)
            out[k] = a
            t = _Tensor()
            t.name = k.encode()
            t.data = a.ctypes.data
            t.ndim = a.ndim
            for i in range(4):
                t.shape[i] = a.shape[i] if i < a.ndim else 1
            items.append(t)
        arr = (_Tensor * len(items))(*items)
        self._ck(self.lib.cvb_cls_train_export(self.h, what, arr, len(items)), "cvb_cls_train_export")
        return {k: torch.from_numpy(v) for k, v in out.items()}

    def cls_train_steps(self) -> int:
        return int(self.lib.cvb_cls_train_steps(self.h))

    def wgrad3x3_f16(self, dz, x, scale=1.0):
        n, h, w, cout = dz.shape
        cin = x.shape[3]
        dw = torch.empty((cout, 9, cin), dtype=torch.float32, device=self.device)
        self._ck(self.lib.cvb_wgrad3x3_f16(self.h, _ptr(dz), _ptr(x), n, h, w, cout, cin, scale, _ptr(dw), _stream()), "cvb_wgrad3x3_f16")
        return dw

    # ---- consumers of the outputs
    def eval_metrics(self, probs, labels, labels_valid, true_labels, k=3, flip=False):
        """probs f32[N,64,13], labels / labels_valid u8[N,64] (or None), true_labels u8[N,64] -> (topk_hits i32[N,k],
        correct i32[N,2]) on the device."""
        n = probs.shape[0]
        assert probs.is_cuda and probs.dtype == torch.float32 and tuple(probs.shape[1:]) == (64, 13) and probs.is_contiguous()
        assert true_labels.is_cuda and true_labels.dtype == torch.uint8 and tuple(true_labels.shape) == (n, 64) and true_labels.is_contiguous()
        hits = torch.empty((n, k), dtype=torch.int32, device=self.device)
        correct = torch.zeros((n, 2), dtype=torch.int32, device=self.device)
        self._ck(self.lib.cvb_eval_metrics(self.h, _ptr(probs), _ptr(labels), _ptr(labels_valid), _ptr(true_labels), n, int(flip), k,
                                           _ptr(hits), _ptr(correct), _stream()), "cvb_eval_metrics")
        return hits, correct

    def quality_scores(self, values, quad=None, found=None):
        """values f32[N,L] (logits), quad f32[N,4,2] or None, found u8[N] or None -> f64[N,4] =
        (quadrangle_regularity, NaN, probability_distribution, probability_confidence)."""
        n = values.shape[0]
        v = values.reshape(n, -1)
        assert v.is_cuda and v.dtype == torch.float32 and v.is_contiguous()
        if quad is not None:
            assert quad.is_cuda and quad.dtype == torch.float32 and quad.numel() == n * 8 and quad.is_contiguous()
        scores = torch.empty((n, 4), dtype=torch.float64, device=self.device)
        self._ck(self.lib.cvb_quality_scores(self.h, _ptr(v), _ptr(quad), _ptr(found), n, v.shape[1], _ptr(scores), _stream()),
                 "cvb_quality_scores")
        return scores

    # ---- JPEG decode front-end
    def decode_jpeg(self, streams):
        """list of ``bytes`` (JPEG files of identical dimensions) -> u8[N,H,W,3] BGR on the device, bit-identical to
        ``cv2.imdecode(buf, cv2.IMREAD_COLOR)``."""
        n = len(streams)
        h, w = jpeg_info(streams[0])
        bufs = [np.frombuffer(b, dtype=np.uint8) for b in streams]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_int64 * n)(*[b.size for b in bufs])
        img = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.cvb_decode_jpeg(self.h, ptrs, sizes, n, h, w, _ptr(img), _stream()), "cvb_decode_jpeg")
        return img

    def launch_count(self) -> int:
        return int(self.lib.cvb_launch_count(self.h))

    def graph_replays(self) -> int:
        return int(self.lib.cvb_graph_replays(self.h))

    def profile(self, enable: bool):
        self._ck(self.lib.cvb_profile(self.h, int(enable)), "cvb_profile")

    def profile_read(self):
        buf = (C.c_float * 7)()
        self._ck(self.lib.cvb_profile_read(self.h, buf, 7), "cvb_profile_read")
        names = ("unet_conv_tc", "unet_aux", "mask_to_quad", "warp", "resnet_stem", "resnet_conv_tc", "head")
        return dict(zip(names, [float(v) for v in buf]))


def jpeg_info(stream: bytes):
    """(height, width) of a JPEG stream; raises NativeError for streams outside the supported subset."""
    lib = load_library()
    buf = np.frombuffer(stream, dtype=np.uint8)
    h, w = C.c_int32(), C.c_int32()
    rc = lib.cvb_jpeg_info(C.c_void_p(buf.ctypes.data), buf.size, C.byref(h), C.byref(w))
    if rc != 0:
        raise NativeError(f"cvb_jpeg_info failed ({rc}): not a baseline 4:2:0 JPEG with dimensions that are multiples of 16")
    return h.value, w.value


def jpeg_coefficients(stream: bytes):
    """Host-side entropy decoding alone: (int16 [H*W*3/2] coefficients, uint16 [3,64] quantisation tables)."""
    lib = load_library()
    h, w = jpeg_info(stream)
    buf = np.frombuffer(stream, dtype=np.uint8)
    coef = np.empty(h * w * 3 // 2, np.int16)
    qt = np.empty((3, 64), np.uint16)
    rc = lib.cvb_jpeg_coefficients(C.c_void_p(buf.ctypes.data), buf.size, C.c_void_p(coef.ctypes.data), C.c_void_p(qt.ctypes.data))
    if rc != 0:
        raise NativeError(f"cvb_jpeg_coefficients failed ({rc})")
    return coef, qt


def fen_strings(fen_tensor):
    """uint8 [N,2,72] -> list of (original_fen, fen)."""
    a = fen_tensor.cpu().numpy()
    out = []
    for i in range(a.shape[0]):
        out.append(tuple(bytes(a[i, j]).split(b"\0", 1)[0].decode() for j in range(2)))
    return out
