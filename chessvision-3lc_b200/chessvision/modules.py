"""``torch.nn.Module`` faces of the two networks of the image->FEN path.

The reference hands plain PyTorch modules around (``UNet(3, 1)`` of chessvision/pytorch_unet/unet/unet_model.py:6-48 and
timm's ``resnet18(num_classes=13, in_chans=1)`` of utils.py:32-39): scripts create them, load checkpoints into them with
``utils.load_model_checkpoint`` and call them on ``u8/255`` batches.  The classes here keep that surface -- parameter and
buffer names, shapes, ``state_dict`` / ``load_state_dict``, ``eval()``, ``__call__`` -- but hold no compute: ``forward``
packs the current parameters into a native context (BN folded, fp16 K-major weights) and runs the hand-written sm_100a
kernels.  Inference only; training goes through ``chessvision.training``.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _native


def _conv(cin: int, cout: int, k: int, bias: bool = False) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, k, padding=k // 2, bias=bias)


class _Holder(nn.Module):
    """Parameters only; ``forward`` of the owning network never calls into it."""

    def forward(self, *_):   # pragma: no cover
        raise RuntimeError("parameter container: call the owning native network instead")


def _double_conv(cin: int, cout: int) -> nn.Sequential:
    # unet_parts.py:11-22: conv3x3(no bias) - BN - ReLU - conv3x3(no bias) - BN - ReLU under the name `double_conv`
    return nn.Sequential(_conv(cin, cout, 3), nn.BatchNorm2d(cout), nn.ReLU(inplace=True), _conv(cout, cout, 3), nn.BatchNorm2d(cout),
                         nn.ReLU(inplace=True))


class _NativeNet(nn.Module):
    _kind = ""

    def __init__(self):
        super().__init__()
        self._engine: _native.Engine | None = None
        self._loaded_version = None
        self.max_batch = 16

    def _weights_fingerprint(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict().values())

    def _context(self) -> _native.Engine:
        """(Re)pack the weights when they have changed since the last call (a context holds one packed copy)."""
        version = self._weights_fingerprint()
        if self._engine is None or version != self._loaded_version:
            if self._engine is not None:
                self._engine.close()
            dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
            self._engine = _native.Engine(dev, max_batch=self.max_batch)
            sd = {k: v.detach().float().cpu() for k, v in self.state_dict().items()}
            (self._engine.load_unet if self._kind == "unet" else self._engine.load_resnet18)(sd)
            self._loaded_version = version
        return self._engine

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("native inference module: use chessvision.training for the training step")
        return super().train(False)


class NativeUNet(_NativeNet):
    """``UNet(n_channels=3, n_classes=1, bilinear=False)`` (unet_model.py:6-48): f32[N,3,256,256] = u8/255 -> logits f32[N,1,256,256]."""
    _kind = "unet"

    def __init__(self, n_channels: int = 3, n_classes: int = 1, bilinear: bool = False):
        super().__init__()
        assert (n_channels, n_classes, bilinear) == (3, 1, False), "the native board extractor is UNet(3, 1, bilinear=False)"
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        w = (64, 128, 256, 512, 1024)
        self.inc = _Holder()
        self.inc.double_conv = _double_conv(3, 64)
        for d in range(1, 5):
            down = _Holder()
            inner = _Holder()
            inner.double_conv = _double_conv(w[d - 1], w[d])
            down.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), inner)
            setattr(self, f"down{d}", down)
        for u in range(1, 5):
            up = _Holder()
            up.up = nn.ConvTranspose2d(w[5 - u], w[5 - u] // 2, kernel_size=2, stride=2)
            up.conv = _Holder()
            up.conv.double_conv = _double_conv(w[5 - u], w[4 - u])
            setattr(self, f"up{u}", up)
        self.outc = _Holder()
        self.outc.conv = nn.Conv2d(64, 1, kernel_size=1)
        super().train(False)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .core import _NetHandle
        assert x.dim() == 4 and tuple(x.shape[1:]) == (3, 256, 256), "expected f32[N,3,256,256]"
        eng = self._context()
        u8 = _NetHandle._as_u8(x).permute(0, 2, 3, 1)
        u8 = u8.repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous().to(eng.device)   # the stem's 2x INTER_AREA undoes this exactly
        logits, _ = eng.unet_forward(u8, 0.5)
        return logits.unsqueeze(1)


class _BasicBlock(_Holder):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class NativeResNet18(_NativeNet):
    """timm ``resnet18(num_classes=13, in_chans=1)`` (utils.py:32-39): f32[64k,1,64,64] = u8/255 -> class logits f32[64k,13]
    (whole boards of 64 squares; returned up to the softmax shift, i.e. ``log_softmax``: argmax and softmax are unchanged)."""
    _kind = "resnet18"

    def __init__(self, num_classes: int = 13, in_chans: int = 1):
        super().__init__()
        assert (num_classes, in_chans) == (13, 1), "the native piece classifier is resnet18(num_classes=13, in_chans=1)"
        self.conv1 = nn.Conv2d(1, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        widths = (64, 128, 256, 512)
        for l, c in enumerate(widths):
            cin = widths[max(l - 1, 0)]
            setattr(self, f"layer{l + 1}", nn.Sequential(_BasicBlock(cin, c, 1 if l == 0 else 2), _BasicBlock(c, c, 1)))
        self.fc = nn.Linear(512, num_classes)
        super().train(False)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .core import _NetHandle
        assert x.dim() == 4 and tuple(x.shape[1:]) == (1, 64, 64) and x.shape[0] % 64 == 0, "expected f32[64k,1,64,64] (whole boards)"
        eng = self._context()
        n = x.shape[0] // 64
        board = _NetHandle._as_u8(x).reshape(n, 8, 8, 64, 64).permute(0, 1, 3, 2, 4).reshape(n, 512, 512).contiguous().to(eng.device)
        probs, _, _, _ = eng.classify(board, False)
        return torch.log(probs.reshape(-1, 13).clamp_min(1e-38))
