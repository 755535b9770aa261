"""ORACLE (test infrastructure, not product code): fp32 PyTorch restatements of the two networks on the hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module.  The product path (``chessvision-3lc_b200/``) never does.

* ``BoardUNet``  restates ``chessvision/pytorch_unet/unet/unet_model.py:6-36`` + ``unet_parts.py:8-77``
  (milesial U-Net, ConvTranspose upsampling, n_channels=3, n_classes=1) with *identical state-dict keys*
  (``inc.double_conv.0.weight`` … ``outc.conv.bias``) so reference checkpoints load unchanged.
* ``PieceResNet18`` restates what ``chessvision/utils.py:32-39`` asks timm for:
  ``timm.create_model("resnet18", num_classes=13, in_chans=1)`` (timm==1.0.15, not vendored in the reference).
  Structure from ``notebooks/model-summary.ipynb:31-124``; state-dict keys equal torchvision/timm ``resnet18``.

Pinned by ``tests/test_oracle_vs_reference.py`` (runs only where /root/reference exists) and by the golden
vectors in ``tests/golden/``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def _two_convs(cin: int, cout: int) -> nn.Sequential:
    # indices 0..5 = conv, bn, relu, conv, bn, relu  (unet_parts.py:15-22)
    layers = []
    for a, b in ((cin, cout), (cout, cout)):
        layers += [nn.Conv2d(a, b, 3, padding=1, bias=False), nn.BatchNorm2d(b), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


class _Block(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.double_conv = _two_convs(cin, cout)

    def forward(self, x):
        return self.double_conv(x)


class _Down(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _Block(cin, cout))  # unet_parts.py:33-36

    def forward(self, x):
        return self.maxpool_conv(x)


class _Up(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, 2, stride=2)  # unet_parts.py:53
        self.conv = _Block(cin, cout)

    def forward(self, deep, skip):
        deep = self.up(deep)
        dh, dw = skip.shape[2] - deep.shape[2], skip.shape[3] - deep.shape[3]
        if dh or dw:  # unet_parts.py:59-63 (no-op at 256x256)
            deep = F.pad(deep, [dw // 2, dw - dw // 2, dh // 2, dh - dh // 2])
        return self.conv(torch.cat([skip, deep], 1))  # skip first, upsampled second (unet_parts.py:67)


class _Head(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        return self.conv(x)


class BoardUNet(nn.Module):
    """UNet(3 -> 1), 31,037,633 parameters."""

    WIDTHS = (64, 128, 256, 512, 1024)

    def __init__(self, n_channels: int = 3, n_classes: int = 1):
        super().__init__()
        w = self.WIDTHS
        self.inc = _Block(n_channels, w[0])
        for i in range(4):
            setattr(self, f"down{i + 1}", _Down(w[i], w[i + 1]))
        for i in range(4):
            setattr(self, f"up{i + 1}", _Up(w[4 - i], w[3 - i]))
        self.outc = _Head(w[0], n_classes)

    def forward(self, x):
        skips = [self.inc(x)]
        for i in range(4):
            skips.append(getattr(self, f"down{i + 1}")(skips[-1]))
        y = skips.pop()
        for i in range(4):
            y = getattr(self, f"up{i + 1}")(y, skips.pop())
        return self.outc(y)


class _Basic(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idn = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + idn)


class PieceResNet18(nn.Module):
    """ResNet-18 (1 -> 13), 11,176,909 parameters."""

    def __init__(self, num_classes: int = 13, in_chans: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_chans, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        c = 64
        for li, width in enumerate((64, 128, 256, 512), start=1):
            blocks = [_Basic(c, width, 1 if li == 1 else 2), _Basic(width, width, 1)]
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
            c = width
        self.fc = nn.Linear(512, num_classes)

    def features(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.max_pool2d(x, 3, 2, 1)
        for li in range(1, 5):
            x = getattr(self, f"layer{li}")(x)
        return x

    def forward(self, x):
        return self.fc(self.features(x).mean((2, 3)))


def load_state(model: nn.Module, path: str) -> nn.Module:
    """Checkpoint layouts accepted by the reference loader (utils.py:55-86)."""
    blob = torch.load(path, map_location="cpu")
    if isinstance(blob, dict):
        for key in ("model_state_dict", "state_dict", "model"):
            if key in blob:
                blob = blob[key]
                break
    model.load_state_dict({k: v.float() for k, v in blob.items()})
    return model.eval()
