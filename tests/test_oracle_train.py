"""The training-step oracle (oracle/train.py) against the reference's own code and hand-computed values.

Where /root/reference exists the unmodified ``chessvision/pytorch_unet/utils/dice_score.py`` and ``unet`` package are
imported and compared; everywhere else known answers pin the loss and the optimizer rule."""
import importlib.util
import math
import sys

import pytest
import torch

from conftest import REFERENCE
from oracle import nets, train as otrain


def test_dice_known_answers():
    t = torch.zeros(2, 1, 4, 4)
    t[0, 0, :2] = 1          # 8 ones
    p = torch.full((2, 1, 4, 4), 0.5)
    # sample 0: inter = 2*4 = 8, sets = 8 + 8 = 16 -> 0.5 ; sample 1: inter 0, sets 8 -> eps/(8+eps)
    want = 0.5 * ((8 + 1e-6) / (16 + 1e-6) + 1e-6 / (8 + 1e-6))
    assert abs(float(otrain.dice_coeff(p, t)) - want) < 1e-7
    assert abs(float(otrain.dice_loss(p, t)) - (1 - want)) < 1e-7
    # all-zero prediction and target: sets_sum == 0 -> replaced by inter (0) -> eps/eps = 1
    z = torch.zeros(1, 1, 4, 4)
    assert float(otrain.dice_coeff(z, z)) == 1.0


def test_loss_is_bce_plus_dice():
    torch.manual_seed(1)
    x, t = torch.randn(3, 1, 8, 8), (torch.rand(3, 1, 8, 8) > 0.5).float()
    bce = (torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))).mean()
    assert torch.allclose(otrain.loss_fn(x, t), bce + otrain.dice_loss(torch.sigmoid(x), t), atol=1e-6)


def test_rmsprop_rule_first_two_steps():
    """RMSprop(alpha .99, eps 1e-8, wd 1e-8, momentum .999) written out by hand."""
    p = torch.nn.Parameter(torch.tensor([0.5, -2.0]))
    opt = torch.optim.RMSprop([p], lr=0.1, weight_decay=otrain.WEIGHT_DECAY, momentum=otrain.MOMENTUM)
    w, sq, buf = [0.5, -2.0], [0.0, 0.0], [0.0, 0.0]
    for g in ([0.3, -0.1], [0.2, 0.4]):
        p.grad = torch.tensor(g)
        opt.step()
        for i in range(2):
            gi = g[i] + 1e-8 * w[i]
            sq[i] = 0.99 * sq[i] + 0.01 * gi * gi
            buf[i] = 0.999 * buf[i] + gi / (math.sqrt(sq[i]) + 1e-8)
            w[i] -= 0.1 * buf[i]
        assert torch.allclose(p.detach(), torch.tensor(w), rtol=1e-5)


@pytest.mark.reference
@pytest.mark.skipif(not (REFERENCE / "chessvision" / "pytorch_unet").exists(), reason="reference checkout not present")
def test_against_unmodified_reference_modules():
    spec = importlib.util.spec_from_file_location("ref_dice", REFERENCE / "chessvision/pytorch_unet/utils/dice_score.py")
    ref_dice = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_dice)
    sys.path.insert(0, str(REFERENCE / "chessvision" / "pytorch_unet"))
    try:
        from unet import UNet as RefUNet
    finally:
        sys.path.pop(0)
    torch.manual_seed(3)
    ref = RefUNet(n_channels=3, n_classes=1)
    mine = nets.BoardUNet()
    mine.load_state_dict(ref.state_dict())          # identical keys
    x = torch.rand(2, 3, 64, 64)
    t = (torch.rand(2, 1, 64, 64) > 0.6).float()
    # reference step body (train_unet.py:309-321), fp32
    ref.train()
    pred = ref(x)
    loss = torch.nn.BCEWithLogitsLoss()(pred, t) + ref_dice.dice_loss(torch.sigmoid(pred), t, multiclass=False, reduce_batch_first=False)
    loss.backward()
    got = otrain.forward_backward(mine, x, t)
    assert torch.allclose(got, loss.detach(), rtol=1e-6, atol=1e-7)
    for (k, a), (_, b) in zip(ref.named_parameters(), mine.named_parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-8), k
    for (k, a), (_, b) in zip(ref.named_buffers(), mine.named_buffers()):
        assert torch.allclose(a.float(), b.float(), rtol=1e-6, atol=1e-8), k
