"""ORACLE (test infrastructure, not product code): fp32 PyTorch restatement of the reference's UNet training step.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
module.  It restates, for ``amp=False`` (plain fp32):

* ``dice_coeff`` / ``dice_loss``  chessvision/pytorch_unet/utils/dice_score.py:5-30 (``multiclass=False``,
  ``reduce_batch_first=False``: per-sample Dice over H,W, mean over the batch, epsilon 1e-6)
* the loss                     scripts/train/train_unet.py:245,309-317  ``BCEWithLogitsLoss()(pred, t) + dice_loss(sigmoid(pred), t)``
* the optimizer                scripts/train/train_unet.py:236-242     ``RMSprop(lr, weight_decay=1e-8, momentum=0.999)``
* one step                     scripts/train/train_unet.py:319-323     zero_grad, backward, ``clip_grad_norm_(params, 1.0)``, step

Pinned by ``tests/test_oracle_train.py``: equal to the reference's own ``dice_score.py`` / ``UNet`` run unmodified from
``/root/reference`` where that checkout exists, and to hand-computed values everywhere.  The reference quirk that
``--amp`` clips *scaled* gradients (no ``unscale_`` before ``clip_grad_norm_``, train_unet.py:320-322 vs upstream
pytorch_unet/train.py:115) is NOT reproduced: the CUDA path follows the fp32 semantics restated here.
"""
from __future__ import annotations

import torch
from torch import nn

from oracle import nets

GRADIENT_CLIPPING = 1.0      # train_unet.py:115 (default of train_model)
WEIGHT_DECAY = 1e-8          # train_unet.py:113
MOMENTUM = 0.999             # train_unet.py:114


def dice_coeff(prob: torch.Tensor, target: torch.Tensor, epsilon: float = 1e-6) -> torch.Tensor:
    assert prob.size() == target.size()
    inter = 2 * (prob * target).sum(dim=(-1, -2))
    sets_sum = prob.sum(dim=(-1, -2)) + target.sum(dim=(-1, -2))
    sets_sum = torch.where(sets_sum == 0, inter, sets_sum)
    return ((inter + epsilon) / (sets_sum + epsilon)).mean()


def dice_loss(prob: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return 1 - dice_coeff(prob, target)


def loss_fn(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return nn.functional.binary_cross_entropy_with_logits(pred, target) + dice_loss(torch.sigmoid(pred), target)


def make_optimizer(model: nn.Module, lr: float) -> torch.optim.Optimizer:
    return torch.optim.RMSprop(model.parameters(), lr=lr, weight_decay=WEIGHT_DECAY, momentum=MOMENTUM, foreach=True)


def forward_backward(model: nn.Module, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
    """loss.backward() of one batch in training mode (BatchNorm batch statistics); returns the loss."""
    model.train()
    for p in model.parameters():
        p.grad = None
    loss = loss_fn(model(images), masks)
    loss.backward()
    return loss.detach()


def train_step(model: nn.Module, optimizer: torch.optim.Optimizer, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
    loss = forward_backward(model, images, masks)
    torch.nn.utils.clip_grad_norm_(model.parameters(), GRADIENT_CLIPPING)
    optimizer.step()
    return loss


def new_model(seed: int = 0) -> nn.Module:
    torch.manual_seed(seed)
    return nets.BoardUNet()
