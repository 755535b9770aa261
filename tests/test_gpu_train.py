"""UNet training step (BASELINE.json configs[4]) on the B200 against the fp32 oracle (oracle/train.py, pinned to the
reference's train_unet.py / dice_score.py by tests/test_oracle_train.py).

Tolerances (fp16 tensor-core operands + fp16 activations/activation-gradients vs the fp32 reference, stated here):
  * weight gradient of a single conv (identical fp16 inputs): relative L2 error <= 2e-3 (fp32 accumulation order only)
  * loss of one step:                      |delta| <= 5e-3
  * gradients of one step, per tensor:     cosine similarity >= 0.98 for every weight tensor with non-negligible norm,
                                           global relative L2 error <= 0.06
  * optimizer given identical gradients:   parameters equal to torch's clip_grad_norm_ + RMSprop within 1e-5 relative
  * 12-step loss trajectory:               within 0.03 of the oracle's at every step
"""
import numpy as np
import pytest
import torch

from conftest import WEIGHTS, load_checkpoint

pytestmark = pytest.mark.gpu


def _batch(b, seed):
    """Synthetic boards: smooth random colour field + a random convex quad mask that the image correlates with."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(256.0), torch.arange(256.0), indexing="ij")
    imgs, masks = [], []
    for _ in range(b):
        cx, cy = 128 + 30 * (torch.rand(2, generator=g) - 0.5)
        r = 60 + 40 * torch.rand(1, generator=g)
        th = 0.6 * (torch.rand(1, generator=g) - 0.5)
        u = (xx - cx) * torch.cos(th) + (yy - cy) * torch.sin(th)
        v = -(xx - cx) * torch.sin(th) + (yy - cy) * torch.cos(th)
        m = ((u.abs() < r) & (v.abs() < r * (0.8 + 0.2 * torch.rand(1, generator=g)))).float()
        base = torch.rand(3, 8, 8, generator=g)
        img = torch.nn.functional.interpolate(base[None], size=(256, 256), mode="bilinear", align_corners=False)[0]
        checker = (((u / (r / 4)).floor() + (v / (r / 4)).floor()) % 2)
        img = (0.6 * img + 0.4 * m * checker + 0.05 * torch.rand(3, 256, 256, generator=g)).clamp(0, 1)
        imgs.append(img)
        masks.append(m[None])
    return torch.stack(imgs).contiguous(), torch.stack(masks).contiguous()


def _fp32_mode():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("n,h,cout,cin", [(2, 16, 128, 64), (1, 32, 64, 64), (2, 16, 256, 256), (1, 32, 64, 128), (3, 16, 128, 128)])
def test_wgrad3x3_matches_torch(engine, n, h, cout, cin):
    _fp32_mode()
    torch.manual_seed(n * 1000 + h + cout + cin)
    dz = (torch.randn(n, h, h, cout, device="cuda") * 0.5).half()
    x = torch.randn(n, h, h, cin, device="cuda").half()
    got = engine.wgrad3x3_f16(dz, x, scale=0.25).reshape(cout, 3, 3, cin)
    want = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3), dz.float().permute(0, 3, 1, 2), padding=1)
    want = 0.25 * want.permute(0, 2, 3, 1)   # [cout][r][s][cin]
    err = float((got - want).norm() / want.norm())
    print(f"wgrad N={n} H={h} Cout={cout} Cin={cin}: relative L2 error {err:.2e}")
    assert err <= 2e-3


@pytest.fixture(scope="module")
def start_state():
    from oracle import train as otrain
    ck = WEIGHTS / "best_extractor.pth"
    model = otrain.new_model(0)
    if ck.exists():   # start from the trained extractor: realistic activation statistics
        model.load_state_dict(load_checkpoint(ck))
    return {k: v.clone() for k, v in model.state_dict().items()}


def test_one_step_gradients_and_optimizer(start_state):
    from chessvision.training import UNetTrainer
    from oracle import train as otrain
    _fp32_mode()
    B = 2
    images, masks = _batch(B, 7)
    # ---- oracle, fp32 on the GPU
    ref = otrain.new_model(0)
    ref.load_state_dict(start_state)
    ref = ref.cuda()
    ref_loss = float(otrain.forward_backward(ref, images.cuda(), masks.cuda()))
    ref_grads = {k: p.grad.detach().cpu() for k, p in ref.named_parameters()}
    # ---- B200 path
    tr = UNetTrainer(start_state, batch_size=B, learning_rate=1e-4)
    try:
        loss = float(tr.forward_backward(images.cuda(), masks.cuda()).item())
        grads = tr.gradients()
        print(f"loss: b200 {loss:.6f}  oracle {ref_loss:.6f}")
        assert abs(loss - ref_loss) <= 5e-3
        num = den = 0.0
        worst = (1.0, None)
        gmax = max(float(g.norm()) for g in ref_grads.values())
        for k, want in ref_grads.items():
            got = grads[k]
            assert got.shape == want.shape, k
            num += float((got - want).pow(2).sum())
            den += float(want.pow(2).sum())
            if float(want.norm()) > 1e-4 * gmax and want.numel() > 64:
                cos = float(torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0))
                if cos < worst[0]:
                    worst = (cos, k)
        rel = (num / den) ** 0.5
        print(f"gradients: global relative L2 error {rel:.4f}, worst per-tensor cosine {worst[0]:.5f} ({worst[1]})")
        assert rel <= 0.06
        assert worst[0] >= 0.98, worst
        # ---- optimizer: torch's clip + RMSprop applied to OUR gradients must give OUR new parameters
        model = otrain.new_model(0)
        model.load_state_dict(start_state)
        opt = torch.optim.RMSprop(model.parameters(), lr=1e-4, weight_decay=otrain.WEIGHT_DECAY, momentum=otrain.MOMENTUM)
        for k, p in model.named_parameters():
            p.grad = grads[k].clone()
        torch.nn.utils.clip_grad_norm_(model.parameters(), otrain.GRADIENT_CLIPPING)
        opt.step()
        tr.engine.train_optimizer_step(1e-4, 1.0)
        new = tr.state_dict()
        for k, p in model.named_parameters():
            assert torch.allclose(new[k], p.detach(), rtol=1e-5, atol=1e-7), k
        # BatchNorm running statistics moved like nn.BatchNorm2d's (momentum 0.1, unbiased variance)
        for k, v in ref.state_dict().items():
            if "running_mean" in k or "running_var" in k:
                assert torch.allclose(new[k], v.cpu(), rtol=2e-2, atol=2e-3), k
    finally:
        tr.close()


def test_loss_trajectory_tracks_the_oracle(start_state):
    from chessvision.training import UNetTrainer
    from oracle import train as otrain
    _fp32_mode()
    B, steps, lr = 2, 12, 2e-5
    ref = otrain.new_model(0)
    ref.load_state_dict(start_state)
    ref = ref.cuda()
    opt = otrain.make_optimizer(ref, lr)
    tr = UNetTrainer(start_state, batch_size=B, learning_rate=lr)
    try:
        a, b = [], []
        for s in range(steps):
            images, masks = _batch(B, 100 + s % 3)
            a.append(float(otrain.train_step(ref, opt, images.cuda(), masks.cuda())))
            b.append(float(tr.step(images, masks).item()))
        print("oracle losses:", np.round(a, 4).tolist())
        print("b200   losses:", np.round(b, 4).tolist())
        assert max(abs(x - y) for x, y in zip(a, b)) <= 0.03
        assert b[-1] < b[0]
        # the trained weights go straight back into the inference path (cvb_load_unet takes the same tensors)
        sd = tr.state_dict()
        assert set(k for k in start_state if "num_batches" not in k) == set(sd)
    finally:
        tr.close()
