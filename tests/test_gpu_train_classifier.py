"""Piece-classifier training step on the B200 against plain fp32 PyTorch running the same loop as
the reference's scripts/train/train_classifier.py:63-113,218-221 (oracle.nets.PieceResNet18 = timm resnet18(13, in_chans=1),
CrossEntropyLoss, Adam(lr=1e-3), StepLR(4, 0.1)).

Both sides are fp32; the only differences are summation orders.  Tolerances, stated here:
  * logits of a forward pass (train and eval mode):  max |delta| <= 2e-4 * max|logits| + 2e-4
  * loss of one step:                                |delta| <= 1e-4
  * gradients of one step, per tensor:               relative L2 error <= 2e-3 (tensors with negligible norm: absolute 1e-6)
  * BatchNorm running statistics after one step:     max |delta| <= 1e-5 + 1e-4 * |value|
  * Adam given the same gradients (one step):        parameters within 1e-6 absolute of torch.optim.Adam's
  * 10-step loss trajectory incl. a StepLR decay:    within 2e-3 of the oracle's at every step
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import WEIGHTS, load_checkpoint

pytestmark = pytest.mark.gpu


def _fp32_mode():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _batch(b, seed):
    """Synthetic squares: smooth background + a blob whose size / brightness encodes the class."""
    g = torch.Generator().manual_seed(seed)
    target = torch.randint(0, 13, (b,), generator=g)
    yy, xx = torch.meshgrid(torch.arange(64.0), torch.arange(64.0), indexing="ij")
    out = []
    for t in target.tolist():
        base = F.interpolate(torch.rand(1, 1, 4, 4, generator=g), size=(64, 64), mode="bilinear", align_corners=False)[0, 0]
        r = 6 + 1.5 * t
        blob = (((xx - 32) ** 2 + (yy - 32) ** 2) < r * r).float() * (0.2 + 0.06 * t)
        out.append((0.5 * base + blob + 0.03 * torch.rand(64, 64, generator=g)).clamp(0, 1))
    return torch.stack(out)[:, None].contiguous(), target


@pytest.fixture(scope="module")
def start_state():
    from oracle import nets
    torch.manual_seed(0)
    model = nets.PieceResNet18()
    ck = WEIGHTS / "best_classifier.pth"
    if ck.exists():
        model.load_state_dict(load_checkpoint(ck))
    return {k: v.clone() for k, v in model.state_dict().items()}


def _oracle(start_state):
    from oracle import nets
    m = nets.PieceResNet18()
    m.load_state_dict(start_state)
    return m.cuda()


def test_forward_train_and_eval_mode(start_state):
    from chessvision.training import ClassifierTrainer
    _fp32_mode()
    B = 48
    data, target = _batch(B, 3)
    tr = ClassifierTrainer(start_state, batch_size=B)
    ref = _oracle(start_state)
    try:
        for training in (False, True):
            ref.train(training)
            with torch.no_grad():
                want = ref(data.cuda())
                want_loss = F.cross_entropy(want, target.cuda())
            if training:
                logits, loss, correct = tr.engine.cls_train_forward(data.cuda(), target.int().cuda(), training=True)
            else:
                logits, loss, correct = tr.evaluate(data, target)
            torch.cuda.synchronize()
            tol = 2e-4 * float(want.abs().max()) + 2e-4
            err = float((logits - want).abs().max())
            print(f"training={training}: logits max err {err:.2e} (tol {tol:.2e}), loss {float(loss):.6f} vs {float(want_loss):.6f}")
            assert err <= tol
            assert abs(float(loss) - float(want_loss)) <= 1e-4
            assert int(correct) == int((want.argmax(1).cpu() == target).sum())
    finally:
        tr.close()


def test_one_step_gradients_running_stats_and_adam(start_state):
    from chessvision.training import ClassifierTrainer
    _fp32_mode()
    B = 64
    data, target = _batch(B, 11)
    ref = _oracle(start_state).train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    opt.zero_grad()
    loss_ref = F.cross_entropy(ref(data.cuda()), target.cuda())
    loss_ref.backward()
    grads_ref = {k: p.grad.detach().cpu().clone() for k, p in ref.named_parameters()}
    tr = ClassifierTrainer(start_state, batch_size=B)
    try:
        loss, correct = tr.forward_backward(data, target)
        torch.cuda.synchronize()
        print(f"loss {float(loss):.6f} vs oracle {float(loss_ref.detach()):.6f}")
        assert abs(float(loss) - float(loss_ref.detach())) <= 1e-4
        got = tr.gradients()
        worst = 0.0
        for k, want in grads_ref.items():
            g = got[k]
            assert g.shape == want.shape, k
            if float(want.norm()) < 1e-6:
                assert float((g - want).abs().max()) <= 1e-6, k
                continue
            err = float((g - want).norm() / want.norm())
            worst = max(worst, err)
            assert err <= 2e-3, f"{k}: relative L2 error {err:.3e}"
        print(f"worst per-tensor relative L2 gradient error {worst:.2e} over {len(grads_ref)} tensors")
        # Adam on the oracle's side with ITS gradients, on ours with ours
        opt.step()
        tr.engine.cls_train_optimizer_step(tr.learning_rate, 1.0)
        sd = tr.state_dict()
        ref_sd = {k: v.detach().cpu() for k, v in ref.state_dict().items()}
        for k, want in ref_sd.items():
            if not want.is_floating_point():
                assert int(sd[k]) == int(want), k
                continue
            if "running_" in k:
                assert float((sd[k] - want).abs().max()) <= 1e-5 + 1e-4 * float(want.abs().max()), k
            else:
                # the first Adam step moves every parameter by ~lr * sign(g): a gradient that differs in the last bits moves
                # it by the same amount unless |g| is at the level of eps
                close = (sd[k] - want).abs() <= 1e-6
                frac = float(close.float().mean())
                assert frac >= 0.999, f"{k}: only {frac:.4f} of the parameters equal torch's Adam step"
        osd = tr.optimizer_state_dict(names=[k for k, _ in ref.named_parameters()])
        tstate = opt.state_dict()["state"]
        for i, (k, p) in enumerate(ref.named_parameters()):
            want = tstate[i]["exp_avg"].cpu()
            if float(want.norm()) > 1e-7:
                assert float((osd["state"][i]["exp_avg"] - want).norm() / want.norm()) <= 2e-3, k
    finally:
        tr.close()


def test_adam_equals_torch_given_identical_gradients(start_state):
    """Feed torch's gradients through the library's flat buffer: the update must equal torch.optim.Adam's to the last bits."""
    from chessvision.training import ClassifierTrainer
    _fp32_mode()
    B = 16
    data, target = _batch(B, 5)
    tr = ClassifierTrainer(start_state, batch_size=B)
    try:
        tr.forward_backward(data, target)
        torch.cuda.synchronize()
        ours = tr.gradients()
        ref = _oracle(start_state).train()
        with torch.no_grad():
            ref(data.cuda())            # the same running-statistics update
        opt = torch.optim.Adam(ref.parameters(), lr=3e-3)
        for k, p in ref.named_parameters():
            p.grad = ours[k].cuda()
        for _ in range(3):
            opt.step()
            tr.engine.cls_train_optimizer_step(3e-3, 1.0)
        sd = tr.state_dict()
        for k, p in ref.named_parameters():
            err = float((sd[k] - p.detach().cpu()).abs().max())
            assert err <= 1e-6, f"{k}: {err:.2e}"
    finally:
        tr.close()


def test_ten_step_trajectory_with_steplr(start_state):
    from chessvision.training import ClassifierTrainer
    _fp32_mode()
    B = 32
    ref = _oracle(start_state).train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.1)
    tr = ClassifierTrainer(start_state, batch_size=B, learning_rate=1e-3, step_size=2, gamma=0.1)
    try:
        ours, theirs = [], []
        for step in range(10):
            data, target = _batch(B, 100 + step)
            opt.zero_grad()
            l = F.cross_entropy(ref(data.cuda()), target.cuda())
            l.backward()
            opt.step()
            theirs.append(float(l.detach()))
            loss, _ = tr.step(data, target)
            ours.append(float(loss))
            if step % 2 == 1:          # an "epoch" of two batches
                sched.step()
                tr.epoch_end()
        print("ours  ", np.round(ours, 4))
        print("oracle", np.round(theirs, 4))
        assert abs(tr.learning_rate - opt.param_groups[0]["lr"]) <= 1e-12
        assert max(abs(a - b) for a, b in zip(ours, theirs)) <= 2e-3
        # the trained weights drop into the inference path (cvb_load_resnet18) and into the reference's module
        from oracle import nets
        m = nets.PieceResNet18()
        m.load_state_dict(tr.state_dict())
    finally:
        tr.close()
