"""Parity of the two networks as run by the CUDA path (fp16 activations/weights, fp32 accumulation, BN folded) against
the fp32 PyTorch oracle (oracle/nets.py) on identical weights and inputs.

Stated tolerances (fp16 vs fp32 reference): UNet logits max-abs <= 2% of the logit range + 0.02,
mask IoU >= 0.99 where the oracle mask is non-trivial; classifier probabilities max-abs <= 0.02 and identical argmax
wherever the oracle's top-1 margin exceeds 0.05."""
import numpy as np
import pytest
import torch

import cvb_synth as synth
from oracle import geometry as og
from oracle import nets

pytestmark = pytest.mark.gpu


def randomize_bn(model, gen):
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=gen) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(m.num_features, generator=gen) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.num_features, generator=gen) * 0.1)


@pytest.fixture(scope="module")
def random_engine():
    from chessvision import _native
    gen = torch.Generator().manual_seed(1234)
    torch.manual_seed(1234)
    unet, cls = nets.BoardUNet().eval(), nets.PieceResNet18().eval()
    with torch.no_grad():
        randomize_bn(unet, gen)
        randomize_bn(cls, gen)
        unet.outc.conv.bias.fill_(0.05)
    eng = _native.Engine(0, max_batch=3)   # 3: forces ragged chunks for N=5
    eng.load_unet(unet.state_dict())
    eng.load_resnet18(cls.state_dict())
    yield eng, unet, cls
    eng.close()


def test_unet_stem(random_engine):
    """Fused INTER_AREA + /255 + conv3x3(3->64) + BN + ReLU on tcgen05 vs the fp32 oracle layer; every image border
    (zero padding) and every 32x64 work-unit seam is covered because whole images are compared."""
    eng, unet, _ = random_engine
    rng = np.random.default_rng(15)
    imgs = np.stack([synth.board_image(rng)[0] for _ in range(2)] + [rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)])
    got = eng.unet_stem(torch.from_numpy(imgs).cuda()).float().cpu().permute(0, 3, 1, 2).numpy()
    with torch.no_grad():
        x = torch.from_numpy(np.stack([og.resize_area_half(i) for i in imgs])).float().div(255).permute(0, 3, 1, 2)
        ref = unet.inc.double_conv[:3](x).numpy()
    err = np.abs(got - ref).max()
    print(f"unet stem: max-abs err {err:.5f} (activation max {ref.max():.3f})")
    assert err <= 2e-3 * max(1.0, float(ref.max()))


def test_resnet_stem(random_engine):
    eng, _, cls = random_engine
    rng = np.random.default_rng(16)
    boards = rng.integers(0, 256, (3, 512, 512), dtype=np.uint8)
    boards[2] = (np.kron((np.indices((8, 8)).sum(0) % 2), np.ones((64, 64))) * 255).astype(np.uint8)
    got = eng.resnet_stem(torch.from_numpy(boards).cuda()).float().cpu().permute(0, 3, 1, 2).numpy()
    with torch.no_grad():
        sq = np.stack([og.extract_squares(b) for b in boards]).reshape(-1, 64, 64, 1)
        x = torch.from_numpy(sq).float().permute(0, 3, 1, 2) / 255.0
        ref = torch.nn.functional.max_pool2d(torch.relu(cls.bn1(cls.conv1(x))), 3, 2, 1).numpy()
    err = np.abs(got - ref).max()
    print(f"resnet stem: max-abs err {err:.5f} (activation max {ref.max():.3f})")
    assert got.shape == ref.shape == (192, 64, 16, 16)
    assert err <= 2e-3 * max(1.0, float(ref.max()))


def test_unet_forward_random_weights(random_engine):
    eng, unet, _ = random_engine
    rng = np.random.default_rng(5)
    imgs = np.stack([synth.board_image(rng)[0] for _ in range(5)])
    logits, mask = eng.unet_forward(torch.from_numpy(imgs).cuda(), 0.5)
    logits, mask = logits.cpu().numpy(), mask.cpu().numpy()
    with torch.no_grad():
        x = torch.from_numpy(np.stack([og.resize_area_half(i) for i in imgs])).float().div(255).permute(0, 3, 1, 2)
        ref = unet.double()(x.double())[:, 0].float().numpy()
        unet.float()
    err = np.abs(logits - ref).max()
    span = ref.max() - ref.min()
    print(f"unet random-init: max-abs err {err:.5f}, logit range {span:.4f}")
    assert err <= 0.02 * span + 0.02
    want = np.stack([og.binary_mask(r, 0.5) for r in ref])
    # masks may differ only where the logit is within the error bound of the decision boundary
    diff = mask != want
    assert np.all(np.abs(ref[diff]) <= err + 1e-6)


def test_classifier_random_weights(random_engine):
    eng, _, cls = random_engine
    rng = np.random.default_rng(6)
    boards = rng.integers(0, 256, (5, 512, 512), dtype=np.uint8)
    boards[1] = (np.kron((np.indices((8, 8)).sum(0) % 2), np.ones((64, 64))) * 200 + 20).astype(np.uint8)
    probs, labels, labels_valid, fen = eng.classify(torch.from_numpy(boards).cuda(), False)
    probs = probs.cpu().numpy()
    with torch.no_grad():
        sq = np.stack([og.extract_squares(b) for b in boards]).reshape(-1, 64, 64, 1)
        ref = torch.softmax(cls(torch.from_numpy(sq).float().permute(0, 3, 1, 2) / 255.0), 1).numpy().reshape(5, 64, 13)
    err = np.abs(probs - ref).max()
    print(f"classifier random-init: probabilities max-abs err {err:.5f}")
    assert err <= 0.02
    assert np.allclose(probs.sum(-1), 1.0, atol=1e-5)
    top2 = np.sort(ref, -1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 0.05
    assert np.array_equal(labels.cpu().numpy()[clear], ref.argmax(-1)[clear])
    # FEN assembly and rule 1 agree with the oracle applied to the kernel's own probabilities
    from chessvision._native import fen_strings
    for i, (orig, fixed) in enumerate(fen_strings(fen)):
        f2, o2, _, _, _ = og.position_from_probabilities(probs[i], False)
        assert (orig, fixed) == (o2, f2)
    probs_f, _, _, fen_f = eng.classify(torch.from_numpy(boards).cuda(), True)
    for i, (orig, fixed) in enumerate(fen_strings(fen_f)):
        f2, o2, _, _, _ = og.position_from_probabilities(probs_f[i].cpu().numpy(), True)
        assert (orig, fixed) == (o2, f2)
