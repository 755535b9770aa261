"""Parity of the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) against a plain PyTorch fp32 reference of the same
op on the same fp16-rounded operands.  Tolerance: |out - ref| <= 1e-3*|ref| + 2e-3 (one fp16 rounding of the output,
fp32 accumulation inside)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def pack(w):  # [Cout,Cin,k,k] -> fp16 [Cout,(r*k+s)*Cin+ci]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().half()


def check(out, ref):
    out = out.float()
    err = (out - ref).abs()
    tol = 1e-3 * ref.abs() + 2e-3
    bad = (err > tol).sum().item()
    assert bad == 0, f"{bad} of {err.numel()} elements out of tolerance; max err {err.max().item():.4g} (ref max {ref.abs().max().item():.4g})"


CASES = [
    # N, H, W, Cin, Cout, k, stride, relu, residual
    (2, 16, 16, 64, 64, 3, 1, True, False),      # BLOCK_N 64; row-streaming kernel, two of eight squares present
    (1, 32, 32, 64, 128, 3, 1, True, False),     # BLOCK_N 128
    (3, 16, 16, 128, 256, 3, 1, False, False),   # BLOCK_N 256, 2 K chunks per tap
    (1, 16, 16, 256, 512, 3, 1, True, False),    # 2 N tiles
    (5, 8, 8, 128, 128, 3, 1, True, True),       # tile = 2 images x 8 x 8, ragged batch, residual
    (19, 4, 4, 256, 256, 3, 1, True, True),      # tile = 8 images x 4 x 4, ragged batch
    (70, 2, 2, 512, 512, 3, 1, True, True),      # tile = 32 images x 2 x 2, ragged batch
    (4, 16, 16, 64, 128, 3, 2, True, False),     # stride 2 through the four parity views
    (4, 16, 16, 64, 128, 1, 2, False, False),    # 1x1 stride-2 downsample
    (9, 8, 8, 128, 256, 3, 2, True, False),
    (33, 4, 4, 256, 512, 1, 2, False, False),
    (1, 64, 64, 64, 64, 1, 1, False, False),     # plain GEMM
    (2, 256, 256, 64, 64, 3, 1, True, False),    # full-size UNet level 0 (many tiles per CTA: exercises the pipeline wrap)
    # vertical-reuse kernel (3x3 s1, Cout 64/128, H%16==0, W%8==0): stationary and streamed weights
    (3, 64, 64, 128, 64, 3, 1, True, False),     # Cin 128 -> 64, weights resident (147 KB), 4 stages
    (2, 32, 32, 128, 128, 3, 1, True, True),     # 128 -> 128, weights streamed (3 stages), residual
    (1, 64, 64, 256, 128, 3, 1, False, False),   # 256 -> 128, 4 K chunks
    (1, 128, 128, 64, 128, 3, 1, True, False),   # 64 -> 128 resident
    (7, 16, 8, 64, 64, 3, 1, True, True),        # a single 16x8 tile per image
    # row-streaming kernel (3x3 s1, Cout 64, W%128==0): N=192 MMAs over a ring of TMEM slots (CVB_RS_MODE picks the staging)
    (3, 128, 128, 64, 64, 3, 1, True, False),    # strips of 64 rows, one 128-pixel segment
    (2, 256, 256, 128, 64, 3, 1, True, False),   # Cin 128 (up4.conv0): two K chunks, 144 KB of resident weights
    (5, 64, 128, 64, 64, 3, 1, True, True),      # residual, one strip per image
    (2, 48, 256, 64, 64, 3, 1, False, False),    # strips of 16 rows, two segments
    (150, 16, 128, 64, 64, 3, 1, True, False),   # more strips than SMs, one 16-row strip each
    # ... in its 16x16 form (ResNet layer1): a streamed row is the same image row of eight squares
    (64, 16, 16, 64, 64, 3, 1, True, True),      # one board of squares, residual (BasicBlock conv2)
    (203, 16, 16, 64, 64, 3, 1, True, False),    # ragged: 203 = 25 * 8 + 3 squares
    (1500, 16, 16, 64, 64, 3, 1, False, True),   # more strips than SMs
    # generic kernel as CTA pairs (cta_group::2, BLOCK_N 128 / 256): many units per cluster (ring and both accumulators wrap),
    # an odd number of M tiles (the last pair's second CTA computes a tile beyond the batch), two N tiles, residual
    (37, 32, 32, 512, 512, 3, 1, True, False),
    (41, 8, 8, 256, 256, 3, 1, True, True),
    (300, 4, 4, 256, 512, 3, 1, True, True),
    (75, 8, 8, 128, 128, 3, 1, False, True),
    (21, 16, 16, 128, 256, 3, 2, True, False),
]


@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride,relu,residual", CASES)
def test_conv2d(engine, N, H, W, Cin, Cout, k, stride, relu, residual):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + H + Cin + Cout + k + stride)
    x = (torch.randn(N, H, W, Cin, generator=g) * 0.5).half().cuda()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).half()
    b = torch.randn(Cout, generator=g).cuda()
    res = (torch.randn(N, H // stride, W // stride, Cout, generator=g)).half().cuda() if residual else None
    out = engine.conv2d_f16(x, pack(w.float()).cuda(), b, k, stride, relu, res)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().cuda(), b, stride=stride, padding=k // 2).permute(0, 2, 3, 1)
    if residual:
        ref = ref + res.float()
    if relu:
        ref = ref.relu()
    check(out, ref)


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 128, 64), (1, 16, 16, 1024, 512), (3, 32, 32, 256, 128)])
def test_convt2x2(engine, N, H, W, Cin, Cout):
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(Cin + Cout)
    x = (torch.randn(N, H, W, Cin, generator=g) * 0.5).half().cuda()
    w = (torch.randn(Cin, Cout, 2, 2, generator=g) / Cin ** 0.5).half()
    b = torch.randn(Cout, generator=g)
    wp = w.float().permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous().half().cuda()  # row = (dy*2+dx)*Cout + co
    out = torch.full((N, 2 * H, 2 * W, 2 * Cout), 7.0, dtype=torch.float16, device="cuda")
    engine.convt2x2_f16(x, wp, b.repeat(4).cuda(), Cout, out, Cout)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.float().cuda(), b.cuda(), stride=2).permute(0, 2, 3, 1)
    check(out[..., Cout:], ref)
    assert (out[..., :Cout] == 7.0).all(), "the skip half of the concat buffer must stay untouched"


@pytest.mark.parametrize("N,H,W,Cin", [(1, 16, 16, 128), (3, 32, 32, 128), (2, 128, 128, 128), (5, 16, 16, 64), (150, 16, 16, 128)])
def test_fused_conv3x3_convt2x2_equals_the_two_kernels(engine, N, H, W, Cin):
    """conv_convt_kernel (conv3x3 -> 128 ch + ReLU, kept on chip, then ConvTranspose2d 128 -> 64) against the same two layers
    run as separate launches: the fp16 rounding point and the accumulation order are the same, so the bytes must be equal;
    and against the fp32 PyTorch reference within the usual tolerance."""
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(N * 7 + H + Cin)
    x = (torch.randn(N, H, W, Cin, generator=g) * 0.5).half().cuda()
    w = (torch.randn(128, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).half()
    b = torch.randn(128, generator=g).cuda()
    w2 = (torch.randn(128, 64, 2, 2, generator=g) / 128 ** 0.5).half()
    b2 = torch.randn(64, generator=g)
    wp = pack(w.float()).cuda()
    w2p = w2.float().permute(2, 3, 1, 0).reshape(256, 128).contiguous().half().cuda()
    fused = torch.full((N, 2 * H, 2 * W, 128), 7.0, dtype=torch.float16, device="cuda")
    engine.conv3x3_convt2x2_f16(x, wp, b, w2p, b2.cuda(), 64, fused, 64)
    mid = engine.conv2d_f16(x, wp, b, 3, 1, True)
    two = torch.full((N, 2 * H, 2 * W, 128), 7.0, dtype=torch.float16, device="cuda")
    engine.convt2x2_f16(mid, w2p, b2.repeat(4).cuda(), 64, two, 64)
    torch.cuda.synchronize()
    # same fp16 rounding point; only the fp32 accumulation ORDER of the 3x3 conv may differ from the kernel the unfused layer
    # picks (vertical-reuse: channel chunk outermost; here: tap outermost), which moves a few results by one fp16 ulp
    diff = fused != two
    assert diff.float().mean().item() < 0.01, f"{diff.sum().item()} halfs differ from the unfused pair"
    check(fused[..., 64:], two[..., 64:].float())
    assert (fused[..., :64] == 7.0).all(), "the skip half of the concat buffer must stay untouched"
    ref_mid = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().cuda(), b, padding=1).relu().half().float()
    ref = F.conv_transpose2d(ref_mid, w2.float().cuda(), b2.cuda(), stride=2).permute(0, 2, 3, 1)
    check(fused[..., 64:], ref)
