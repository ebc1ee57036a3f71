"""tcgen05 implicit-GEMM conv vs a plain PyTorch fp32 reference of the same op (GPU)."""
import pytest
import torch
import torch.nn.functional as F

from pointdreamer_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(cuda, B, H, W, C1, C2, Cout, taps, bn, bias=True, residual=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x1 = (torch.randn(B, H, W, C1, generator=g)).half().to(cuda)
    x2 = (torch.randn(B, H, W, C2, generator=g)).half().to(cuda) if C2 else None
    k = 3 if taps == 9 else 1
    C = C1 + C2
    w = (torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5).half().to(cuda)
    b = torch.randn(Cout, generator=g).to(cuda) if bias else None
    res = torch.randn(B, H, W, Cout, generator=g).half().to(cuda) if residual else None
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=cuda)
    # [Cout][tap][C]
    wk = w.permute(0, 2, 3, 1).reshape(Cout, taps * C).contiguous()
    _lib.call("pdr_conv_tc", x1, x2, wk, b, res,
              out, B, H, W, C1, C2, Cout, taps, bn)
    torch.cuda.synchronize()
    xin = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), b, padding=k // 2)
    ref = ref.half()  # reference rounds the conv output to fp16 ...
    if residual:
        ref = (ref.float() + res.float().permute(0, 3, 1, 2)).half()  # ... then adds in fp16
    ref = ref.permute(0, 2, 3, 1).float()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"conv B{B} {H}x{W} C{C1}+{C2}->{Cout} taps{taps} bn{bn}: max_abs_err={err:.3e} "
          f"ref_max={scale:.3f}")
    return err, scale


@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("taps", [1, 9])
def test_conv_basic(cuda, bn, taps):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    err, scale = _run(cuda, 2, 16, 16, 64, 0, 256, taps, bn)
    assert err <= 4e-3 * max(scale, 1.0)


@pytest.mark.parametrize("shape", [
    (1, 8, 8, 128, 0, 256, 9),      # bb=2 > B: batch rows out of bounds
    (3, 8, 8, 64, 64, 128, 9),      # two sources, odd batch
    (2, 32, 32, 128, 64, 256, 9),   # multi-tile spatial, two sources
    (1, 64, 64, 256, 0, 256, 9),    # full 256->256 layer shape
    (2, 16, 16, 192, 0, 384, 1),    # qkv-like 1x1
    (8, 4, 4, 64, 0, 128, 9),       # tiny image, bb=8
])
def test_conv_shapes(cuda, shape):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, C1, C2, Cout, taps = shape
    err, scale = _run(cuda, B, H, W, C1, C2, Cout, taps, 0, residual=True)
    assert err <= 4e-3 * max(scale, 1.0)


def test_conv_many_tiles_persistent(cuda):
    """More tiles than SMs: exercises the persistent loop, TMEM double buffering, ring wrap."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    err, scale = _run(cuda, 4, 128, 128, 64, 0, 256, 9, 256)
    assert err <= 4e-3 * max(scale, 1.0)
    err, scale = _run(cuda, 2, 128, 128, 128, 0, 128, 9, 128)
    assert err <= 4e-3 * max(scale, 1.0)


@pytest.mark.parametrize("shape", [
    (2, 16, 16, 64, 0, 256, 9),      # 4 m-tiles -> 2 pairs, single K segment
    (2, 32, 32, 128, 64, 512, 9),    # two sources, 2 n-tiles
    (1, 64, 64, 256, 0, 256, 9),
    (4, 128, 128, 64, 0, 256, 9),    # many pairs per cluster: ring wrap + TMEM double buffering
    (8, 8, 8, 128, 0, 256, 1),       # bb = 2
])
def test_conv_two_cta(cuda, shape):
    """cta_group::2 variant (bn=512 selects it): SM pairs share the weight tile."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, C1, C2, Cout, taps = shape
    err, scale = _run(cuda, B, H, W, C1, C2, Cout, taps, 512, residual=True)
    assert err <= 4e-3 * max(scale, 1.0)


@pytest.mark.parametrize("shape", [
    # B, H, W, C, S1, S2, Cout, bn
    (2, 16, 16, 128, 64, 0, 128, 128),     # one skip source
    (2, 16, 16, 256, 128, 64, 256, 256),   # concatenated skip input (up path), 1-CTA
    (1, 8, 8, 128, 192, 64, 128, 64),      # bb = 2, BN = 64
    (2, 32, 32, 256, 256, 256, 256, 512),  # 2-CTA kernel
    (4, 64, 64, 64, 64, 64, 256, 0),       # auto tile, many tiles (ring wrap)
])
def test_conv_fused_skip(cuda, shape):
    """out_layers.3 + skip_connection in one GEMM (pdr_conv_tc_skip) vs fp32 torch convs."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, C, S1, S2, Cout, bn = shape
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B, H, W, C, generator=g).half().to(cuda)
    s1 = torch.randn(B, H, W, S1, generator=g).half().to(cuda)
    s2 = torch.randn(B, H, W, S2, generator=g).half().to(cuda) if S2 else None
    w3 = (torch.randn(Cout, C, 3, 3, generator=g) / (9 * C) ** 0.5).half().to(cuda)
    ws = (torch.randn(Cout, S1 + S2, 1, 1, generator=g) / (S1 + S2) ** 0.5).half().to(cuda)
    b = torch.randn(Cout, generator=g).to(cuda)
    wk = torch.cat([w3.permute(0, 2, 3, 1).reshape(Cout, 9 * C), ws.reshape(Cout, S1 + S2)], 1).contiguous()
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=cuda)
    _lib.call("pdr_conv_tc_skip", x, wk, b, s1, s2, out, B, H, W, C, S1, S2, Cout, bn)
    torch.cuda.synchronize()
    sin = s1 if s2 is None else torch.cat([s1, s2], -1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w3.float(), b, padding=1) + \
        F.conv2d(sin.float().permute(0, 3, 1, 2), ws.float())
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"fused skip B{B} {H}x{W} C{C}+skip{S1}+{S2}->{Cout} bn{bn}: max_abs_err={err:.3e} ref_max={scale:.3f}")
    assert err <= 4e-3 * max(scale, 1.0)


def test_unet_engine_split_k_layers_batch_invariant(cuda):
    """The 8x8 layers run split-K (decided for a fixed reference batch): a chain's output bits are
    the same at batch 1, 3 and 8 of the full-size model."""
    from pointdreamer_b200.unet import UNetEngine, random_state_dict, DEFAULT_MODEL_CONFIG
    sd = random_state_dict(DEFAULT_MODEL_CONFIG, 7, cuda)
    eng = UNetEngine(sd, DEFAULT_MODEL_CONFIG, device=cuda)
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(8, 3, 256, 256, generator=g).to(cuda)
    t = torch.full((8,), 310.0, device=cuda)
    y8 = eng(x, t, n_out=3).clone()
    y3 = eng(x[:3].contiguous(), t[:3].contiguous(), n_out=3).clone()
    y1 = eng(x[5:6].contiguous(), t[5:6].contiguous(), n_out=3).clone()
    assert torch.isfinite(y8).all()
    assert torch.equal(y3, y8[:3])
    assert torch.equal(y1[0], y8[5])
