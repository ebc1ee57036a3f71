"""Non-GEMM U-Net kernels against plain fp32 PyTorch references of the same ops (with the
reference's fp16 rounding points, see oracle/unet.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

from pointdreamer_b200 import _lib
from oracle import unet as ounet

pytestmark = pytest.mark.gpu


def h(x):
    return x.half().float()


def nhwc(x):  # [B,C,H,W] -> [B,H,W,C] fp16 contiguous
    return x.permute(0, 2, 3, 1).contiguous().half()


def nchw(x):  # [B,H,W,C] fp16 -> [B,C,H,W] fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def test_linear_modes(cuda):
    g = torch.Generator().manual_seed(0)
    B, K, N = 5, 256, 1000
    x = torch.randn(B, K, generator=g).to(cuda)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    for mode, fn in [(0, lambda v: v), (1, F.silu)]:
        out = torch.empty(B, N, device=cuda)
        out16 = torch.empty(B, N, device=cuda, dtype=torch.float16)
        _lib.call("pdr_linear", x, W, b, B, K, N, mode, out, out16)
        ref = F.linear(fn(x), W, b)
        assert (out - ref).abs().max().item() < 2e-5
        assert torch.equal(out16, out.half())
    t = torch.tensor([990.0, 370.0, 0.0, 12.0, 555.0], device=cuda)
    out = torch.empty(B, N, device=cuda)
    _lib.call("pdr_linear", t, W, b, B, K, N, 2, out, None)
    ref = F.linear(ounet.timestep_embedding(t.cpu(), K).to(cuda), W, b)
    assert (out - ref).abs().max().item() < 2e-4
    # batch > 8 exercises the chunk loop
    B2 = 19
    x2 = torch.randn(B2, K, generator=g).to(cuda)
    out = torch.empty(B2, N, device=cuda)
    _lib.call("pdr_linear", x2, W, b, B2, K, N, 0, out, None)
    assert (out - F.linear(x2, W, b)).abs().max().item() < 2e-5


@pytest.mark.parametrize("C", [64, 256])
def test_stem_conv(cuda, C):
    g = torch.Generator().manual_seed(1)
    B, H, W = 2, 32, 64
    x = torch.randn(B, 3, H, W, generator=g).to(cuda)
    w = (torch.randn(C, 3, 3, 3, generator=g) / math.sqrt(27)).to(cuda)
    b = torch.randn(C, generator=g).to(cuda)
    wk = w.half().permute(0, 2, 3, 1).reshape(C, 27).contiguous()
    out = torch.empty(B, H, W, C, device=cuda, dtype=torch.float16)
    _lib.call("pdr_stem_conv", x, wk, b, B, H, W, C, out)
    ref = F.conv2d(h(x), h(w), b, padding=1)
    err = (nchw(out) - ref).abs().max().item()
    print("stem max err", err)
    assert err < 4e-3


def _gn_ref(x, gamma, beta, film, silu, resample):
    y = h(F.group_norm(x, 32, gamma, beta, eps=1e-5))
    if film is not None:
        C = x.shape[1]
        scale, shift = h(film[:, :C])[..., None, None], h(film[:, C:])[..., None, None]
        y = h(h(y * h(1 + scale)) + shift)
    if silu:
        y = h(F.silu(y))
    if resample == 1:
        y = h(F.avg_pool2d(y, 2))
    elif resample == 2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    return y


@pytest.mark.parametrize("C1,C2,film,silu,resample", [
    (64, 0, False, True, 0), (256, 0, True, True, 0), (128, 64, False, True, 0),
    (512, 256, False, True, 0), (256, 0, False, True, 1), (192, 0, False, True, 2),
    (256, 0, False, False, 0), (1024, 1024, False, True, 0),
])
def test_group_norm(cuda, C1, C2, film, silu, resample):
    g = torch.Generator().manual_seed(2)
    B, H, W = 3, 16, 16
    C = C1 + C2
    x = (torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3).to(cuda)
    x = h(x)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(C, generator=g)).to(cuda)
    fl = None
    fl16 = None
    stride, off = 0, 0
    if film:
        stride, off = 3 * 2 * C, 2 * C  # embedded in a wider table at an offset
        table = (0.3 * torch.randn(B, stride, generator=g)).to(cuda)
        fl16 = table.half().contiguous()
        fl = fl16[:, off:off + 2 * C].float()
    x1 = nhwc(x[:, :C1])
    x2 = nhwc(x[:, C1:]) if C2 else None
    Ho = H // 2 if resample == 1 else (H * 2 if resample == 2 else H)
    out = torch.empty(B, Ho, Ho, C, device=cuda, dtype=torch.float16)
    ws = torch.empty(B * 64 * 2 * C, device=cuda)
    stats = torch.empty(B * 64, device=cuda)
    _lib.call("pdr_group_norm", x1, x2, B, H, W, C1, C2, gamma, beta, fl16, stride, off,
              1 if silu else 0, resample, ws, stats, out)
    ref = _gn_ref(x, gamma, beta, fl, silu, resample)
    err = (nchw(out) - ref).abs()
    # fp16 rounding boundaries can flip by one ulp
    tol = 2e-3 * ref.abs().clamp(min=1.0)
    print("gn", C1, C2, film, silu, resample, "max err", err.max().item(),
          "frac>1ulp", (err > tol).float().mean().item())
    assert (err > 4 * tol).sum().item() == 0
    assert (err > tol).float().mean().item() < 1e-3
    st = stats.view(B, 32, 2)
    xg = x.view(B, 32, -1)
    assert (st[:, :, 0] - xg.mean(-1)).abs().max().item() < 1e-5
    assert (st[:, :, 1] - 1 / torch.sqrt(xg.var(-1, unbiased=False) + 1e-5)).abs().max().item() < 1e-4


@pytest.mark.parametrize("mode", [1, 2])
def test_resample(cuda, mode):
    g = torch.Generator().manual_seed(3)
    B, H, W, C = 2, 8, 16, 128
    x = h(torch.randn(B, C, H, W, generator=g).to(cuda))
    Ho, Wo = (H // 2, W // 2) if mode == 1 else (H * 2, W * 2)
    out = torch.empty(B, Ho, Wo, C, device=cuda, dtype=torch.float16)
    _lib.call("pdr_resample", nhwc(x), B, H, W, C, mode, out)
    ref = h(F.avg_pool2d(x, 2)) if mode == 1 else F.interpolate(x, scale_factor=2, mode="nearest")
    assert torch.equal(nchw(out), ref)


@pytest.mark.parametrize("T,heads", [(64, 2), (256, 4), (1024, 3)])
def test_attention(cuda, T, heads):
    g = torch.Generator().manual_seed(4)
    B, dh = 2, 64
    C = heads * dh
    qkv = h((torch.randn(B, 3 * C, T, generator=g)).to(cuda))  # [B, 3C, T] like the reference
    out = torch.empty(B, T, C, device=cuda, dtype=torch.float16)
    _lib.call("pdr_attention", qkv.permute(0, 2, 1).contiguous().half(), B, T, heads, out)
    # QKVAttentionLegacy (unet.py:337-354) with fp16 rounding points
    q, k, v = qkv.reshape(B * heads, dh * 3, T).split(dh, dim=1)
    scale = 1 / math.sqrt(math.sqrt(dh))
    w = h(torch.einsum("bct,bcs->bts", h(q * scale), h(k * scale)))
    w = h(torch.softmax(w.float(), dim=-1))
    a = h(torch.einsum("bts,bcs->bct", w, v)).reshape(B, -1, T)
    got = out.float().permute(0, 2, 1)
    err = (got - a).abs()
    print("attention", T, heads, "max err", err.max().item(), "ref max", a.abs().max().item())
    assert err.max().item() < 4e-3 * max(1.0, a.abs().max().item())


@pytest.mark.parametrize("T,heads,B", [(128, 2, 1), (256, 4, 2), (1024, 3, 2)])
def test_attention_prescaled_tcgen05(cuda, T, heads, B):
    """pdr_attention_prescaled: q and k arrive already scaled (what the engine's qkv projection
    emits); T % 128 == 0 runs attention_tc_kernel (tcgen05: S and O in tensor memory, V read as an
    MN-major operand).  Same reference and the same bound as the mma.sync kernel's test."""
    g = torch.Generator().manual_seed(9)
    dh = 64
    C = heads * dh
    scale = 1 / math.sqrt(math.sqrt(dh))
    qkv = h((torch.randn(B, 3 * C, T, generator=g)).to(cuda))  # [B, 3C, T] like the reference
    q, k, v = qkv.reshape(B * heads, dh * 3, T).split(dh, dim=1)
    qs, ks = h(q * scale), h(k * scale)
    pre = torch.cat([qs, ks, v], 1).reshape(B, 3 * C, T)
    out = torch.empty(B, T, C, device=cuda, dtype=torch.float16)
    _lib.call("pdr_attention_prescaled", pre.permute(0, 2, 1).contiguous().half(), B, T, heads, out)
    w = h(torch.einsum("bct,bcs->bts", qs, ks))
    w = h(torch.softmax(w.float(), dim=-1))
    a = h(torch.einsum("bts,bcs->bct", w, v)).reshape(B, -1, T)
    got = out.float().permute(0, 2, 1)
    err = (got - a).abs()
    print("attention (prescaled)", T, heads, "max err", err.max().item(), "ref max", a.abs().max().item())
    assert err.max().item() < 4e-3 * max(1.0, a.abs().max().item())
    # the raw-input entry point (mma.sync kernel) gives the same tensor up to P rounding flips
    out2 = torch.empty_like(out)
    _lib.call("pdr_attention", qkv.permute(0, 2, 1).contiguous().half(), B, T, heads, out2)
    assert (out.float() - out2.float()).abs().max().item() < 4e-3 * max(1.0, a.abs().max().item())


@pytest.mark.parametrize("C,n_out", [(64, 6), (256, 3)])
def test_head(cuda, C, n_out):
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 16, 64
    x = h((torch.randn(B, C, H, W, generator=g) * 2).to(cuda))
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(C, generator=g)).to(cuda)
    w = (torch.randn(6, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(cuda)
    b = torch.randn(6, generator=g).to(cuda)
    out = torch.empty(B, n_out, H, W, device=cuda)
    ws = torch.empty(B * 64 * 2 * C, device=cuda)
    stats = torch.empty(B * 64, device=cuda)
    _lib.call("pdr_unet_head", nhwc(x), gamma, beta, w, b, B, H, W, C, n_out, ws, stats, out)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(F.silu(F.group_norm(x, 32, gamma, beta, eps=1e-5)), w, b, padding=1)[:, :n_out]
    err = (out - ref).abs().max().item()
    print("head max err", err)
    assert err < 1e-4
