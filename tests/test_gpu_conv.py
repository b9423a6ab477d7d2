"""Layer-level checks of the fused convolution kernels through the C-ABI test hook: the tcgen05 (f16x3)
kernel and the fp32 SIMT kernel against torch's conv1d on the CPU, including the fused
GroupNorm-apply/SiLU/resample prologue and the bias/residual epilogue."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
MODES = {"fp32": 0, "f16x3": 1, "bf16": 2}


def _run(lib, dev, x_nlc, w, bias, scale, shift, silu, resample, res, math):
    from eegldm import _lib
    B, Tin, Cin = x_nlc.shape
    Cout, _, k = w.shape
    Tc = {0: Tin, 1: Tin // 2, 2: Tin * 2}[resample]
    xd = x_nlc.contiguous().to(dev)
    sd = scale.contiguous().to(dev) if scale is not None else None
    hd = shift.contiguous().to(dev) if shift is not None else None
    rd = res.contiguous().to(dev) if res is not None else None
    out = torch.full((B, Tc, Cout), float("nan"), device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    wc, bc = w.contiguous(), bias.contiguous() if bias is not None else None
    _lib.check(lib.eegldm_test_conv(p(xd), p(sd), p(hd), int(silu), int(resample), p(wc), p(bc), p(rd), B, Tin, Cin, Cout, k,
                                    MODES[math], p(out), None))
    return out.cpu()


def _ref(x_nlc, w, bias, scale, shift, silu, resample, res):
    x = x_nlc.double()
    if scale is not None:
        x = x * scale[:, None, :].double() + shift[:, None, :].double()
    if silu:
        x = F.silu(x)
    x = x.transpose(1, 2)
    if resample == 1:
        x = F.avg_pool1d(x, 2, 2)
    elif resample == 2:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    y = F.conv1d(x, w.double(), bias.double() if bias is not None else None, padding=w.shape[2] // 2).transpose(1, 2)
    if res is not None:
        y = y + res.double()
    return y.float()


CASES = [
    # B, Tin, Cin, Cout, k, affine+silu, resample, residual
    (1, 128, 32, 128, 1, False, 0, False),
    (1, 128, 32, 128, 3, False, 0, False),
    (2, 192, 128, 128, 3, True, 0, True),
    (3, 48, 64, 256, 3, True, 0, False),       # partial last tile (9 segments), several samples per tile
    (2, 64, 256, 128, 3, True, 1, False),      # AvgPool1d(2,2) folded into the load
    (2, 32, 128, 384, 3, True, 2, True),       # nearest x2 folded into the load
    (3, 48, 256, 256, 3, True, 2, False),      # nearest x2 -> 3-tap conv in polyphase form (up-sampling ResBlock, 256-wide tiles)
    (2, 96, 512, 512, 3, True, 2, False),      # the same with two N tiles per phase; tiles straddle samples
    (5, 16, 512, 512, 1, False, 0, True),
    (2, 768, 384, 128, 3, True, 0, False),
    (2, 64, 512, 256, 3, True, 0, True),       # long K -> 256-channel tiles (single TMEM accumulator set)
    (3, 48, 256, 512, 3, False, 0, False),
    (2, 192, 1024, 512, 1, True, 0, True),
]


@pytest.mark.parametrize("math", ["fp32", "f16x3", "bf16"])
@pytest.mark.parametrize("case", CASES)
def test_fused_conv_matches_torch(built_lib, cuda_device, case, math):
    B, Tin, Cin, Cout, k, aff, rs, has_res = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, Tin, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    bias = 0.1 * torch.randn(Cout, generator=g)
    scale = 1 + 0.2 * torch.randn(B, Cin, generator=g) if aff else None
    shift = 0.2 * torch.randn(B, Cin, generator=g) if aff else None
    Tc = {0: Tin, 1: Tin // 2, 2: Tin * 2}[rs]
    res = torch.randn(B, Tc, Cout, generator=g) if has_res else None
    y = _run(built_lib, cuda_device, x, w, bias, scale, shift, aff, rs, res, math)
    ref = _ref(x, w, bias, scale, shift, aff, rs, res)
    assert torch.isfinite(y).all()
    err = (y - ref).abs().max().item()
    if math == "bf16":      # fast mode: single bf16 product, ~2^-8 relative operand error; NOT a parity mode
        assert err < 5e-2, err
    else:                   # fp32 SIMT and f16x3: fp32-level agreement
        torch.testing.assert_close(y, ref, rtol=1e-4, atol=2e-5)


# tile shapes of the tcgen05 kernel: (CTA-pair mask, minimum weight stages for 256-wide tiles, fuse_epilogues mask).  The default is
# (1, 1, 15): CTA pairs on the 256-wide tiles (wherever Cout % 256 == 0), single-CTA MMAs on the 128-wide ones; (x, 10**6) forces 128-wide tiles; pair bit 0 / bit 1 =
# cta_group::2 for the 256- / 128-wide launches; fuse bit 2 (4) = activation operands produced inside the conv kernel (no act_split)
SHAPES = [(3, 1, 3), (3, 10 ** 6, 3), (3, 1, 15), (3, 10 ** 6, 15), (1, 1, 143), (0, 1, 15), (0, 24, 1), (0, 10 ** 6, 5)]
BIG_CASES = [
    (40, 768, 128, 128, 3, True, 0, True),     # 240 M tiles: every persistent CTA walks several tiles (ring wrap, both TMEM sets)
    (37, 192, 512, 512, 3, True, 0, True),     # tiles straddle samples, odd tile count (idle slot in the last pair), 2-4 N tiles
    (33, 192, 512, 1536, 1, False, 0, False),  # the qkv shape
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("case", CASES + BIG_CASES)
def test_conv_tile_shapes(built_lib, cuda_device, case, shape):
    from eegldm import _lib
    B, Tin, Cin, Cout, k, aff, rs, has_res = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, Tin, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    bias = 0.1 * torch.randn(Cout, generator=g)
    scale = 1 + 0.2 * torch.randn(B, Cin, generator=g) if aff else None
    shift = 0.2 * torch.randn(B, Cin, generator=g) if aff else None
    Tc = {0: Tin, 1: Tin // 2, 2: Tin * 2}[rs]
    res = torch.randn(B, Tc, Cout, generator=g) if has_res else None
    ref = _ref(x, w, bias, scale, shift, aff, rs, res)
    _lib.check(built_lib.eegldm_set_conv_tuning(*shape))
    try:
        y = _run(built_lib, cuda_device, x, w, bias, scale, shift, aff, rs, res, "f16x3")
    finally:
        _lib.check(built_lib.eegldm_set_conv_tuning(1, 1, 15))
    assert torch.isfinite(y).all()
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("case", BIG_CASES)
def test_conv_default_shape_big(built_lib, cuda_device, case):
    test_conv_tile_shapes(built_lib, cuda_device, case, (1, 1, 15))     # default: CTA pairs on 256-wide tiles, fused producer


GN_CASES = [(3, 768, 128, 128, 3), (5, 192, 256, 512, 3), (2, 384, 128, 256, 1), (7, 48, 64, 128, 3), (3, 192, 512, 1024, 1)]


@pytest.mark.parametrize("case", GN_CASES)
def test_conv_epilogue_groupnorm_statistics(built_lib, cuda_device, case):
    """The conv epilogue's GroupNorm(32) statistics of its own output (consumer: Normalize, unet.py:71-74) against
    torch's mean / biased variance of that output; data with a large common offset exercises the cancellation-free path."""
    from eegldm import _lib
    B, T, Cin, Cout, k = case
    G = 32
    g = torch.Generator().manual_seed(Cin + Cout + T)
    x = torch.randn(B, T, Cin, generator=g) + 3.0
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    bias = 5.0 + torch.randn(Cout, generator=g)
    xd = x.to(cuda_device)
    out = torch.empty(B, T, Cout, device=cuda_device)
    mean = torch.empty(B, G, device=cuda_device)
    rstd = torch.empty(B, G, device=cuda_device)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(built_lib.eegldm_test_conv_gn(p(xd), p(w.contiguous()), p(bias.contiguous()), B, T, Cin, Cout, k, G, p(out), p(mean),
                                             p(rstd), None))
    y = out.cpu().double().reshape(B, T, G, Cout // G)
    ref_mean = y.mean(dim=(1, 3))
    ref_var = y.var(dim=(1, 3), unbiased=False)
    torch.testing.assert_close(mean.cpu().double(), ref_mean, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rstd.cpu().double(), 1.0 / torch.sqrt(ref_var + 1e-6), rtol=2e-5, atol=1e-6)


ATTN_CASES = [(2, 192, 1, 512), (1, 128, 1, 128), (3, 64, 2, 128), (2, 256, 1, 256), (1, 32, 4, 128),
              # longer than one 256-key score tile: key blocks of 256 / 192 / 256 merged afterwards (the raw-signal DM variant's T = 768)
              (2, 768, 1, 512), (1, 384, 2, 128), (3, 512, 1, 256)]


@pytest.mark.parametrize("case", ATTN_CASES)
@pytest.mark.parametrize("math", ["f16x3", "bf16"])
def test_attention_in_kernel_split(built_lib, cuda_device, case, math):
    """The same attention reading fp32 q, k, v and splitting them in its own producer warps (fuse bit 4; T <= 208, else the
    pre-split path is taken)."""
    from eegldm import _lib
    _lib.check(built_lib.eegldm_set_conv_tuning(0, 1, 29))
    try:
        test_attention_matches_torch(built_lib, cuda_device, case, math)
    finally:
        _lib.check(built_lib.eegldm_set_conv_tuning(1, 1, 15))


@pytest.mark.parametrize("math", ["fp32", "f16x3", "bf16"])
@pytest.mark.parametrize("case", ATTN_CASES)
def test_attention_matches_torch(built_lib, cuda_device, case, math):
    """QKVAttentionLegacy.forward (unet.py:107-125) on the legacy [q;k;v]-per-head channel layout."""
    from eegldm import _lib
    B, T, H, ch = case
    g = torch.Generator().manual_seed(T + ch)
    qkv = torch.randn(B, H * 3 * ch, T, generator=g)                       # reference layout [B, 3C, T]
    q, k, v = qkv.double().reshape(B * H, 3 * ch, T).split(ch, dim=1)
    scale = 1.0 / (ch ** 0.25)
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale), dim=-1)
    ref = torch.einsum("bts,bcs->bct", w, v).reshape(B, H * ch, T).float()
    xd = qkv.transpose(1, 2).contiguous().to(cuda_device)                  # engine layout [B][T][3C]
    out = torch.full((B, T, H * ch), float("nan"), device=cuda_device)
    _lib.check(built_lib.eegldm_test_attention(C.c_void_p(xd.data_ptr()), B, T, H, ch, MODES[math], C.c_void_p(out.data_ptr()), None))
    y = out.cpu().transpose(1, 2)
    assert torch.isfinite(y).all()
    if math == "bf16":
        assert (y - ref).abs().max().item() < 2e-2
    else:
        torch.testing.assert_close(y, ref, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("case", [(3, 192, 1, 512), (2, 64, 2, 128), (5, 256, 1, 256), (2, 32, 4, 128), (3, 96, 1, 128)])
def test_fused_qkv_conv_attention(built_lib, cuda_device, case):
    """qkv conv -> attention with q, k, v handed over as fp16 hi/lo operand images written by the conv epilogue."""
    from eegldm import _lib
    B, T, H, ch = case
    Cc = H * ch
    g = torch.Generator().manual_seed(T + ch + H)
    x = torch.randn(B, T, Cc, generator=g)
    w = torch.randn(3 * Cc, Cc, 1, generator=g) / Cc ** 0.5
    bias = 0.1 * torch.randn(3 * Cc, generator=g)
    qkv = F.conv1d(x.double().transpose(1, 2), w.double(), bias.double())              # [B, 3C, T], legacy head layout
    q, k, v = qkv.reshape(B * H, 3 * ch, T).split(ch, dim=1)
    scale = 1.0 / (ch ** 0.25)
    wgt = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale), dim=-1)
    ref = torch.einsum("bts,bcs->bct", wgt, v).reshape(B, Cc, T).float()
    xd = x.to(cuda_device)
    out = torch.full((B, T, Cc), float("nan"), device=cuda_device)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(built_lib.eegldm_test_qkv_attention(p(xd), p(w.contiguous()), p(bias.contiguous()), B, T, H, ch, p(out), None))
    y = out.cpu().transpose(1, 2)
    assert torch.isfinite(y).all()
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=2e-5)
