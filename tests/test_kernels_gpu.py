"""GPU parity tests of the individual sm_100a kernels, called through the C ABI (ctypes), against
the CPU oracle / plain torch fp32 on the same seeded inputs."""
import pytest
import torch
import torch.nn.functional as F

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from rangeldm_b200 import _lib
    _lib.lib()
    return _lib


def cl(x):      # (B,C,W,H) -> (B,W,H,C)
    return x.permute(0, 2, 3, 1).contiguous()


def ref_layout(x):   # (B,W,H,C) -> (B,C,W,H)
    return x.permute(0, 3, 1, 2).contiguous()


def pack_w(w, split=False):  # (Cout,Cin,kW,kH) -> [planes][tap][Cout][Cin] fp16
    co, ci, k0, k1 = w.shape
    t = w.permute(2, 3, 0, 1).reshape(k0 * k1, co, ci)
    hi = t.half()
    if not split:
        return hi.contiguous()
    return torch.cat([hi, (t - hi.float()).half()], 0).contiguous()


def padw(x, circular=True):
    """(B,W,H,C) -> the W-padded operand layout (B,W+2,H,C): wrap (or zero) halo columns."""
    if circular:
        return torch.cat([x[:, -1:], x, x[:, :1]], dim=1).contiguous()
    z = torch.zeros_like(x[:, :1])
    return torch.cat([z, x, z], dim=1).contiguous()


def unpadw(x):
    return x[:, 1:-1].contiguous()


def split_half(x):
    hi = x.half()
    return hi, (x - hi.float()).half()


def oracle_conv(x, w, b, stride, pad_lo, ks, circular=True):
    """x (B,C,W,H) fp32; reference semantics `ldm/utils.py:46-49` (+ asymmetric `:109-111`)."""
    if ks == 1:
        return F.conv2d(x, w, b, stride)
    if pad_lo == 1:
        if circular:
            x = F.pad(x, (0, 0, 1, 1), mode="circular")
            x = F.pad(x, (1, 1, 0, 0))
        else:
            x = F.pad(x, (1, 1, 1, 1))
    else:
        x = F.pad(x, (0, 0, 0, 1), mode="circular") if circular else F.pad(x, (0, 0, 0, 1))
        x = F.pad(x, (0, 1, 0, 0))
    return F.conv2d(x, w, b, stride)


CONV_CASES = [
    # B, W, H, Cin, Cout, ks, stride, pad_lo, split_k, circular
    (2, 16, 16, 64, 64, 3, 1, 1, 1, 1),
    (1, 32, 8, 128, 128, 3, 1, 1, 1, 1),
    (1, 32, 8, 128, 128, 3, 1, 1, 2, 1),      # cluster split-K x2 (DSMEM reduction)
    (2, 16, 8, 256, 128, 3, 1, 1, 8, 1),      # cluster split-K x8
    (4, 32, 2, 256, 256, 3, 1, 1, 4, 1),      # 64-pixel images: one tile spans two images
    (2, 16, 4, 128, 256, 3, 1, 1, 0, 1),      # Ho = 4: column boxes smaller than a swizzle atom
    (3, 8, 2, 256, 128, 3, 1, 1, 0, 1),       # Ho = 2, partial last tile (M = 48)
    (1, 4, 64, 64, 64, 3, 1, 1, 1, 1),        # Ho = 64 (decoder top level geometry)
    (1, 32, 16, 128, 128, 3, 2, 1, 1, 1),     # UNet Downsample2D(padding=1)
    (1, 32, 16, 64, 64, 3, 2, 0, 1, 1),       # VAE-encoder Downsample2D(padding=0), asymmetric pad
    (2, 16, 8, 128, 384, 1, 1, 0, 1, 1),      # attention qkv projection as a 1x1 "conv"
    (1, 16, 8, 384, 128, 3, 1, 1, 0, 1),      # skip-concat width
    (1, 16, 8, 64, 64, 3, 1, 1, 1, 0),        # zero padding on W as well (no surgery)
    (8, 256, 16, 128, 128, 3, 1, 1, 0, 1),    # C3 top-level layer at full size (persistent kernel: 256 tiles)
    (8, 256, 16, 64, 256, 3, 1, 1, 0, 1),     # persistent, two N tiles per M tile
    (2, 512, 32, 64, 64, 3, 1, 1, 0, 1),      # persistent, BLOCK_N = 64 (decoder geometry), 256 tiles
    (5, 64, 32, 64, 128, 1, 1, 0, 0, 1),      # persistent 1x1, 80 tiles -> not persistent; sanity
    (8, 512, 8, 64, 128, 3, 1, 1, 0, 1),      # persistent halo windows at Ho = 8 (nuScenes latent geometry)
    (3, 256, 16, 128, 256, 3, 1, 1, 0, 0),    # persistent halo windows, odd batch, zero-padded W
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_tc_matches_cuda_core_restatement_and_oracle(L, case):
    B, W, H, Cin, Cout, ks, stride, pad_lo, split, circ = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(B, Cin, W, H, generator=g)
    w = torch.randn(Cout, Cin, ks, ks, generator=g) / (Cin * ks * ks) ** 0.5
    b = torch.randn(Cout, generator=g)
    temb = torch.randn(B, Cout + 8, generator=g)
    Wo, Ho = W // stride, H // stride
    res = torch.randn(B, Cout, Wo, Ho, generator=g)
    xh = padw(cl(x).half(), bool(circ)).cuda()
    wt = pack_w(w).cuda()
    bd, td, rd = b.cuda(), temb.cuda(), cl(res).cuda()
    out_tc = torch.full((B, Wo, Ho, Cout), float("nan"), device="cuda")
    out_rf = torch.full((B, Wo, Ho, Cout), float("nan"), device="cuda")
    L.call("rldm_conv_tc", L.ptr(xh), None, L.ptr(wt), L.ptr(bd), L.ptr(td), Cout + 8, L.ptr(rd), L.ptr(out_tc),
           B, W, H, Cin, Cout, ks, stride, pad_lo, circ, split, None)
    L.call("rldm_conv_ref", L.ptr(xh), None, L.ptr(wt), L.ptr(bd), L.ptr(td), Cout + 8, L.ptr(rd), L.ptr(out_rf),
           B, W, H, Cin, Cout, ks, stride, pad_lo, circ)
    torch.cuda.synchronize()
    # same fp16 operands on both sides: only the fp32 summation order differs
    assert relerr(out_tc, out_rf) < 2e-5
    # oracle with fp16-rounded operands (exact products in fp32)
    y = oracle_conv(unpadw(xh).float().cpu().permute(0, 3, 1, 2), wt.float().cpu().reshape(ks, ks, Cout, Cin).permute(2, 3, 0, 1),
                    b, stride, pad_lo, ks, bool(circ)) + temb[:, :Cout, None, None] + res
    assert relerr(ref_layout(out_tc.cpu()), y) < 2e-5
    # and against the un-rounded fp32 oracle within the north-star tolerance
    y32 = oracle_conv(x, w, b, stride, pad_lo, ks, bool(circ)) + temb[:, :Cout, None, None] + res
    assert relerr(ref_layout(out_tc.cpu()), y32) < 1e-3
    # split-fp16 ("fp16x3", the engine default): hi+lo operands, 3 MMAs per K step -> ~fp32 accuracy
    xh2, xl2 = split_half(cl(x))
    xh2, xl2, wt2 = padw(xh2, bool(circ)).cuda(), padw(xl2, bool(circ)).cuda(), pack_w(w, split=True).cuda()
    out3 = torch.full((B, Wo, Ho, Cout), float("nan"), device="cuda")
    out3r = torch.full((B, Wo, Ho, Cout), float("nan"), device="cuda")
    fused = Wo * Ho >= 64
    stats = torch.zeros(B, Cout // 2, 2, dtype=torch.float64, device="cuda")
    L.call("rldm_conv_tc", L.ptr(xh2), L.ptr(xl2), L.ptr(wt2), L.ptr(bd), L.ptr(td), Cout + 8, L.ptr(rd), L.ptr(out3),
           B, W, H, Cin, Cout, ks, stride, pad_lo, circ, split, L.ptr(stats) if fused else None)
    if fused:     # channel-pair moments accumulated by the epilogue == moments of the tensor it wrote
        og = out3.double().reshape(B, Wo * Ho, Cout // 2, 2)
        assert torch.allclose(stats[:, :, 0], og.sum((1, 3)), rtol=1e-5, atol=1e-3)
        assert torch.allclose(stats[:, :, 1], (og * og).sum((1, 3)), rtol=1e-5, atol=1e-3)
    # split-K through the cluster/DSMEM reduction is deterministic: a second launch is bit-identical
    out3b = torch.full((B, Wo, Ho, Cout), float("nan"), device="cuda")
    L.call("rldm_conv_tc", L.ptr(xh2), L.ptr(xl2), L.ptr(wt2), L.ptr(bd), L.ptr(td), Cout + 8, L.ptr(rd), L.ptr(out3b),
           B, W, H, Cin, Cout, ks, stride, pad_lo, circ, split, None)
    assert torch.equal(out3, out3b)
    # explicit precision through rldm_conv_tc_ex: terms = 3 is the call above, bit for bit ...
    out3e = torch.full((B, Wo, Ho, Cout), float("nan"), device="cuda")
    L.call("rldm_conv_tc_ex", L.ptr(xh2), L.ptr(xl2), L.ptr(wt2), L.ptr(bd), L.ptr(td), Cout + 8, L.ptr(rd), L.ptr(out3e),
           B, W, H, Cin, Cout, ks, stride, pad_lo, circ, split, None, None, None, None, 0, 3)
    assert torch.equal(out3, out3e)
    # ... terms = 2 ("fp16x2": activations single fp16, weights hi+lo -> Xh*Wh + Xh*Wl): exact for the fp16-rounded
    # activations and the un-rounded weights
    out2 = torch.full((B, Wo, Ho, Cout), float("nan"), device="cuda")
    st2 = torch.zeros(B, Cout // 2, 2, dtype=torch.float64, device="cuda")
    L.call("rldm_conv_tc_ex", L.ptr(xh2), None, L.ptr(wt2), L.ptr(bd), L.ptr(td), Cout + 8, L.ptr(rd), L.ptr(out2),
           B, W, H, Cin, Cout, ks, stride, pad_lo, circ, split, L.ptr(st2) if fused else None, None, None, None, 0, 2)
    y2 = oracle_conv(unpadw(xh2).float().cpu().permute(0, 3, 1, 2), w, b, stride, pad_lo, ks, bool(circ)) \
        + temb[:, :Cout, None, None] + res
    assert relerr(ref_layout(out2.cpu()), y2) < 1e-5
    assert relerr(ref_layout(out2.cpu()), y32) < 1e-3
    if fused:
        og = out2.double().reshape(B, Wo * Ho, Cout // 2, 2)
        assert torch.allclose(st2[:, :, 0], og.sum((1, 3)), rtol=1e-5, atol=1e-3)
    # ... terms = 1 == the legacy plain-fp16 call
    out1 = torch.full((B, Wo, Ho, Cout), float("nan"), device="cuda")
    L.call("rldm_conv_tc_ex", L.ptr(xh), None, L.ptr(wt), L.ptr(bd), L.ptr(td), Cout + 8, L.ptr(rd), L.ptr(out1),
           B, W, H, Cin, Cout, ks, stride, pad_lo, circ, split, None, None, None, None, 0, 1)
    assert torch.equal(out1, out_tc)
    with pytest.raises(L.RldmError):      # x3 needs the low-order activation plane
        L.call("rldm_conv_tc_ex", L.ptr(xh2), None, L.ptr(wt2), L.ptr(bd), None, 0, None, L.ptr(out1),
               B, W, H, Cin, Cout, ks, stride, pad_lo, circ, split, None, None, None, None, 0, 3)
    L.call("rldm_conv_ref", L.ptr(xh2), L.ptr(xl2), L.ptr(wt2), L.ptr(bd), L.ptr(td), Cout + 8, L.ptr(rd),
           L.ptr(out3r), B, W, H, Cin, Cout, ks, stride, pad_lo, circ)
    assert relerr(out3, out3r) < 1e-5
    assert relerr(ref_layout(out3.cpu()), y32) < 1e-5


def test_conv_tc_golden_reference_conv(L, golden):
    g = golden("circ_conv.pt")                # produced by the reference's own Conv2d
    for wk, bk, yk, stride in (("w1", "b1", "y1", 1), ("w2", "b2", "y2", 2)):
        x, w, b, y = g["x"], g[wk], g[bk], g[yk]
        B, Cin, W, H = x.shape
        Cout = w.shape[0]
        out = torch.empty(B, W // stride, H // stride, Cout, device="cuda")
        xh, wt, bd = padw(cl(x).half()).cuda(), pack_w(w).cuda(), b.cuda()     # keep the operands alive across the call
        L.call("rldm_conv_tc", L.ptr(xh), None, L.ptr(wt), L.ptr(bd), None, 0, None,
               L.ptr(out), B, W, H, Cin, Cout, 3, stride, 1, 1, 0, None)
        assert relerr(ref_layout(out.cpu()), y) < 1e-3
        xh2, xl2 = split_half(cl(x))
        xh2, xl2, wt2 = padw(xh2).cuda(), padw(xl2).cuda(), pack_w(w, split=True).cuda()
        L.call("rldm_conv_tc", L.ptr(xh2), L.ptr(xl2), L.ptr(wt2), L.ptr(bd), None, 0, None,
               L.ptr(out), B, W, H, Cin, Cout, 3, stride, 1, 1, 0, None)
        assert relerr(ref_layout(out.cpu()), y) < 1e-5


SHORTCUT_CASES = [
    # B, W, H, Cin (3x3 operand), Cin2 (1x1 shortcut operand), Cout, split_k
    (2, 32, 8, 128, 256, 128, 0),       # UNet up-block geometry: K loop 18 + 4 steps
    (4, 32, 2, 256, 512, 256, 8),       # 64-pixel images, cluster split-K x8 straddling the two operands
    (1, 16, 8, 256, 128, 256, 4),       # down-block width change 128 -> 256
    (8, 256, 16, 128, 256, 128, 0),     # C3 top level at full size: persistent kernel
    (1, 512, 32, 64, 128, 64, 0),       # decoder geometry, BLOCK_N = 64, persistent
]


@pytest.mark.parametrize("case", SHORTCUT_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_tc_with_fused_shortcut(L, case):
    """ResnetBlock2D tail: conv2(a) + conv_shortcut(x_raw) + biases in ONE launch == the two reference convs."""
    B, W, H, Cin, Cin2, Cout, split = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    a = torch.randn(B, Cin, W, H, generator=g)
    xr = torch.randn(B, Cin2, W, H, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    w2 = torch.randn(Cout, Cin2, 1, 1, generator=g) / Cin2 ** 0.5
    b, b2 = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    y = oracle_conv(a, w, b, 1, 1, 3) + F.conv2d(xr, w2, b2)
    ah, al = split_half(cl(a)); xh, xl = split_half(cl(xr))
    ah, al, xh, xl = padw(ah).cuda(), padw(al).cuda(), padw(xh).cuda(), padw(xl).cuda()
    wt, wt2, bd = pack_w(w, split=True).cuda(), pack_w(w2, split=True).cuda(), (b + b2).cuda()
    out = torch.full((B, W, H, Cout), float("nan"), device="cuda")
    stats = torch.zeros(B, Cout // 2, 2, dtype=torch.float64, device="cuda")
    L.call("rldm_conv_tc_shortcut", L.ptr(ah), L.ptr(al), L.ptr(wt), L.ptr(bd), None, 0, None, L.ptr(out),
           B, W, H, Cin, Cout, 3, 1, 1, 1, split, L.ptr(stats) if W * H >= 64 else None,
           L.ptr(xh), L.ptr(xl), L.ptr(wt2), Cin2)
    torch.cuda.synchronize()
    assert relerr(ref_layout(out.cpu()), y) < 1e-5
    if W * H >= 64:
        og = out.double().reshape(B, W * H, Cout // 2, 2)
        assert torch.allclose(stats[:, :, 0], og.sum((1, 3)), rtol=1e-5, atol=1e-3)
    # plain-fp16 operands (one MMA term) take the same path
    out1 = torch.full((B, W, H, Cout), float("nan"), device="cuda")
    w1h, w2h = pack_w(w).cuda(), pack_w(w2).cuda()          # keep the operands alive across the call
    L.call("rldm_conv_tc_shortcut", L.ptr(ah), None, L.ptr(w1h), L.ptr(bd), None, 0, None, L.ptr(out1),
           B, W, H, Cin, Cout, 3, 1, 1, 1, split, None, L.ptr(xh), None, L.ptr(w2h), Cin2)
    torch.cuda.synchronize()
    assert relerr(ref_layout(out1.cpu()), y) < 2e-3
    with pytest.raises(L.RldmError):    # stride 2 cannot carry a same-grid shortcut
        L.call("rldm_conv_tc_shortcut", L.ptr(ah), L.ptr(al), L.ptr(wt), L.ptr(bd), None, 0, None, L.ptr(out),
               B, W, H, Cin, Cout, 3, 2, 1, 1, split, None, L.ptr(xh), L.ptr(xl), L.ptr(wt2), Cin2)


def test_conv_tc_rejects_bad_shapes(L):
    x = torch.zeros(1, 10, 8, 48, dtype=torch.half, device="cuda")
    with pytest.raises(L.RldmError):
        L.call("rldm_conv_tc", L.ptr(x), None, L.ptr(x), None, None, 0, None, L.ptr(x), 1, 8, 8, 48, 64, 3, 1, 1, 1, 0,
               None)


@pytest.mark.parametrize("shape", [(2, 64, 0, 16, 8, 1), (2, 128, 256, 8, 4, 1), (1, 256, 128, 16, 2, 2),
                                   (2, 64, 0, 32, 64, 1), (3, 512, 512, 4, 2, 1)])
def test_gn_stats_and_prep(L, shape):
    B, C0, C1, W, H, up = shape
    g = torch.Generator().manual_seed(C0 + C1 + W)
    x0 = torch.randn(B, C0, W, H, generator=g) * 2 + 0.5
    x1 = torch.randn(B, C1, W, H, generator=g) - 1 if C1 else None
    C = C0 + C1
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    G, eps = 32, 1e-5
    sums = torch.zeros(B, G, 2, dtype=torch.float64, device="cuda")
    x0d = cl(x0).cuda()
    x1d = cl(x1).cuda() if C1 else None
    L.call("rldm_gn_stats", L.ptr(x0d), C0, L.ptr(x1d), C1, L.ptr(sums), B, W * H, G)
    xc = torch.cat([x0, x1], 1) if C1 else x0
    xg = xc.double().reshape(B, G, -1)
    assert torch.allclose(sums[:, :, 0].cpu(), xg.sum(-1), rtol=1e-6, atol=1e-4)
    assert torch.allclose(sums[:, :, 1].cpu(), (xg * xg).sum(-1), rtol=1e-6, atol=1e-4)
    out = torch.empty(B, W * up + 2, H * up, C, dtype=torch.half, device="cuda")
    gd, bd = gamma.cuda(), beta.cuda()
    out_lo = torch.empty_like(out)
    L.call("rldm_prep", L.ptr(x0d), C0, L.ptr(x1d), C1, L.ptr(sums), None, None, L.ptr(gd), L.ptr(bd), eps, G, 1,
           up, 1, L.ptr(out), L.ptr(out_lo), None, None, B, W, H)
    y = F.silu(F.group_norm(xc, G, gamma, beta, eps))
    if up == 2:
        y = F.interpolate(y, scale_factor=2.0, mode="nearest")
    assert relerr(ref_layout(unpadw(out).float().cpu()), y) < 1.5e-3          # fp16 output rounding
    assert relerr(ref_layout(unpadw(out.float() + out_lo.float()).cpu()), y) < 5e-6      # hi + lo: split-fp16
    assert torch.equal(out[:, 0], out[:, -2]) and torch.equal(out[:, -1], out[:, 1])   # circular halo columns
    # raw cast path (no norm, no silu), zero halo
    L.call("rldm_prep", L.ptr(x0d), C0, L.ptr(x1d), C1, None, None, None, None, None, 0.0, 0, 0, up, 0, L.ptr(out), None, None, None,
           B, W, H)
    yr = F.interpolate(xc, scale_factor=2.0, mode="nearest") if up == 2 else xc
    assert torch.equal(ref_layout(unpadw(out).cpu()), yr.half())
    assert float(out[:, 0].abs().max()) == 0.0 and float(out[:, -1].abs().max()) == 0.0
    # dual output: normalised+SiLU operand and the raw operand from one launch
    raw, raw_lo = torch.empty_like(out), torch.empty_like(out)
    L.call("rldm_prep", L.ptr(x0d), C0, L.ptr(x1d), C1, L.ptr(sums), None, None, L.ptr(gd), L.ptr(bd), eps, G, 1,
           up, 1, L.ptr(out), L.ptr(out_lo), L.ptr(raw), L.ptr(raw_lo), B, W, H)
    assert relerr(ref_layout(unpadw(out.float() + out_lo.float()).cpu()), y) < 5e-6
    assert relerr(ref_layout(unpadw(raw.float() + raw_lo.float()).cpu()), yr) < 5e-6
    # channel-pair moments (what conv epilogues accumulate) instead of group sums: same normalisation
    if (C // G) % 2 == 0:
        def pairs(x):
            xp = x.double().reshape(B, x.shape[1] // 2, -1)
            return torch.stack([xp.sum(-1), (xp * xp).sum(-1)], -1).contiguous().cuda()
        p0, p1 = pairs(x0), (pairs(x1) if C1 else None)
        L.call("rldm_prep", L.ptr(x0d), C0, L.ptr(x1d), C1, None, L.ptr(p0), L.ptr(p1), L.ptr(gd), L.ptr(bd), eps, G, 1,
               up, 1, L.ptr(out), L.ptr(out_lo), None, None, B, W, H)
        assert relerr(ref_layout(unpadw(out.float() + out_lo.float()).cpu()), y) < 5e-6


@pytest.mark.parametrize("kernel", ["tcgen05", "mma", "cudacore"])
@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 256, 128), (2, 1024, 128), (1, 40, 64), (1, 320, 64), (3, 128, 64),
                                   (1, 1024, 16), (1, 512, 256), (1, 2048, 32), (5, 1024, 64)])
def test_attention_core(L, shape, kernel, monkeypatch):
    """Default dispatch (tcgen05 kernel for N a multiple of 128 and >= 256; other shapes fall through to the mma.sync
    and CUDA-core kernels), the mma.sync kernel (N % 64 == 0) and the CUDA-core kernel (ragged N) forced where they apply."""
    B, N, C = shape
    if kernel == "cudacore":
        monkeypatch.setenv("RLDM_ATTN_CUDACORE", "1")     # the ragged-N kernel, forced for every shape
    if kernel == "mma":
        monkeypatch.setenv("RLDM_ATTN_MMASYNC", "1")
    L.lib().rldm_reload_env()                              # switches are read once; re-read after changing them
    if kernel != "tcgen05" and N > 1024:
        monkeypatch.undo()
        L.lib().rldm_reload_env()
        pytest.skip("large N only exercised on the tcgen05 kernel")
    g = torch.Generator().manual_seed(N)
    qkv = torch.randn(B, N, 3 * C, generator=g)
    Hh = 8                                                 # tokens are (w, h) with H = 8; output is W-padded
    outp = torch.zeros(B, N // Hh + 2, Hh, C, dtype=torch.half, device="cuda")
    qd = qkv.cuda()
    outp_lo = torch.zeros_like(outp)
    L.call("rldm_attention", L.ptr(qd), L.ptr(outp), L.ptr(outp_lo), B, N, C, Hh)
    out, out_lo = unpadw(outp).reshape(B, N, C), unpadw(outp_lo).reshape(B, N, C)
    q, k, v = qkv.split(C, dim=-1)
    sp = lambda t: t.view(B, N, C // 8, 8).transpose(1, 2)
    y = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B, N, C)
    # single-plane output (the consumer runs below split-fp16 x3): the tcgen05 kernel then carries P as one fp16 plane
    outp1 = torch.zeros_like(outp)
    L.call("rldm_attention", L.ptr(qd), L.ptr(outp1), None, B, N, C, Hh)
    out1 = unpadw(outp1).reshape(B, N, C)
    monkeypatch.undo()
    L.lib().rldm_reload_env()
    assert relerr(out.float().cpu(), y) < 1e-3
    assert relerr((out.float() + out_lo.float()).cpu(), y) < 1e-5
    assert relerr(out1.float().cpu(), y) < 1e-3


@pytest.mark.parametrize("rows", [3, 20, 37])
def test_time_embedding(L, rows):
    """rows = the batch of one forward, or the timestep table of a trajectory (one row per sampling step); 37 rows
    exercise more than two shared-memory row chunks."""
    from oracle import nets
    torch.manual_seed(0)
    B, D0, D4 = rows, 128, 512
    te = nets.TimestepEmbedding(D0, D4)
    projs = [torch.nn.Linear(D4, c) for c in (128, 256, 128, 4)]       # T = 516: a ragged last block of output rows
    t = torch.tensor([999.0, 47.0, 0.0] + [float(940 - 23 * i) for i in range(rows - 3)])
    wp = torch.cat([p.weight for p in projs]).detach()
    bp = torch.cat([p.bias for p in projs]).detach()
    T = wp.shape[0]
    scratch = torch.empty(2, B, D4, device="cuda")
    out = torch.empty(B, T, device="cuda")
    dv = [x.detach().cuda().contiguous() for x in (t, te.linear_1.weight, te.linear_1.bias, te.linear_2.weight,
                                                   te.linear_2.bias, wp, bp)]
    L.call("rldm_temb", *[L.ptr(x) for x in dv], L.ptr(scratch), L.ptr(out), B, D0, D4, T)
    with torch.no_grad():
        emb = te(nets.sinusoidal_timestep(t, D0))
        y = torch.cat([p(F.silu(emb)) for p in projs], dim=1)
    assert relerr(out.cpu(), y, f"temb_rows{rows}") < 2e-5


def test_conv_in_and_conv_out(L):
    g = torch.Generator().manual_seed(9)
    B, W, H = 2, 16, 8
    lat, pe = torch.randn(B, 4, W, H, generator=g), torch.randn(B, 1, W, H, generator=g)
    w = torch.randn(128, 5, 3, 3, generator=g) * 0.2
    b = torch.randn(128, generator=g)
    out = torch.empty(B, W, H, 128, device="cuda")
    ld, pd, wd, bd = lat.cuda(), pe.cuda(), w.permute(2, 3, 1, 0).contiguous().cuda(), b.cuda()
    L.call("rldm_conv_in", L.ptr(ld), 4, L.ptr(pd), 1, L.ptr(wd), L.ptr(bd), L.ptr(out), B, W, H, 128, 1)
    y = oracle_conv(torch.cat([lat, pe], 1), w, b, 1, 1, 3)
    assert relerr(ref_layout(out.cpu()), y) < 1e-5
    for Cout, Cin in ((4, 128), (2, 64)):
        x = torch.randn(B, Cin, W, H, generator=g)
        w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1
        b = torch.randn(Cout, generator=g)
        xh = padw(cl(x).half()).cuda()
        out = torch.empty(B, Cout, W, H, device="cuda")
        wd, bd = w.permute(2, 3, 0, 1).contiguous().cuda(), b.cuda()
        L.call("rldm_conv_out", L.ptr(xh), None, L.ptr(wd), L.ptr(bd), L.ptr(out), B, W, H, Cin, Cout, 1)
        y = oracle_conv(unpadw(xh).float().cpu().permute(0, 3, 1, 2), w, b, 1, 1, 3)
        assert relerr(out.cpu(), y) < 1e-5
        xh2, xl2 = split_half(cl(x))
        xh2, xl2 = padw(xh2).cuda(), padw(xl2).cuda()
        L.call("rldm_conv_out", L.ptr(xh2), L.ptr(xl2), L.ptr(wd), L.ptr(bd), L.ptr(out), B, W, H, Cin, Cout, 1)
        assert relerr(out.cpu(), oracle_conv(x, w, b, 1, 1, 3)) < 1e-5


def test_conv_in_fused_pair_moments(L):
    g = torch.Generator().manual_seed(19)
    for (B, W, H, c0, c1, Cout) in ((3, 32, 8, 4, 1, 128), (2, 64, 16, 4, 0, 256), (2, 16, 16, 2, 0, 64)):
        x0, x1 = torch.randn(B, c0, W, H, generator=g), torch.randn(B, max(c1, 1), W, H, generator=g)
        w = torch.randn(Cout, c0 + c1, 3, 3, generator=g) * 0.2
        b = torch.randn(Cout, generator=g)
        out = torch.empty(B, W, H, Cout, device="cuda")
        stats = torch.zeros(B, Cout // 2, 2, dtype=torch.float64, device="cuda")
        x0d, x1d, wd, bd = x0.cuda(), x1.cuda(), w.permute(2, 3, 1, 0).contiguous().cuda(), b.cuda()
        L.call("rldm_conv_in_stats", L.ptr(x0d), c0, L.ptr(x1d) if c1 else None, c1, L.ptr(wd), L.ptr(bd), L.ptr(out),
               B, W, H, Cout, 1, L.ptr(stats))
        xin = torch.cat([x0, x1], 1) if c1 else x0
        assert relerr(ref_layout(out.cpu()), oracle_conv(xin, w, b, 1, 1, 3)) < 1e-5
        og = out.double().reshape(B, W * H, Cout // 2, 2)
        assert torch.allclose(stats[:, :, 0], og.sum((1, 3)), rtol=1e-5, atol=1e-3)
        assert torch.allclose(stats[:, :, 1], (og * og).sum((1, 3)), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("shape", [(2, 16, 8, 128, 4, 1), (1, 64, 16, 128, 4, 1), (2, 32, 64, 64, 2, 1),
                                   (1, 24, 16, 256, 8, 1), (1, 10, 4, 64, 2, 0), (2, 7, 32, 64, 4, 1)])
def test_norm_conv_out_fused(L, shape):
    """GroupNorm + SiLU + 3x3 conv_out in one launch == F.group_norm -> F.silu -> reference circular conv."""
    B, W, H, Cin, Cout, circ = shape
    g = torch.Generator().manual_seed(W * 31 + Cin)
    x = torch.randn(B, Cin, W, H, generator=g) * 1.5 + 0.3
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1
    b = torch.randn(Cout, generator=g)
    G, eps = 32, 1e-6
    gamma, beta = torch.randn(Cin, generator=g), torch.randn(Cin, generator=g)
    y = oracle_conv(F.silu(F.group_norm(x, G, gamma, beta, eps)), w, b, 1, 1, 3, circular=bool(circ))
    xd, wd, bd, gd, btd = cl(x).cuda(), w.permute(2, 3, 0, 1).contiguous().cuda(), b.cuda(), gamma.cuda(), beta.cuda()
    xg = x.double().reshape(B, G, -1)
    sums = torch.stack([xg.sum(-1), (xg * xg).sum(-1)], -1).contiguous().cuda()
    xp = x.double().reshape(B, Cin // 2, -1)
    pairs = torch.stack([xp.sum(-1), (xp * xp).sum(-1)], -1).contiguous().cuda()
    for sm, pr in ((sums, None), (None, pairs)):
        out = torch.full((B, Cout, W, H), float("nan"), device="cuda")
        L.call("rldm_norm_conv_out", L.ptr(xd), L.ptr(sm), L.ptr(pr), L.ptr(gd), L.ptr(btd), eps, G, 1, L.ptr(wd), L.ptr(bd),
               L.ptr(out), B, W, H, Cin, Cout, circ)
        assert relerr(out.cpu(), y) < 1e-5
    # no normalisation, no activation: a plain 3x3 conv to the reference layout
    out = torch.full((B, Cout, W, H), float("nan"), device="cuda")
    L.call("rldm_norm_conv_out", L.ptr(xd), None, None, None, None, 0.0, 0, 0, L.ptr(wd), L.ptr(bd), L.ptr(out),
           B, W, H, Cin, Cout, circ)
    assert relerr(out.cpu(), oracle_conv(x, w, b, 1, 1, 3, circular=bool(circ))) < 1e-5


def test_layout_helpers(L):
    x = torch.randn(2, 5, 8, 4)
    d = torch.empty(2, 8, 4, 5, device="cuda")
    xd = x.cuda()
    L.call("rldm_ref_to_cl", L.ptr(xd), L.ptr(d), 2, 5, 8, 4)
    assert torch.equal(d.cpu(), cl(x))
    r = torch.empty(2, 5, 8, 4, device="cuda")
    L.call("rldm_cl_to_ref", L.ptr(d), L.ptr(r), 2, 5, 8, 4)
    assert torch.equal(r.cpu(), x)


@pytest.mark.parametrize("kind", ["ddim", "ddpm", "dpm"])
def test_scheduler_step_matches_oracle(kind):
    import rangeldm_b200 as R
    from oracle import schedulers as O
    n = 20
    if kind == "ddim":
        s, o = R.DDIMScheduler(clip_sample=False), O.OracleDDIMScheduler()
    elif kind == "ddpm":
        s, o = R.DDPMScheduler(clip_sample=False), O.OracleDDPMScheduler()
    else:
        s, o = R.DPMSolverMultistepScheduler(timestep_spacing="leading"), O.OracleDPMSolverMultistepScheduler(
            timestep_spacing="leading")
    s.set_timesteps(n)
    o.set_timesteps(n)
    assert torch.equal(s.timesteps, o.timesteps)             # integer table: bit exact
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 16, 6, generator=g)
    xa, xb = x.cuda(), x.clone()
    for t in s.timesteps:
        eps = torch.randn(x.shape, generator=g)
        noise = torch.randn(x.shape, generator=g)
        if kind == "ddpm":
            xa = s.step(eps.cuda(), t, xa, variance_noise=noise.cuda()).prev_sample
            xb = o.step(eps, t, xb, variance_noise=noise)
        else:
            xa = s.step(eps.cuda(), t, xa).prev_sample
            xb = o.step(eps, t, xb)
        assert relerr(xa, xb) < 1e-5, int(t)


def test_range_to_points_matches_reference_golden(L, golden, tmp_path):
    """rldm_range_to_points (through RangeImageGeometry) against the reference's own to_pc_torch outputs, the CPU
    oracle at KITTI size, and the .bin writer contract (float32 N x 4, |xyz| < 90, original order)."""
    import numpy as np
    import rangeldm_b200 as R
    from oracle import geometry as G
    d = golden("range_to_points.pt")
    for name, kw in (("linear", {}), ("log", {"log": True}), ("inverse", {"inverse": True})):
        geom = R.RangeImageGeometry(d["incl"].numpy(), d["height"].numpy(), mean=d["mean"], std=d["std"], **kw)
        pts = geom.to_pc_torch(d["image"].cuda())
        assert pts.shape == d[name].shape
        assert relerr(pts.cpu(), d[name]) < 1e-5
    geom = R.RangeImageGeometry(d["incl"].numpy(), d["height"].numpy())
    # full KITTI size (BASELINE shape 2 x 1024 x 64), against the oracle
    g = torch.Generator().manual_seed(5)
    img = torch.rand(3, 2, 1024, 64, generator=g) * 1.3 - 0.5
    pts, depth = geom.to_pc_torch(img.cuda(), return_depth=True)
    ref = G.to_points(img, d["incl"], d["height"])
    assert relerr(pts.cpu(), ref) < 1e-5
    assert torch.allclose(depth.cpu(), ref[..., :3].norm(dim=-1), rtol=1e-5, atol=1e-4)
    assert torch.equal(pts[..., 3].cpu(), img[:, 1].reshape(3, -1))             # remission is copied bit-exactly
    # single-channel image -> xyz only; writer: rows with depth < 90 m in order, float32 N x 4 on disk
    assert geom.to_pc_torch(img[:, :1].cuda()).shape == (3, 65536, 3)
    paths = geom.write_bins(d["image"].cuda(), str(tmp_path), first_index=7)
    assert [p.rsplit("/", 1)[1] for p in paths] == ["7.bin", "8.bin"]
    rows = np.fromfile(paths[0], dtype=np.float32).reshape(-1, 4)
    want = d["masked_rows"].numpy()
    assert rows.shape == want.shape and np.allclose(rows, want, rtol=1e-5, atol=1e-4)
    with pytest.raises(RuntimeError):
        geom.to_pc_torch(img)                                                    # CPU tensor: no fallback


def test_points_to_voxel_matches_reference_golden(L, golden):
    """rldm_points_to_voxel against the reference's own to_voxel output: (1) the splat kernel alone, fed the reference's
    points: agreement to float-atomic rounding order; (2) chained behind rldm_range_to_points (RangeImageGeometry.to_voxel),
    where the ~2 ulp libm difference of the point coordinates (CUDA sincosf vs the CPU's cosf, 8e-6 m at 60 m range) is
    amplified by the voxel pitch; (3) the reference's default 1x1024x1024 BEV grid at KITTI size against the CPU oracle."""
    import ctypes
    import rangeldm_b200 as R
    from oracle import geometry as G
    d = golden("range_to_points.pt")
    D, Hg, Wg = d["voxel_grid"]

    def check(v, ref, name, tol):
        """log-densities to `tol`; features are a ratio that is ill-conditioned where the total vote weight is ~1e-4
        (a 1e-6-voxel shift of one point moves it by 1e-3 there, in the reference too), so they are compared as the
        well-conditioned weighted sums feature * clamp(density)."""
        assert v.shape == ref.shape
        Dd = v.shape[1] // 2
        assert relerr(v[:, :Dd].cpu(), ref[:, :Dd], name + "_density") < tol
        w = torch.expm1(ref[:, :Dd]).clamp(min=1e-4)
        assert relerr(v[:, Dd:].cpu() * w, ref[:, Dd:] * w, name + "_weighted_feature") < 2 * tol

    # (1) splat + finalize kernels alone
    pts = d["linear"].cuda().contiguous()
    B, N, P = pts.shape
    scratch = torch.empty((2, B, D * Hg * Wg), device="cuda")
    vox = torch.empty((B, 2 * D, Hg, Wg), device="cuda")
    rng = (ctypes.c_float * 6)(*d["voxel_range"])
    L.call("rldm_points_to_voxel", L.ptr(pts), B, N, P, rng, D, Hg, Wg, 1, L.ptr(scratch), L.ptr(vox))
    check(vox, d["voxel"], "to_voxel_kernel", 1e-6)
    assert (vox[:, D:].cpu() - d["voxel"][:, D:]).abs().max() < 1e-3
    # (2) range image -> points -> volume
    geom = R.RangeImageGeometry(d["incl"].numpy(), d["height"].numpy(), mean=d["mean"], std=d["std"],
                                grid_sizes=d["voxel_grid"], pc_range=d["voxel_range"])
    v = geom.to_voxel(d["image"].cuda())
    check(v, d["voxel"], "to_voxel_golden", 5e-5)
    assert (R.RangeImageGeometry.bev_image(v[0]) == G.bev_image(d["voxel"][0])).mean() > 0.999   # uint8 truncation ties
    # (3) default BEV grid (0.05 m pitch), KITTI image size
    geom = R.RangeImageGeometry(d["incl"].numpy(), d["height"].numpy())
    img = torch.rand(2, 2, 1024, 64, generator=torch.Generator().manual_seed(3)) * 1.2 - 0.45
    v = geom.to_voxel(img.cuda())
    ref = G.to_voxel(img, d["incl"], d["height"])
    assert v.shape == (2, 2, 1024, 1024)
    check(v, ref, "to_voxel_kitti", 5e-4)
    # un-normalised densities: total vote weight == number of points that fall inside the grid (z is a single layer)
    geom.normalize_volume_densities = False
    dens = geom.to_voxel(img.cuda())[:, 0]
    pts = G.to_points(img, d["incl"], d["height"])
    inside = ((pts[..., 0].abs() <= 25.6) & (pts[..., 1].abs() <= 25.6)).sum(1).float()
    assert torch.allclose(dens.sum((1, 2)).cpu(), inside, rtol=2e-3)


def test_points_to_range_matches_reference_golden(L, golden):
    """rldm_points_to_range (RangeImageGeometry.from_points) against the reference's projection + miss-value fill +
    normalisation: identical pixel assignment (nearest return wins) and masks, values to float rounding; then the
    round trip range image -> points -> range image at KITTI size reproduces every pixel."""
    import rangeldm_b200 as R
    d = golden("range_to_points.pt")
    pts = d["proj_points"].cuda()
    for name, kw in (("linear", {}), ("log", {"log": True}), ("inverse", {"inverse": True})):
        geom = R.RangeImageGeometry(d["incl"].numpy(), d["height"].numpy(), **kw)
        out = geom.from_points(pts, width=128)
        assert torch.equal(out["mask"].cpu(), d["proj_mask_" + name])
        assert torch.equal(out["car_window_mask"].cpu(), d["proj_car_" + name])
        ref = d["proj_" + name]
        assert out["jpg"].shape == ref.shape
        assert torch.equal(out["jpg"][1].cpu(), ref[1])                        # remission of the winning point: exact
        assert relerr(out["jpg"][0].cpu(), ref[0], "points_to_range_" + name) < 1e-6
    # round trip at KITTI size: every pixel of a dense range image comes back (beam assignment, column rounding)
    geom = R.RangeImageGeometry(d["incl"].numpy(), d["height"].numpy())
    g = torch.Generator().manual_seed(12)
    img = torch.stack([torch.rand(1024, 64, generator=g) * 1.2 - 0.4, torch.rand(1024, 64, generator=g)])[None]
    cloud = geom.to_pc_torch(img.cuda())[0]
    back = geom.from_points(cloud, width=1024)
    assert bool(back["mask"].all()) and not bool(back["car_window_mask"].any())
    assert relerr(back["jpg"].cpu(), img[0], "range_round_trip") < 1e-5
    # empty cloud: everything is the fill value, nothing is a car window
    empty = geom.from_points(torch.zeros(0, 4, device="cuda"), width=64)
    assert not bool(empty["mask"].any()) and not bool(empty["car_window_mask"].any())
    assert torch.allclose(empty["jpg"][0].cpu(), torch.full((64, 64), (100.0 - 20.0) / 40.0))
    with pytest.raises(RuntimeError):
        geom.from_points(torch.zeros(4, 4))
