"""GPU tests of the fused multi-layer kernel (csrc/fused_levels.cu): a run of small ops as ONE persistent launch must
give what the same ops give as separate launches (same operands, same arithmetic; only the fp32 summation order of
the K split differs), and both must match the fp32 oracle."""
import ctypes

import pytest
import torch

from conftest import relerr
from test_models_gpu import make_unet, make_vae

pytestmark = pytest.mark.gpu


def _forward_both(u, x, t, monkeypatch):
    from rangeldm_b200 import _lib, engine
    outs, nodes = [], []
    monkeypatch.setattr(engine, "FUSE_PREP", False)     # (the experiment fuses whole prep + conv runs instead)
    for fuse in (True, False):
        monkeypatch.setattr(engine, "FUSE_LEVELS", fuse)
        u.invalidate_plans()
        outs.append(u(x, t).sample)
        prog = u.plan(x.shape[0], x.shape[2], x.shape[3]).prog
        nodes.append((len(prog.exec_ops), sum(1 for o in prog.exec_ops if o.kind == _lib.OP_FUSED)))
    u.invalidate_plans()
    return outs, nodes


@pytest.fixture
def x3(monkeypatch):
    from rangeldm_b200 import engine
    monkeypatch.setattr(engine, "PRECISION", 3)
    monkeypatch.setattr(engine, "PRECISION_TOP", 3)
    monkeypatch.setattr(engine, "PRECISION_TOP_SAMPLER", 3)
    monkeypatch.setattr(engine, "PRECISION_VAE", 3)
    yield


def test_fused_levels_tiny_unet_matches_unfused_and_oracle(monkeypatch, x3):
    from oracle import nets
    from oracle.make_golden import TINY_UNET, seeded
    ou = seeded(nets.OracleUNet2DModel, 4321, **TINY_UNET)
    u = make_unet(TINY_UNET, ou)
    x = torch.randn(2, 5, 32, 8, generator=torch.Generator().manual_seed(0))
    (a, b), nodes = _forward_both(u, x.cuda(), torch.tensor(500), monkeypatch)
    assert nodes[0][1] >= 1 and nodes[1][1] == 0 and nodes[0][0] < nodes[1][0]
    assert relerr(a, b, "fused_vs_unfused_tiny") < 5e-6
    with torch.no_grad():
        ref = ou(x, torch.tensor(500))
    assert relerr(a, ref, "fused_tiny_vs_oracle") < 1e-4
    # replays of the same compiled run (the grid barrier's generation word keeps counting across launches)
    c = u(x.cuda(), torch.tensor(500)).sample
    d = u(x.cuda(), torch.tensor(500)).sample
    assert relerr(c, d) < 1e-6


@pytest.mark.parametrize("batch", [8, 3, 1])
def test_fused_levels_c3_unet_matches_unfused_and_oracle(batch, monkeypatch, x3):
    """BENCH shape (batch 8), a ragged batch (partial 128-pixel tiles at the 32x2 level) and a single image."""
    from oracle import nets
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C3)
    u = make_unet(nets.UNET_C3, ou)
    x = torch.randn(batch, 5, 256, 16, generator=torch.Generator().manual_seed(20 + batch))
    t = torch.tensor([17, 500, 999, 47, 940, 3, 250, 800][:batch])
    (a, b), nodes = _forward_both(u, x.cuda(), t, monkeypatch)
    assert nodes[0][1] >= 3 and nodes[0][0] <= 40, nodes          # <= 40 graph nodes per forward instead of 168
    assert relerr(a, b, f"fused_vs_unfused_c3_b{batch}") < 1e-5
    with torch.no_grad():
        ref = ou(x, t)
    assert relerr(a, ref, f"fused_c3_b{batch}_vs_oracle") < 1e-3


def test_fused_levels_reduced_precision_and_graph_replay(monkeypatch):
    """fp16x2 / fp16 operands inside a fused run (1 or 2 weight planes, no activation lo plane), captured in a CUDA
    graph and replayed.  With single-fp16 activations the summation-order noise of the K split (1e-7) flips the odd
    fp16 rounding of an activation (2^-11 of that element), so fused and unfused programs agree to ~1e-4 rather than to
    1e-6; both stay inside the tolerance against the fp32 oracle.  (The bit-level check of the reduced-precision conv
    path is test_fused_conv_matches_standalone_conv below.)"""
    from rangeldm_b200 import engine
    from oracle import nets
    from oracle.make_golden import TINY_UNET, seeded
    ou = seeded(nets.OracleUNet2DModel, 4321, **TINY_UNET)
    u = make_unet(TINY_UNET, ou)
    x = torch.randn(2, 5, 32, 8, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = ou(x, torch.tensor(300))
    x = x.cuda()
    for terms in (2, 1):
        monkeypatch.setattr(engine, "PRECISION", terms)
        monkeypatch.setattr(engine, "PRECISION_TOP", terms)
        (a, b), nodes = _forward_both(u, x, 300, monkeypatch)
        assert nodes[0][1] >= 1
        assert relerr(a, b, f"fused_vs_unfused_terms{terms}") < 2e-3
        assert relerr(a, ref, f"fused_terms{terms}_vs_oracle") < 1e-3 and relerr(b, ref) < 1e-3
        monkeypatch.setattr(engine, "FUSE_LEVELS", True)
        monkeypatch.setattr(engine, "FUSE_PREP", False)
        u.invalidate_plans()
        plan = u.plan(2, 32, 8)
        plan.x_in.copy_(x); plan.t_buf.fill_(300.0)
        plan.prog.run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            plan.prog.run()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        assert relerr(plan.out, a) < 1e-6
    u.invalidate_plans()


FUSED_CONV_CASES = [
    # B, W, H, Cin, Cout, ks, stride, residual, temb
    (8, 32, 2, 256, 256, 3, 1, True, True),      # level 3 of C3: 64-pixel images, two per tile, K split 12 ways
    (3, 32, 2, 512, 256, 3, 1, False, False),    # ragged batch: partial last tile
    (2, 64, 4, 128, 256, 3, 1, False, True),
    (8, 128, 8, 128, 128, 3, 2, False, False),   # Downsample2D: stride 2 -> 64x4
    (8, 64, 4, 256, 768, 1, 1, False, False),    # qkv projection
    (8, 128, 8, 128, 128, 1, 1, True, False),    # out projection + residual at level 1
]


@pytest.mark.parametrize("terms", [3, 2, 1])
@pytest.mark.parametrize("case", FUSED_CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_fused_conv_matches_standalone_conv(case, terms, monkeypatch):
    """[raw-cast prep, conv] as one fused run against the same two ops as separate launches: identical fp16 operands on
    both sides, so the outputs agree to fp32 summation order and the GroupNorm moments to 1e-6, for every operand
    precision, with bias + time embedding + residual in the epilogue."""
    import rangeldm_b200 as R
    from rangeldm_b200 import engine, models
    monkeypatch.setattr(engine, "FUSE_LEVELS", True)           # (an opt-in experiment: see engine.FUSE_LEVELS)
    monkeypatch.setattr(engine, "FUSE_PREP", False)
    B, W, H, Cin, Cout, ks, stride, use_res, use_temb = case
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    conv = models.LoRACompatibleConv(Cin, Cout, ks, stride=stride, padding=ks // 2).to(dev)
    conv.circular = True
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (Cin * ks * ks) ** 0.5)
        conv.bias.copy_(torch.randn(Cout, generator=g))
    x = torch.randn(B, W, H, Cin, generator=g).to(dev)
    Wo, Ho = W // stride, H // stride
    res = torch.randn(B, Wo, Ho, Cout, generator=g).to(dev)
    temb = torch.randn(B, Cout + 16, generator=g).to(dev)
    outs = []
    for fuse in (True, False):
        pg = engine.Program(dev, fuse=fuse)
        bd = engine.Builder(pg, B, cache={}, terms_of=lambda w: terms)
        xa = engine.Act(pg.hold(x.clone()), B, W, H, Cin)
        ra = engine.Act(pg.hold(res.clone()), B, Wo, Ho, Cout) if use_res else None
        a = bd.prep(xa, None, None, terms=terms)
        out = bd.conv(a, W, H, conv, temb=(temb, Cout + 16) if use_temb else None, residual=ra, stats=True, terms=terms)
        bd.finish()
        pg.finalize()
        assert sum(1 for o in pg.exec_ops if o.kind == R._lib.OP_FUSED) == (1 if fuse else 0)
        for _ in range(2):          # second run: memset + barrier generation carry over
            pg.run()
        torch.cuda.synchronize()
        outs.append((out.t.clone(), out.stats.clone() if out.stats is not None else None, pg))
    (fa, fs, _), (ua, us, _) = outs
    assert relerr(fa, ua, f"fused_conv_vs_standalone_terms{terms}") < 2e-6
    if fs is not None:
        assert torch.allclose(fs, us, rtol=1e-6, atol=1e-3)


def test_fused_create_rejects_unsupported_ops():
    from rangeldm_b200 import _lib
    from rangeldm_b200._lib import RldmOp
    op = RldmOp()
    op.kind = _lib.OP_SCHED_STEP
    assert _lib.lib().rldm_fused_supported(ctypes.byref(op)) == 0
    h = ctypes.c_void_p()
    arr = (RldmOp * 1)(op)
    assert _lib.lib().rldm_fused_create(arr, 1, None, 0, ctypes.byref(h)) != 0 and not h.value
    assert b"cannot run inside a fused segment" in _lib.lib().rldm_last_error()


# ---- convolutions that produce their own operand (rldm_conv_tc_fused) -------------------------------------------------
OWN_OPERAND_CASES = [
    # B, W, H, C0, C1, Cout, ks, stride, up, norm, silu
    (8, 32, 2, 256, 0, 256, 3, 1, 1, True, True),       # level 3 ResnetBlock2D conv: GroupNorm + SiLU, two images per tile
    (8, 32, 2, 256, 256, 256, 3, 1, 1, True, True),     # up-block conv1 over a skip concat (512 channels, K split 8 ways)
    (3, 64, 4, 256, 128, 256, 3, 1, 1, True, True),     # 384-channel concat: one GroupNorm group straddles the two sources
    (8, 64, 4, 256, 0, 256, 3, 2, 1, False, False),     # Downsample2D: raw cast, stride 2
    (8, 32, 2, 256, 0, 256, 3, 1, 2, False, False),     # Upsample2D: nearest 2x folded into the operand production
    (2, 64, 4, 256, 0, 768, 1, 1, 1, True, False),      # attention: GroupNorm without SiLU, qkv projection
    (5, 128, 8, 128, 0, 128, 3, 1, 1, True, True),      # level 1, ragged batch: 40 tiles, K split 2
]


def _pair_moments(x):
    """[B][C/2][2] doubles: (sum, sum of squares) per channel pair over the pixels of (B, W, H, C)."""
    B, W, H, C = x.shape
    xd = x.double().reshape(B, W * H, C // 2, 2)
    return torch.stack([xd.sum((1, 3)), (xd * xd).sum((1, 3))], dim=-1).contiguous()


@pytest.mark.parametrize("terms", [3, 1])
@pytest.mark.parametrize("case", OWN_OPERAND_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_own_operand_matches_prep_then_conv(case, terms, monkeypatch):
    """A small-layer convolution that turns the fp32 stream into its own fp16 operand inside the kernel against the same
    convolution behind a rldm_prep launch: the operands are bit-identical, the kernel and its summation order are the
    same, so the outputs are BIT-IDENTICAL (the GroupNorm moments agree to the order of the double atomics)."""
    import rangeldm_b200 as R
    from rangeldm_b200 import engine, models
    B, W, H, C0, C1, Cout, ks, stride, up, use_norm, silu = case
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    Cin = C0 + C1
    conv = models.LoRACompatibleConv(Cin, Cout, ks, stride=stride, padding=ks // 2).to(dev)
    conv.circular = True
    norm = torch.nn.GroupNorm(32, Cin, eps=1e-5).to(dev) if use_norm else None
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (Cin * ks * ks) ** 0.5)
        conv.bias.copy_(torch.randn(Cout, generator=g))
        if norm is not None:
            norm.weight.copy_(1 + 0.3 * torch.randn(Cin, generator=g)); norm.bias.copy_(0.3 * torch.randn(Cin, generator=g))
    x0 = (torch.randn(B, W, H, C0, generator=g) * 1.7 + 0.3).to(dev)
    x1 = torch.randn(B, W, H, C1, generator=g).to(dev) if C1 else None
    outs = []
    for own in (True, False):
        monkeypatch.setattr(engine, "FUSE_PREP", own)
        pg = engine.Program(dev)
        bd = engine.Builder(pg, B, cache={}, terms_of=lambda w: terms)
        a0 = engine.Act(pg.hold(x0.clone()), B, W, H, C0, stats=pg.hold(_pair_moments(x0)))
        a1 = engine.Act(pg.hold(x1.clone()), B, W, H, C1, stats=pg.hold(_pair_moments(x1))) if C1 else None
        opnd = bd.prep(a0, a1, norm, silu=silu, up=up, terms=terms, defer=True)
        out = bd.conv(opnd, W * up, H * up, conv, stats=True, terms=terms)
        bd.finish()
        pg.finalize()
        kinds = [o.kind for o in pg.ops]
        assert kinds.count(R._lib.OP_PREP) == (0 if own else 1), kinds
        for _ in range(2):
            pg.run()
        torch.cuda.synchronize()
        outs.append((out.t.clone(), out.stats.clone(), pg))
    (fa, fs, _), (ua, us, _) = outs
    assert torch.equal(fa, ua), relerr(fa, ua)
    assert torch.allclose(fs, us, rtol=1e-9, atol=1e-6)
    # and against the fp32 PyTorch reference of the same ops
    xin = x0 if x1 is None else torch.cat([x0, x1], dim=-1)
    xr = xin.permute(0, 3, 1, 2).cpu()
    with torch.no_grad():
        y = xr
        if norm is not None:
            y = torch.nn.functional.group_norm(y, 32, norm.weight.cpu(), norm.bias.cpu(), 1e-5)
        if silu:
            y = torch.nn.functional.silu(y)
        if up == 2:
            y = torch.nn.functional.interpolate(y, scale_factor=2.0, mode="nearest")
        from oracle.nets import circ_conv2d
        ref = circ_conv2d(y, conv.weight.cpu(), conv.bias.cpu(), stride, ks // 2)
    assert relerr(fa.permute(0, 3, 1, 2).cpu(), ref, f"conv_own_operand_terms{terms}") < (1e-5 if terms == 3 else 2e-3)


def test_resnet_with_folded_shortcut_own_operands(monkeypatch, x3):
    """A whole ResnetBlock2D over a skip concat with its 1x1 conv_shortcut folded into conv2: conv1 produces its operand
    from (h | skip) with GroupNorm + SiLU, conv2 produces BOTH its operand (from conv1's output) and the raw shortcut
    operand (from (h | skip)) -- against the same block behind rldm_prep launches."""
    from rangeldm_b200 import engine, models
    from rangeldm_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(77)
    B, W, H = 8, 64, 4
    rb = models.ResnetBlock2D(384, 256, 512).to(dev)
    for prm in rb.parameters():
        with torch.no_grad():
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.05 if prm.ndim > 1 else 0.3) + (1.0 if prm.ndim == 1 and prm.numel() in (384, 256) else 0.0))
    for m in rb.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.circular = True
    x0 = torch.randn(B, W, H, 256, generator=g).to(dev)
    x1 = torch.randn(B, W, H, 128, generator=g).to(dev)
    temb = torch.randn(B, 256, generator=g).to(dev)
    outs = []
    for own in (True, False):
        monkeypatch.setattr(engine, "FUSE_PREP", own)
        pg = engine.Program(dev)
        bd = engine.Builder(pg, B, cache={})
        bd.temb, bd.temb_rows = (pg.hold(temb.clone()), 256), {id(rb): 0}
        a0 = engine.Act(pg.hold(x0.clone()), B, W, H, 256, stats=pg.hold(_pair_moments(x0)))
        a1 = engine.Act(pg.hold(x1.clone()), B, W, H, 128, stats=pg.hold(_pair_moments(x1)))
        out = bd.resnet(rb, a0, a1, free_inputs=False)
        bd.finish(); pg.finalize()
        assert [o.kind for o in pg.ops].count(_lib.OP_PREP) == (0 if own else 2)
        pg.run(); torch.cuda.synchronize()
        outs.append(out.t.clone())
    assert relerr(outs[0], outs[1], "resnet_own_operands_vs_prep") < 2e-6


# ---- convolutions that emit the next GroupNorm's operand (rldm_conv_tc_emit) -------------------------------------------
EMIT_CASES = [
    # B, W, H, Cin, Cout, ks, stride, residual, silu
    (8, 32, 2, 256, 256, 3, 1, False, True),       # level 3 conv1: two images per tile, K split 8 ways inside the cluster
    (3, 32, 2, 256, 256, 1, 1, True, False),       # attention out-projection, ragged batch: half-empty last tile, no K split
    (8, 64, 4, 256, 256, 3, 1, True, True),        # level 2: an image is two M tiles -> cluster (2, 1, 4)
    (2, 64, 4, 256, 256, 1, 1, True, True),
    (8, 128, 8, 128, 128, 3, 1, False, True),      # level 1: eight M tiles per image -> cluster (8, 1, 1)
    (5, 128, 8, 128, 128, 1, 1, True, False),
    (8, 128, 8, 128, 128, 3, 2, False, True),      # Downsample2D: stride 2, output on the 64 x 4 grid
    (4, 32, 2, 512, 256, 3, 1, False, True),       # 72 K steps
]


@pytest.fixture
def emit_any_cluster(monkeypatch):
    """RLDM_EMIT_MAXCLM=8: emitting convolutions also for images of 2..8 M tiles (the default keeps them to one tile)."""
    import rangeldm_b200 as R
    monkeypatch.setenv("RLDM_EMIT_MAXCLM", "8")
    R._lib.lib().rldm_reload_env()
    yield
    monkeypatch.delenv("RLDM_EMIT_MAXCLM")
    R._lib.lib().rldm_reload_env()


@pytest.mark.parametrize("case", EMIT_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_emit_matches_conv_then_prep(case, monkeypatch, emit_any_cluster):
    """conv -> GroupNorm (+ SiLU) -> fp16 operand: produced by the convolution's own epilogue (cluster-complete moments)
    against the same convolution followed by a rldm_prep launch.  The fp32 output and its channel-pair moments are the
    same kernel's (equal up to the K-split geometry); the operand differs by the summation order of the moments only:
    at most one fp16 ulp on a handful of values."""
    import rangeldm_b200 as R
    from rangeldm_b200 import engine, models
    B, W, H, Cin, Cout, ks, stride, use_res, silu = case
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(abs(hash(case)) % (2 ** 31))
    conv = models.LoRACompatibleConv(Cin, Cout, ks, stride=stride, padding=ks // 2).to(dev)
    conv.circular = True
    norm = torch.nn.GroupNorm(32, Cout, eps=1e-5).to(dev)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) / (Cin * ks * ks) ** 0.5)
        conv.bias.copy_(torch.randn(Cout, generator=g))
        norm.weight.copy_(1 + 0.3 * torch.randn(Cout, generator=g)); norm.bias.copy_(0.3 * torch.randn(Cout, generator=g))
    x0 = (torch.randn(B, W, H, Cin, generator=g) * 1.3 + 0.2).to(dev)
    res = torch.randn(B, W // stride, H // stride, Cout, generator=g).to(dev) if use_res else None
    outs = []
    for emit in (True, False):
        monkeypatch.setattr(engine, "EMIT_PREP", emit)
        pg = engine.Program(dev)
        pg.no_reuse = True
        bd = engine.Builder(pg, B, cache={}, terms_of=lambda w: 1)
        a0 = engine.Act(pg.hold(x0.clone()), B, W, H, Cin)
        ra = engine.Act(pg.hold(res.clone()), B, W // stride, H // stride, Cout) if use_res else None
        opnd = bd.prep(a0, None, None, terms=1)
        h = bd.conv(opnd, W, H, conv, residual=ra, stats=True, terms=1)
        o2 = bd.prep(h, None, norm, silu=silu, terms=1)
        bd.finish(); pg.finalize()
        kinds = [o.kind for o in pg.ops]
        assert kinds.count(R._lib.OP_PREP) == (1 if emit else 2), kinds
        for _ in range(2):
            pg.run()
        torch.cuda.synchronize()
        outs.append((h.t.clone(), h.stats.clone(), o2.hi.clone()))
    (fa, fs, fo), (ua, us, uo) = outs
    assert relerr(fa, ua, "conv_emit_fp32_out") < 2e-6
    assert torch.allclose(fs, us, rtol=1e-6, atol=1e-3)
    d = (fo.float() - uo.float()).abs()
    tol = 2e-3 * uo.float().abs().clamp_min(1.0)            # two fp16 ulps
    assert bool((d <= tol).all()), float((d / tol).max())
    assert float((d > 0).float().mean()) < 0.02             # and almost every value is bit-identical
    # halo columns of the W-padded operand: wrap of the image's first / last column
    assert torch.equal(fo[:, 0], fo[:, -2]) and torch.equal(fo[:, -1], fo[:, 1])


def test_unet_with_emitting_convolutions_matches_oracle(monkeypatch, emit_any_cluster):
    """C3 UNet forward, batch 3, with emitting convolutions on EVERY level (RLDM_EMIT_MAXCLM=8; the default emits on level
    3 only): 40 prep launches fewer, same parity gate as the default path."""
    import rangeldm_b200 as R
    from rangeldm_b200 import engine
    from oracle import nets
    from oracle.make_golden import seeded
    monkeypatch.setattr(engine, "EMIT_PREP", True)
    dev = torch.device("cuda")
    ou = seeded(nets.OracleUNet2DModel, 11, **nets.UNET_C3)
    u = R.UNet2DModel(**nets.UNET_C3)
    R.replace_down(u); R.replace_conv(u)
    u.load_state_dict(ou.state_dict())
    u = u.to(dev)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 5, 256, 16, generator=g)
    t = torch.tensor([900, 417, 12])
    with torch.no_grad():
        ref = ou(x, t)
    y = u(x.to(dev), t.to(dev)).sample
    plan = u.plan(3, 256, 16, 1)
    assert sum(1 for op in plan.prog.ops if op.kind == R._lib.OP_CONV_TC and op.p[19]) == 40
    assert relerr(y.cpu(), ref, "unet_c3_emit_prep") < 5e-4


# ---- nearest-2x upsampling folded into the convolution (rldm_conv_tc_up2) ----------------------------------------------
@pytest.mark.parametrize("terms", [3, 1])
@pytest.mark.parametrize("shape", [(8, 256, 16, 256, 256), (3, 512, 32, 128, 128)], ids=lambda s: "x".join(map(str, s)))
def test_upsample_folded_into_the_convolution(shape, terms, monkeypatch):
    """Upsample2D of the decoder: four 2x2 phase convolutions over the low-resolution operand against nearest-2x in the
    prep pass + the 3x3 convolution, and against the fp32 PyTorch reference (circular along W, zero pad along H)."""
    import rangeldm_b200 as R
    from rangeldm_b200 import engine, models
    B, W, H, Cin, Cout = shape
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(31 + W)
    us = models.Upsample2D(Cin, use_conv=True, out_channels=Cout).to(dev)
    us.conv.circular = True
    with torch.no_grad():
        us.conv.weight.copy_(torch.randn(us.conv.weight.shape, generator=g) / (9 * Cin) ** 0.5)
        us.conv.bias.copy_(torch.randn(Cout, generator=g))
    x0 = torch.randn(B, W, H, Cin, generator=g).to(dev)
    outs = []
    for fold in (True, False):
        monkeypatch.setattr(engine, "FOLD_UPSAMPLE", fold)
        pg = engine.Program(dev)
        bd = engine.Builder(pg, B, cache={}, terms_of=lambda w: terms)
        a0 = engine.Act(pg.hold(x0.clone()), B, W, H, Cin)
        out = bd.upsample(us, a0)
        bd.finish(); pg.finalize()
        kinds = [o.kind for o in pg.ops]
        assert kinds.count(R._lib.OP_CONV_UP2) == (1 if fold else 0) and kinds.count(R._lib.OP_CONV_TC) == (0 if fold else 1)
        for _ in range(2):
            pg.run()
        torch.cuda.synchronize()
        outs.append((out.t.clone(), out.stats.clone()))
    (fa, fs), (ua, us_) = outs
    assert fa.shape == (B, 2 * W, 2 * H, Cout)
    tol = 1e-5 if terms == 3 else 2e-3       # (the combined phase weights are rounded once more than the 3x3 taps)
    assert relerr(fa, ua, f"upsample_folded_vs_prep_terms{terms}") < tol
    assert torch.allclose(fs, us_, rtol=1e-5 if terms == 3 else 1e-2, atol=1e-1)
    from oracle.nets import circ_conv2d
    xr = torch.nn.functional.interpolate(x0.permute(0, 3, 1, 2).cpu(), scale_factor=2.0, mode="nearest")
    with torch.no_grad():
        ref = circ_conv2d(xr, us.conv.weight.cpu(), us.conv.bias.cpu(), 1, 1)
    assert relerr(fa.permute(0, 3, 1, 2).cpu(), ref, f"upsample_folded_vs_torch_terms{terms}") < (1e-5 if terms == 3 else 2e-3)
