"""GPU box: where does the time of a convolution that produces its own operand (rldm_conv_tc_fused, RLDM_FUSE_PREP=1)
go?  For a few UNet level 1-3 shapes: graphed time of [conv with own operand] against [rldm_prep + conv], and the
clock64() stamps of CTA 0 (entry, prologue done, operand produced, first stage landed, last MMA, accumulator, end).
   python scripts/own_operand_probe.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rangeldm_b200 import engine, models, _lib

CASES = [
    # B, W, H, C0, C1, Cout, ks, stride, up, norm, silu
    (8, 32, 2, 256, 0, 256, 3, 1, 1, True, True),
    (8, 32, 2, 256, 256, 256, 3, 1, 1, True, True),
    (8, 64, 4, 256, 0, 256, 3, 1, 1, True, True),
    (8, 64, 4, 256, 0, 768, 1, 1, 1, True, False),
    (8, 128, 8, 128, 0, 128, 3, 1, 1, True, True),
    (8, 128, 8, 256, 128, 128, 3, 1, 1, True, True),
]


def pair_moments(x):
    B, W, H, C = x.shape
    xd = x.double().reshape(B, W * H, C // 2, 2)
    return torch.stack([xd.sum((1, 3)), (xd * xd).sum((1, 3))], dim=-1).contiguous()


def graphed_us(pg, reps=20, inner=20):
    pg.run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            pg.run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * inner)


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    terms = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    L = _lib.lib()
    L.rldm_debug_conv_timestamps.argtypes = [ctypes.c_void_p]
    L.rldm_debug_conv_timestamps.restype = None
    stamps = torch.zeros(16, dtype=torch.int64, device=dev)
    for case in CASES:
        B, W, H, C0, C1, Cout, ks, stride, up, use_norm, silu = case
        g = torch.Generator().manual_seed(1)
        Cin = C0 + C1
        conv = models.LoRACompatibleConv(Cin, Cout, ks, stride=stride, padding=ks // 2).to(dev)
        conv.circular = True
        norm = torch.nn.GroupNorm(32, Cin, eps=1e-5).to(dev) if use_norm else None
        x0 = torch.randn(B, W, H, C0, generator=g).to(dev)
        x1 = torch.randn(B, W, H, C1, generator=g).to(dev) if C1 else None
        res = {}
        for own in (True, False):
            engine.FUSE_PREP = own
            pg = engine.Program(dev)
            bd = engine.Builder(pg, B, cache={}, terms_of=lambda w: terms)
            a0 = engine.Act(pg.hold(x0.clone()), B, W, H, C0, stats=pg.hold(pair_moments(x0)))
            a1 = engine.Act(pg.hold(x1.clone()), B, W, H, C1, stats=pg.hold(pair_moments(x1))) if C1 else None
            # three convolutions in a row off the same input: the middle one sees a conv before and after it
            for _ in range(3):
                opnd = bd.prep(a0, a1, norm, silu=silu, up=up, terms=terms, defer=True)
                bd.conv(opnd, W * up, H * up, conv, stats=True, terms=terms)
            bd.finish(); pg.finalize()
            n_ops = len(pg.ops)
            res[own] = graphed_us(pg) / 3
            if True:
                stamps.zero_()
                L.rldm_debug_conv_timestamps(ctypes.c_void_p(stamps.data_ptr()))
                pg.run(); torch.cuda.synchronize()
                L.rldm_debug_conv_timestamps(None)
                s = stamps.cpu().tolist()
                t0 = s[1]
                order = (("operand", 9), ("stage0", 2), ("last_mma", 3), ("accum", 4), ("staged", 6), ("cluster_sync", 7),
                         ("reduced", 8), ("end", 5))
                res[own, "rel"] = {k: (s[i] - t0) for k, i in order if s[i]}
        print(f"{case}: own-operand conv {res[True]:.2f} us   prep + conv {res[False]:.2f} us")
        print(f"    own : cycles since griddepcontrol.wait {res[True, 'rel']}")
        print(f"    prep: cycles since griddepcontrol.wait {res[False, 'rel']}")
