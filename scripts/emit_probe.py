"""GPU box: cost of a convolution that emits the next GroupNorm's operand (rldm_conv_tc_emit) against the same
convolution followed by a rldm_prep launch; clock64() stamps of CTA 0 of the emitting launch.
   python scripts/emit_probe.py [terms]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rangeldm_b200 import engine, models, _lib

CASES = [
    # B, W, H, Cin, Cout, ks, residual
    (8, 32, 2, 256, 256, 3, False),
    (8, 32, 2, 256, 256, 1, True),
    (8, 64, 4, 256, 256, 3, False),
    (8, 64, 4, 256, 256, 1, True),
    (8, 128, 8, 128, 128, 3, False),
    (8, 128, 8, 128, 128, 1, True),
]


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
    if len(sys.argv) > 2:                      # one case only (for ncu)
        CASES = [CASES[int(sys.argv[2])]]
    L = _lib.lib()
    L.rldm_debug_conv_timestamps.argtypes = [ctypes.c_void_p]
    L.rldm_debug_conv_timestamps.restype = None
    stamps = torch.zeros(16, dtype=torch.int64, device=dev)
    for case in CASES:
        B, W, H, Cin, Cout, ks, use_res = case
        g = torch.Generator().manual_seed(1)
        conv = models.LoRACompatibleConv(Cin, Cout, ks, padding=ks // 2).to(dev); conv.circular = True
        conv2 = models.LoRACompatibleConv(Cout, Cout, 3, padding=1).to(dev); conv2.circular = True
        norm = torch.nn.GroupNorm(32, Cout, eps=1e-5).to(dev)
        x0 = torch.randn(B, W, H, Cin, generator=g).to(dev)
        res = {}
        for emit in (True, False):
            engine.EMIT_PREP = emit
            pg = engine.Program(dev)
            bd = engine.Builder(pg, B, cache={}, terms_of=lambda w: terms)
            a0 = engine.Act(pg.hold(x0.clone()), B, W, H, Cin)
            opnd = bd.prep(a0, None, None, terms=terms)
            n_head = len(pg.ops)
            for _ in range(3):      # three (conv -> GroupNorm + SiLU -> conv) chains off the same input
                h = bd.conv(opnd, W, H, conv, residual=a0 if use_res and Cin == Cout else None, stats=True, terms=terms)
                o2 = bd.prep(h, None, norm, silu=True, terms=terms)
                bd.conv(o2, W, H, conv2, stats=True, terms=terms)
            bd.finish(); pg.finalize()
            res[emit] = graphed_us(pg) / 3
            res[emit, "ops"] = len(pg.ops)
            if emit:        # stamps of the first emitting conv alone
                sub = engine.Program(dev)
                for op in pg.ops[:n_head + 1]:
                    sub.append(op)
                sub.finalize()
                sub.run(); torch.cuda.synchronize()
                stamps.zero_()
                L.rldm_debug_conv_timestamps(ctypes.c_void_p(stamps.data_ptr()))
                sub.run(); torch.cuda.synchronize()
                L.rldm_debug_conv_timestamps(None)
                s = stamps.cpu().tolist()
                t0 = s[1]
                order = (("stage0", 2), ("last_mma", 3), ("accum", 4), ("staged", 6), ("cluster_sync", 7), ("reduced", 8),
                         ("stats", 5), ("published+sync", 10), ("gathered", 11), ("emitted", 12))
                rel = {k: (s[i] - t0) for k, i in order if s[i]}
        print(f"{case}: conv(emit) + conv {res[True]:.2f} us ({res[True, 'ops']} ops)   conv + prep + conv {res[False]:.2f} us ({res[False, 'ops']} ops)")
        print(f"    emitting conv, cycles since griddepcontrol.wait: {rel}")
