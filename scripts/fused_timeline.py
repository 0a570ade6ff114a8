"""GPU box: per-phase device timeline of the fused-levels kernel (CTA 0): work time and grid-barrier time of every
phase of every fused run of one C3 UNet forward.   python scripts/fused_timeline.py [batch]"""
import os, sys, ctypes, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rangeldm_b200 import _lib

KN = {0: "prep", 1: "conv_main", 2: "conv_fin", 3: "attn"}
if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0")
    pipe = bench.build_pipeline(dev)
    plan = pipe.unet.plan(B, 256, 16, 1, sampler=True)
    plan.x_in.normal_(); plan.t_buf.fill_(500.0)
    lib = _lib.lib()
    lib.rldm_fused_debug.restype = ctypes.c_int
    lib.rldm_fused_debug.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.rldm_fused_phase_kind.restype = ctypes.c_int
    lib.rldm_fused_phase_kind.argtypes = [ctypes.c_void_p, ctypes.c_int]
    prog = plan.prog
    bufs = []
    for op, (i, j) in zip(prog.exec_ops, prog.exec_src):
        if op.kind == _lib.OP_FUSED:
            buf = torch.zeros(3 * 4096, dtype=torch.int64, device=dev)
            n = lib.rldm_fused_debug(op.p[0], buf.data_ptr())
            bufs.append((op, i, j, buf, n))
    for _ in range(3):
        prog.run()
    torch.cuda.synchronize()
    tot_w, tot_b, cnt = collections.defaultdict(float), collections.defaultdict(float), collections.Counter()
    for op, i, j, buf, n in bufs:
        t = buf[:3 * n].cpu().view(n, 3).double()
        work = (t[:, 1] - t[:, 0]) / 1e3
        barr = (t[:, 2] - t[:, 1]) / 1e3
        gap = (t[1:, 0] - t[:-1, 2]) / 1e3
        print(f"fused run ops[{i}:{j}]: {n} phases, total {(t[-1, 2] - t[0, 0]) / 1e3:.1f} us, work {work.sum():.1f}, barrier {barr.sum():.1f}, "
              f"descriptor/gap {gap.sum():.1f}")
        for k in range(n):
            kind = KN[lib.rldm_fused_phase_kind(op.p[0], k)]
            tot_w[kind] += work[k].item(); tot_b[kind] += barr[k].item(); cnt[kind] += 1
        if "--phases" in sys.argv:
            for k in range(min(n, 60)):
                print(f"    {k:3d} {KN[lib.rldm_fused_phase_kind(op.p[0], k)]:10s} work {work[k]:7.2f} us  barrier {barr[k]:6.2f} us")
    for kind in tot_w:
        print(f"{kind:10s} n={cnt[kind]:4d}  work avg {tot_w[kind] / cnt[kind]:6.2f} us  barrier avg {tot_b[kind] / cnt[kind]:6.2f} us")
    for op, i, j, buf, n in bufs:
        lib.rldm_fused_debug(op.p[0], None)
