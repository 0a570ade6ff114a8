"""GPU box: serialised, cache-warm per-op durations of one C3 UNet forward and one KITTI decode, measured inside a
CUDA graph with `rldm_run_timed` (a %globaltimer stamp after every op), next to the real graphed time of the
same program.   python scripts/timeline.py [batch] [--ops]"""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rangeldm_b200 import _lib

NAMES = {1: "gn_stats", 2: "prep", 3: "conv_tc", 4: "conv_in", 5: "conv_out", 6: "attention", 7: "temb",
         8: "sched_step", 9: "memset", 10: "conv_ref", 11: "scale", 12: "norm_conv_out", 13: "fused_levels",
         14: "conv_up2"}


def describe(op):
    i = op.i
    if op.kind == 3:
        return f"conv B{i[1]} {i[2]}x{i[3]} {i[4]}->{i[5]} k{i[6]} s{i[7]}" + (" +res" if op.p[4] else "")
    if op.kind == 2:
        return f"prep {i[6]}x{i[7]} C{i[0]}+{i[1]} up{i[4]}" + (" +raw" if op.p[7] else "")
    if op.kind == 6:
        return f"attention N{i[1]} C{i[2]}"
    if op.kind == 14:
        return f"conv_up2 B{i[0]} {i[1]}x{i[2]} -> {2 * i[1]}x{2 * i[2]} {i[3]}->{i[4]} (4 phase launches)"
    if op.kind == 13:
        return f"fused run of {op.n} ops"
    return NAMES.get(op.kind, str(op.kind))


timed_profile = bench.timed_profile


def graphed_ms(prog, reps=10):
    prog.run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        prog.run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(args[0]) if args else 8
    dev = torch.device("cuda:0")
    pipe = bench.build_pipeline(dev)
    if "--c2" in sys.argv:          # RangeDM pixel UNet (BASELINE configs[1]), batch 1 unless given
        import rangeldm_b200 as R
        from configs_bench import UNET_C2
        torch.manual_seed(0)
        u2 = R.UNet2DModel(**UNET_C2)
        R.replace_down(u2); R.replace_conv(u2)
        u2 = u2.to(dev)
        B = int(args[0]) if args else 1
        uplan = u2.plan(B, 1024, 64, 1)
        uplan.x_in.normal_(); uplan.t_buf.fill_(500.0)
        plans = (("unet_c2", uplan),)
    else:
        uplan = pipe.unet.plan(B, 256, 16, 1, sampler=True)
        dplan = pipe.vae.decoder_plan(B, 256, 16)
        uplan.x_in.normal_(); uplan.t_buf.fill_(500.0); dplan.z_in.normal_()
        plans = (("unet", uplan), ("decoder", dplan))
    for name, plan in plans:
        prog = plan.prog
        us = timed_profile(prog)
        real = graphed_ms(prog)
        tot = collections.defaultdict(float); cnt = collections.Counter()
        for op, u in zip(prog.exec_ops, us):
            k = NAMES.get(op.kind, str(op.kind)); tot[k] += u; cnt[k] += 1
        print(f"{name}: batch {B}, {len(prog.ops)} ops in {len(prog.exec_ops)} launches-nodes; graphed {real * 1e3:.0f} us; serialised+stamped sum {sum(us):.0f} us")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            print(f"   {k:10s} n={cnt[k]:3d} total {v:8.1f} us  avg {v / cnt[k]:6.1f}")
        if "--ops" in sys.argv:
            for idx, (op, u) in enumerate(zip(prog.exec_ops, us)):
                print(f"   {idx:3d} {u:7.1f} us  {describe(op)}")
