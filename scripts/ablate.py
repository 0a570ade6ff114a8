"""GPU box: where the time of one C3 UNet forward really goes INSIDE the CUDA graph.  The stamped timeline
(scripts/timeline.py) serialises every op behind a stamp kernel (+~2 us each); here whole classes of ops are REMOVED
from the program and the remaining graph is timed (results are garbage, the timing is not): the difference to the
full graph is the true in-graph cost of the removed class, launch overlap included.
    python scripts/ablate.py [batch]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rangeldm_b200 import _lib
from rangeldm_b200._lib import RldmOp

K = _lib


def level_of(op):
    """azimuth width of the tensor an op works on (256 = top level ... 32 = level 3); None for level-less ops"""
    if op.kind in (K.OP_CONV_TC, K.OP_CONV_REF):
        return op.i[2] // op.i[7] if op.i[7] == 2 else op.i[2]          # stride-2 convs count for the level they produce
    if op.kind == K.OP_PREP:
        return op.i[6] * op.i[4]
    if op.kind == K.OP_ATTENTION:
        return op.i[1] // op.i[3]
    return None


def graph_time(ops, dev, reps=20):
    arr = (RldmOp * len(ops))(*ops)
    lib = _lib.lib()
    run = lambda: _lib.check(lib.rldm_run(arr, len(ops), _lib.stream_ptr()))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0")
    pipe = bench.build_pipeline(dev)
    res = {}
    for name, plan in (("unet", pipe.unet.plan(B, 256, 16, 1, sampler=True)), ("decoder", pipe.vae.decoder_plan(B, 256, 16))):
        if name == "unet":
            plan.x_in.normal_(); plan.t_buf.fill_(500.0)
        else:
            plan.z_in.normal_()
        ops = list(plan.prog.ops)
        full = graph_time(ops, dev)
        out = {"full_us": round(full, 1), "n_ops": len(ops)}
        cases = {"prep": lambda o: o.kind == K.OP_PREP, "attention": lambda o: o.kind == K.OP_ATTENTION,
                 "temb": lambda o: o.kind == K.OP_TEMB,
                 "conv_1x1": lambda o: o.kind == K.OP_CONV_TC and o.i[6] == 1,
                 "conv_3x3": lambda o: o.kind == K.OP_CONV_TC and o.i[6] == 3,
                 "conv_in_out": lambda o: o.kind in (K.OP_CONV_IN, K.OP_NORM_CONV_OUT)}
        levels = sorted({level_of(o) for o in ops if level_of(o) is not None}, reverse=True)
        for lv in levels:
            cases[f"level_{lv}_all"] = lambda o, lv=lv: level_of(o) == lv
            cases[f"level_{lv}_conv"] = lambda o, lv=lv: level_of(o) == lv and o.kind == K.OP_CONV_TC
            cases[f"level_{lv}_prep"] = lambda o, lv=lv: level_of(o) == lv and o.kind == K.OP_PREP
            cases[f"level_{lv}_attention"] = lambda o, lv=lv: level_of(o) == lv and o.kind == K.OP_ATTENTION
        for cname, pred in cases.items():
            kept = [o for o in ops if not pred(o)]
            removed = len(ops) - len(kept)
            if removed == 0:
                continue
            t = graph_time(kept, dev)
            out[cname] = {"removed_ops": removed, "saves_us": round(full - t, 1), "per_op_us": round((full - t) / removed, 2)}
        res[name] = out
        print(name, json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ablate.json"), "w"), indent=1)
