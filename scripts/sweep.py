"""GPU box: whole-trajectory throughput of the C3 workload against the per-GPU batch, plus graphed UNet-only and
decoder-only times.   python scripts/sweep.py [batches...]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from rangeldm_b200.pipelines import FusedSampler, make_pos_encoding

dev = torch.device("cuda:0")
pipe = bench.build_pipeline(dev)
pipe.scheduler.set_timesteps(bench.STEPS)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def graphed(prog):
    prog.run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        prog.run()
    return g.replay


for B in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]:
    s = FusedSampler(pipe.unet, pipe.scheduler, pipe.vae, B, 1)
    noise = torch.randn(B, 4, 256, 16, device=dev)
    pos = make_pos_encoding(B, 256, 16, dev)
    s.load(noise, pos)
    ms = timed(s.replay)
    u = timed(graphed(s.plan.prog))
    d = timed(graphed(s.dec.prog))
    print(json.dumps({"batch": B, "ms": round(ms, 3), "images_per_s": round(B / ms * 1e3, 2), "unet_ms": round(u, 3),
                      "decoder_ms": round(d, 3), "unet_tflops": round(B * 34.07 / u, 1),
                      "decoder_tflops": round(B * 157.46 / d, 1)}), flush=True)
    del s
    torch.cuda.empty_cache()
