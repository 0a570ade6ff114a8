"""Profiling driver (GPU box): one UNet forward and/or one VAE decode of the C3 workload, un-graphed, bracketed
by cudaProfilerStart/Stop so `ncu --profile-from-start off` sees exactly those launches.
   python scripts/profile_one.py [unet|dec|both] [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "both"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
pipe = bench.build_pipeline(dev)
uplan = pipe.unet.plan(B, 256, 16, 1, sampler=True)
dplan = pipe.vae.decoder_plan(B, 256, 16)
uplan.x_in.normal_(); uplan.t_buf.fill_(500.0); dplan.z_in.normal_()
for _ in range(2):
    uplan.prog.run(); dplan.prog.run()
torch.cuda.synchronize()
def timed(prog, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): prog.run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"un-graphed: unet {timed(uplan.prog):.3f} ms ({len(uplan.prog.ops)} ops), decoder {timed(dplan.prog):.3f} ms ({len(dplan.prog.ops)} ops), batch {B}")
torch.cuda.synchronize()
torch.cuda.profiler.start()
if which in ("unet", "both"):
    uplan.prog.run()
if which in ("dec", "both"):
    dplan.prog.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
