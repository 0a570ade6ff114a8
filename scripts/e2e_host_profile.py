"""GPU box: where does the host time of one `LDMPipelineRange.__call__` go (C3, batch 8)?  python scripts/e2e_host_profile.py"""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda:0")
pipe = bench.build_pipeline(dev)
pipe.set_progress_bar_config(disable=True)
g = torch.Generator().manual_seed(0)
host = torch.empty((8, 2, 1024, 64), pin_memory=True)
for _ in range(3):
    host.copy_(pipe(batch_size=8, generator=g, num_inference_steps=20, output_type="torch"))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    host.copy_(pipe(batch_size=8, generator=g, num_inference_steps=20, output_type="torch"))
torch.cuda.synchronize()
print(f"per call {(time.perf_counter() - t0) * 100:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    host.copy_(pipe(batch_size=8, generator=g, num_inference_steps=20, output_type="torch"))
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
