"""GPU box: per-node cost of chains of small kernels inside a CUDA graph, with and without programmatic dependent
launch (RLDM_PDL=1).   python scripts/launch_gap.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rangeldm_b200 import _lib as L
L.lib()
x = torch.randn(4096, device="cuda"); y = torch.empty_like(x)
B, W, H, C = 8, 32, 2, 256
xa = torch.randn(B, W, H, C, device="cuda")
hi = torch.empty(B, W + 2, H, C, dtype=torch.half, device="cuda"); lo = torch.empty_like(hi)
w = (torch.randn(2 * 9, C, C, device="cuda") * 0.02).half()
out = torch.empty(B, W, H, C, device="cuda")
def chain(kind, n):
    for _ in range(n):
        if kind == "scale":
            L.call("rldm_scale", L.ptr(x), 1.0001, L.ptr(y), x.numel())
        elif kind == "prep":
            L.call("rldm_prep", L.ptr(xa), C, None, 0, None, None, None, None, None, 0.0, 0, 0, 1, 1, L.ptr(hi), L.ptr(lo), None, None, B, W, H)
        elif kind == "conv":
            L.call("rldm_conv_tc", L.ptr(hi), L.ptr(lo), L.ptr(w), None, None, 0, None, L.ptr(out), B, W, H, C, C, 3, 1, 1, 1, 0, None)
        else:
            L.call("rldm_prep", L.ptr(xa), C, None, 0, None, None, None, None, None, 0.0, 0, 0, 1, 1, L.ptr(hi), L.ptr(lo), None, None, B, W, H)
            L.call("rldm_conv_tc", L.ptr(hi), L.ptr(lo), L.ptr(w), None, None, 0, None, L.ptr(out), B, W, H, C, C, 3, 1, 1, 1, 0, None)
for kind in ("scale", "prep", "conv", "prep+conv"):
    n = 100
    chain(kind, 2); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain(kind, n)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) * 1e3 / (5 * n)
    print(f"PDL={os.environ.get('RLDM_PDL', '0')} {kind:10s} {per:6.2f} us per graph node" + (" pair" if "+" in kind else ""))
