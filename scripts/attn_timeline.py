"""GPU box: clock64 timeline of CTA 0 of the tcgen05 attention kernel.  python scripts/attn_timeline.py [N] [C]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rangeldm_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B, H = 8, 8
lib = L.lib()
lib.rldm_debug_attn_timestamps.argtypes = [ctypes.c_void_p]
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
qkv = torch.randn(B, N, 3 * C, device="cuda")
out = torch.zeros(B, N // H + 2, H, C, dtype=torch.half, device="cuda"); lo = torch.zeros_like(out)
for _ in range(3):
    L.call("rldm_attention", L.ptr(qkv), L.ptr(out), L.ptr(lo), B, N, C, H)
torch.cuda.synchronize()
lib.rldm_debug_attn_timestamps(buf.data_ptr())
L.call("rldm_attention", L.ptr(qkv), L.ptr(out), L.ptr(lo), B, N, C, H)
torch.cuda.synchronize()
lib.rldm_debug_attn_timestamps(None)
t = buf.cpu().tolist(); t0 = t[127]; T = N // 128
rel = lambda i: t[i] - t0 if t[i] else None
print("setup done (after pdl_wait):", rel(124))
print("loader0: start", rel(0), "its tiles (0,3,6) ready:", [rel(1 + i) for i in range((T + 2) // 3)])
print("mma: QK issued:", [rel(32 + j) for j in range(min(T, 15))])
print("mma: PV issued:", [rel(48 + j) for j in range(min(T, 15))])
for k in range(T):
    print(f"softmax tile {k}: S ready {rel(64+4*k)}, max known {rel(65+4*k)}, P written {rel(67+4*k)}")
print("combine: all PV done", rel(125), "output written", rel(126))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    L.call("rldm_attention", L.ptr(qkv), L.ptr(out), L.ptr(lo), B, N, C, H)
e1.record(); torch.cuda.synchronize()
print("us per launch:", e0.elapsed_time(e1) * 50)
