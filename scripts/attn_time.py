"""GPU box: device-timed launches of rldm_attention (L2-warm, back to back), with the split-fp16 output pair (P carried
as hi+lo in TMEM) and with a single-plane output (P as one fp16 plane).  python scripts/attn_time.py [N] [C] [B]
RLDM_ATTN_MMASYNC=1 selects the mma.sync kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rangeldm_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
H = 8
qkv = torch.randn(B, N, 3 * C, device="cuda")
out = torch.zeros(B, N // H + 2, H, C, dtype=torch.half, device="cuda"); lo = torch.zeros_like(out)
for name, lo_ptr in (("hi+lo", L.ptr(lo)), ("hi only", None)):
    for _ in range(5):
        L.call("rldm_attention", L.ptr(qkv), L.ptr(out), lo_ptr, B, N, C, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        for _ in range(20):
            L.call("rldm_attention", L.ptr(qkv), L.ptr(out), lo_ptr, B, N, C, H)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 50)
    print(f"N={N} C={C} B={B} output {name}: {best:.1f} us per launch")
