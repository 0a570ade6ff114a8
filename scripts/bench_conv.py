"""Micro-benchmark of rldm_conv_tc on the layer shapes of the C3 workload (GPU box).
   python scripts/bench_conv.py            # prints us / launch and TFLOP/s (algorithmic, 2*M*N*K)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rangeldm_b200 import _lib as L

SHAPES = [  # B, W, H, Cin, Cout, ks, stride, name
    (8, 256, 16, 128, 128, 3, 1, "unet L0 128->128"),
    (8, 256, 16, 256, 128, 3, 1, "unet L0 256->128 (up3)"),
    (8, 128, 8, 128, 128, 3, 1, "unet L1 128->128"),
    (8, 128, 8, 256, 128, 3, 1, "unet L1 256->128"),
    (8, 64, 4, 256, 256, 3, 1, "unet L2 256->256"),
    (8, 64, 4, 512, 256, 3, 1, "unet L2 512->256"),
    (8, 32, 2, 512, 256, 3, 1, "unet L3 512->256"),
    (8, 128, 8, 128, 384, 1, 1, "unet L1 qkv"),
    (8, 512, 32, 128, 128, 3, 1, "dec 128->128 @512x32"),
    (8, 1024, 64, 64, 64, 3, 1, "dec 64->64 @1024x64"),
]
L.lib()
def run(split_mode):
    for (B, W, H, Cin, Cout, ks, stride, name) in SHAPES:
        x = torch.randn(B, W + 2, H, Cin, device="cuda").half()
        xl = (torch.randn(B, W + 2, H, Cin, device="cuda") * 1e-3).half()
        planes = 2 if split_mode else 1
        w = (torch.randn(planes * ks * ks, Cout, Cin, device="cuda") * 0.02).half()
        bias = torch.randn(Cout, device="cuda")
        out = torch.empty(B, W // stride, H // stride, Cout, device="cuda")
        res = torch.randn_like(out)
        stats = torch.zeros(B, 32, 2, dtype=torch.float64, device="cuda")
        def call():
            L.call("rldm_conv_tc", L.ptr(x), L.ptr(xl) if split_mode else None, L.ptr(w), L.ptr(bias), None, 0, L.ptr(res),
                   L.ptr(out), B, W, H, Cin, Cout, ks, stride, 1 if ks == 3 else 0, 1, 0, None)
        for _ in range(3): call()
        torch.cuda.synchronize()
        # graph of n back-to-back launches: what the kernel costs inside a trajectory graph (no host work, PDL edges)
        n = 20
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n): call()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * n)
        fl = 2.0 * B * (W // stride) * (H // stride) * Cout * Cin * ks * ks
        print(f"  {name:28s} {us:8.1f} us   {fl / us / 1e6:7.1f} TFLOP/s algorithmic")
print("split-fp16 (3 MMA terms):"); run(True)
print("plain fp16 (1 MMA term):"); run(False)

# ---- per-CTA timeline of the per-tap kernel (clock64 stamps of CTA 0), top-level layer
import ctypes
lib = L.lib()
lib.rldm_debug_conv_timestamps.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.rldm_debug_conv_timestamps(buf.data_ptr())
for (B, W, H, Cin, Cout, ks, stride, name) in SHAPES[:3] + SHAPES[6:7] + SHAPES[8:]:
    x = torch.randn(B, W + 2, H, Cin, device="cuda").half(); xl = (x.float() * 1e-3).half()
    w = (torch.randn(2 * ks * ks, Cout, Cin, device="cuda") * 0.02).half()
    out = torch.empty(B, W // stride, H // stride, Cout, device="cuda"); res = torch.randn_like(out)
    for _ in range(3):
        L.call("rldm_conv_tc", L.ptr(x), L.ptr(xl), L.ptr(w), None, None, 0, L.ptr(res), L.ptr(out), B, W, H, Cin, Cout, ks,
               stride, 1 if ks == 3 else 0, 1, 0, None)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    d = [(t[i] - t[0]) for i in range(6)]
    print(f"  timeline {name:26s} cycles: prologue {d[1]}, first stage +{d[2]-d[1]}, mainloop issue +{d[3]-d[2]}, "
          f"acc ready +{d[4]-d[3]}, epilogue +{d[5]-d[4]} [tmem->smem {t[6]-t[4]}, sync {t[7]-t[6]}, rows {t[8]-t[7]}, tail {t[5]-t[8]}], total {d[5]}")
lib.rldm_debug_conv_timestamps(None)
