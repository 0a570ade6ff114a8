"""GPU box: clock64 timeline of CTA 0 of the role-swapped persistent convolution (rldm_conv_tc on a top-level shape).
   python scripts/conv_wt_timeline.py [terms=3]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rangeldm_b200 import _lib as L
lib = L.lib()
lib.rldm_debug_conv_timestamps.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
TERMS = int(sys.argv[1]) if len(sys.argv) > 1 else 3
SHAPES = [(8, 256, 16, 128, 128, "unet L0 128->128"), (8, 256, 16, 256, 256, "dec 256->256 @256x16"),
          (8, 512, 32, 128, 128, "dec 128->128 @512x32")]
for with_res in (False, True):
    for (B, W, H, Cin, Cout, name) in SHAPES:
        x = torch.randn(B, W + 2, H, Cin, device="cuda").half(); xl = (x.float() * 1e-3).half()
        w = (torch.randn(18 if TERMS >= 2 else 9, Cout, Cin, device="cuda") * 0.02).half()
        out = torch.empty(B, W, H, Cout, device="cuda"); res = torch.randn_like(out)
        stats = torch.zeros(B, Cout // 2, 2, dtype=torch.float64, device="cuda")
        def call():
            L.call("rldm_conv_tc_ex", L.ptr(x), L.ptr(xl) if TERMS == 3 else None, L.ptr(w), None, None, 0,
                   L.ptr(res) if with_res else None, L.ptr(out), B, W, H, Cin, Cout, 3, 1, 1, 1, 0, L.ptr(stats),
                   None, None, None, 0, TERMS)
        for _ in range(3): call()
        torch.cuda.synchronize()
        lib.rldm_debug_conv_timestamps(buf.data_ptr())
        call(); torch.cuda.synchronize()
        lib.rldm_debug_conv_timestamps(None)
        t = buf.cpu().tolist(); d = [t[i] - t[0] for i in range(9)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): call()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:24s} res={int(with_res)} {e0.elapsed_time(e1) * 50:6.1f} us/launch | cycles from entry: setup {d[1]}, first stage "
              f"{d[2]}, last MMA issued {d[3]}, acc complete {d[4]}, first chunk in regs {d[6]}, stores issued {d[8]}, done {d[5]}")
