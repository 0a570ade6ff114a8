"""Debug helper (GPU box): run the same UNet program twice and report run-to-run differences per block."""
import os, sys
os.environ["RLDM_NOFREE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nets
from oracle.make_golden import TINY_UNET, seeded
from test_models_gpu import make_unet
from rangeldm_b200 import _lib

o = seeded(nets.OracleUNet2DModel, 4321, **TINY_UNET)
m = make_unet(TINY_UNET, o)
x = torch.randn(2, 5, 32, 8, generator=torch.Generator().manual_seed(0)).cuda()
plan = m.plan(2, 32, 8)
names = {id(mod): n for n, mod in m.named_modules()}
snaps = []
for rep in range(3):
    plan.run(x, 500)
    torch.cuda.synchronize()
    snaps.append([(names.get(id(mod), "?"), act.t.clone()) for mod, act in plan.prog.taps] + [("out", plan.out.clone())])
for (n, a), (_, b), (_, c) in zip(*snaps):
    d1 = (a - b).abs().max().item(); d2 = (a - c).abs().max().item()
    print(f"{n:40s} max|run0-run1| {d1:.3e}  max|run0-run2| {d2:.3e}  scale {a.abs().max().item():.3e}")
# op-level: re-run each op of the program in isolation twice and compare its output buffers is hard; instead
# run whole program op by op with syncs (serialised) and compare with the back-to-back run
lib = _lib.lib()
plan.run(x, 500); torch.cuda.synchronize()
ref = [act.t.clone() for mod, act in plan.prog.taps]
plan.x_in.copy_(x); plan.t_buf.fill_(500.0)
for op in plan.prog.ops:
    arr = (_lib.RldmOp * 1)(op)
    _lib.check(lib.rldm_run(arr, 1, _lib.stream_ptr()))
    torch.cuda.synchronize()
for (mod, act), r in zip(plan.prog.taps, ref):
    print(f"serialised vs back-to-back {names.get(id(mod), '?'):40s} {(act.t - r).abs().max().item():.3e}")
