"""Debug helper (GPU box): compare every block output of the engine with the oracle's module outputs.
   RLDM_NOFREE=1 python scripts/debug_layers.py [unet|enc|dec]"""
import os, sys
os.environ["RLDM_NOFREE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rangeldm_b200 as R
from oracle import nets
from oracle.make_golden import TINY_UNET, TINY_VAE, seeded
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_models_gpu import make_unet, make_vae

which = sys.argv[1] if len(sys.argv) > 1 else "unet"
g = torch.Generator().manual_seed(0)
if which == "unet":
    o = seeded(nets.OracleUNet2DModel, 4321, **TINY_UNET)
    m = make_unet(TINY_UNET, o)
    x = torch.randn(2, 5, 32, 8, generator=g)
    plan = m.plan(2, 32, 8)
    omods, mmods = dict(o.named_modules()), dict(m.named_modules())
    run_o = lambda: o(x, torch.tensor(500))
    run_m = lambda: plan.run(x.cuda(), 500)
else:
    ov = seeded(nets.OracleAutoencoderKL, 1234, **TINY_VAE)
    v = make_vae(ov, [64, 128], 1)
    if which == "enc":
        x = torch.randn(2, 2, 64, 16, generator=g)
        plan = v._plan("enc", 2, 64, 16)
        run_o = lambda: ov.encode_moments(x)
    else:
        x = torch.randn(2, 4, 32, 8, generator=g)
        plan = v._plan("dec", 2, 32, 8)
        run_o = lambda: ov.decode(x)
    run_m = lambda: plan.run(x.cuda())
    omods, mmods = dict(ov.named_modules()), dict(v.named_modules())
outs = {}
hooks = [mod.register_forward_hook(lambda mod, i, out, n=n: outs.__setitem__(n, out)) for n, mod in omods.items()]
with torch.no_grad():
    ref = run_o()
got = run_m()
names = {id(mod): n for n, mod in mmods.items()}
for mod, act in plan.prog.taps:
    n = names.get(id(mod), "?")
    if n in outs:
        a = act.t.float().cpu().permute(0, 3, 1, 2)
        b = outs[n]
        print(f"{n:45s} {tuple(b.shape)}  rel {((a - b).abs().max() / b.abs().max()).item():.3e}")
print("final rel", ((got.cpu() - ref).abs().max() / ref.abs().max()).item())
