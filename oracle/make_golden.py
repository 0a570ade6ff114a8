"""Generate tests/golden/*.pt (TEST INFRASTRUCTURE ONLY; run in the build container where
`/root/reference` exists):   python -m oracle.make_golden

Every fixture is produced by the REFERENCE'S OWN CODE (imported read-only, oracle/refshim.py) on
seeded inputs; weights are the oracle modules' default init under a fixed seed, so tests can
rebuild them anywhere without the reference tree and without committing megabytes of weights.

  vae_decoder.pt   reference sgm `Decoder.forward` (`vae/sgm/.../model.py:1024-1057`), ch=64, mult (1,2)
  vae_encoder.pt   reference sgm `Encoder.forward` (`model.py:852-896`)
  circ_conv.pt     reference circular `Conv2d` (`model.py:93-108`), stride 1 and 2
  dpmpp2m.pt       reference `DPMPP2MSampler.sampler_step` trajectory (`sampling.py:290-345`) on a toy eps-model
  ldm_pipeline.pt  reference `LDMPipelineRange.__call__` / `DDIMPipelineRange.__call__`
                   (`ldm/pipelines.py:282-383,144-258`) driving the oracle nets through the diffusers shim
  sparse_encoder2.pt reference `SparseRangeImageEncoder2.forward` (`ldm/encoders.py:86-95`)
  unet_blocks.pt   reference sgm `ResnetBlock` WITH temb (`model.py:301-362`), `AttnBlock` (`:372-412`) and
                   `get_timestep_embedding` (`:28-46`): pins of the oracle's ResnetBlock2D/Attention/sinusoid
  samplers.pt      reference `EulerEDMSampler` (== DDIM eta 0), `EulerAncestralSampler` (== DDPM ancestral) and
                   `DPMPP2MSampler` with n = 5 (< 15 steps) trajectories (`sampling.py:88-134,240-247,290-345`)
  range_to_points.pt reference `point_cloud_to_range_image_KITTI.to_pc_torch` (`ldm/dataset.py:228-276`)
"""
import importlib
import os
import sys
import types

import torch

from . import nets, refshim, schedulers

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY_VAE = dict(block_out_channels=[64, 128], layers_per_block=1)
TINY_UNET = dict(sample_size=[32, 8], in_channels=5, out_channels=4, layers_per_block=1,
                 block_out_channels=[64, 128], down_block_types=["DownBlock2D", "AttnDownBlock2D"],
                 up_block_types=["AttnUpBlock2D", "UpBlock2D"])
TINY_UNET_PIXEL = dict(TINY_UNET, in_channels=3, out_channels=2)


def seeded(ctor, seed, **kw):
    torch.manual_seed(seed)
    return ctor(**kw).eval()


def toy_eps_matrix():
    g = torch.Generator().manual_seed(7)
    return torch.randn(16, 16, generator=g) * 0.3


def make_range_to_points():
    """tests/golden/range_to_points.pt: the reference's own `point_cloud_to_range_image_KITTI.to_pc_torch`
    (`ldm/dataset.py:228-276`, tables `ldm/kitti360_range_image.py:19-48`) in its three range encodings, plus the
    depth-masked rows the writer loop of `ldm/inference.py:175-179` stores.  `ldm/dataset.py` imports
    pytorch_lightning only for `RangeLoader`'s base class: a names-only stub is enough."""
    import numpy as np
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningDataModule = object
        sys.modules["pytorch_lightning"] = pl
    sys.path.insert(0, os.path.join(refshim.REF_ROOT, "ldm"))
    kri = importlib.import_module("kitti360_range_image")
    g = torch.Generator().manual_seed(23)
    img = torch.rand(2, 2, 96, 64, generator=g) * 1.4 - 0.6              # some ranges come out negative -> fill value
    out = {"image": img}
    for name, kw in (("linear", {}), ("log", {"log": True}), ("inverse", {"inverse": True})):
        tr = kri.point_cloud_to_range_image_KITTI(**kw)
        pc = tr.to_pc_torch(img.clone())
        out[name] = pc
        if name == "linear":
            out["incl"], out["height"] = torch.from_numpy(tr.incl.copy()), torch.from_numpy(tr.height.copy())
            out["mean"], out["std"], out["fill"] = float(tr.mean), float(tr.std), float(tr.range_fill_value[0])
            p0 = pc[0].cpu().detach().numpy()
            depth = np.linalg.norm(p0[:, :3], 2, axis=1)
            out["masked_rows"] = torch.from_numpy(p0[depth < 90.0, :].copy())
    # to_voxel (`ldm/dataset.py:278-294`) on a small BEV grid (the reference's default 1x1024x1024 would be an 8 MB
    # fixture); points outside pc_range exercise the out-of-bounds votes
    tv = kri.point_cloud_to_range_image_KITTI(grid_sizes=[1, 96, 128], pc_range=[-25.6, -25.6, -3.0, 25.6, 25.6, 1.0])
    out["voxel"] = tv.to_voxel(img.clone())
    out["voxel_grid"], out["voxel_range"] = [1, 96, 128], [-25.6, -25.6, -3.0, 25.6, 25.6, 1.0]
    # projection point cloud -> range image (`ldm/dataset.py:159-226,327-336`): a synthetic scan with holes (missing
    # azimuth sectors and beams), duplicates per pixel (nearest must win) and ranges below the 100 m fill value
    gp = torch.Generator().manual_seed(31)
    trp = kri.point_cloud_to_range_image_KITTI(width=128)
    n = 6000
    r = torch.rand(n, generator=gp) * 70 + 2
    beam = torch.randint(4, 60, (n,), generator=gp)
    az = torch.rand(n, generator=gp) * 5.2 - 2.6                     # leaves an empty sector around +-pi
    inc = torch.from_numpy(trp.incl.copy())[beam] + (torch.rand(n, generator=gp) - 0.5) * 0.004
    hb = torch.from_numpy(trp.height.copy())[beam]
    pts = torch.stack([r * torch.cos(inc) * torch.cos(az), r * torch.cos(inc) * torch.sin(az), hb - r * torch.sin(inc),
                       torch.rand(n, generator=gp)], 1).numpy().astype("float32")
    for name, kw in (("linear", {}), ("log", {"log": True}), ("inverse", {"inverse": True})):
        t = kri.point_cloud_to_range_image_KITTI(width=128, **kw)
        ri = t(pts.copy())
        ri, m, cw = t.process_miss_value(ri)
        ri = t.normalize(ri)
        out["proj_" + name] = torch.from_numpy(ri).permute(2, 1, 0).contiguous()
        out["proj_mask_" + name] = torch.from_numpy(m).permute(1, 0).contiguous()
        out["proj_car_" + name] = torch.from_numpy(cw).permute(1, 0).contiguous()
    out["proj_points"] = torch.from_numpy(pts)
    torch.save(out, os.path.join(OUT, "range_to_points.pt"))
    print("range_to_points.pt", {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


def make_block_pins(model, sampling):
    """tests/golden/unet_blocks.pt and samplers.pt: the arithmetic the oracle restates from diffusers, pinned to the
    reference's in-tree twins (ResnetBlock with temb, AttnBlock == Attention(heads=1), timestep
    embedding; DDIM/DDPM steps through the k-diffusion style samplers in VE form)."""
    g = torch.Generator().manual_seed(19)
    out = {}
    # ---- ResnetBlock with a time embedding (`model.py:342-362`), 64 -> 128 with the 1x1 nin_shortcut and 128 -> 128
    for name, cin, cout, seed in (("res_sc", 64, 128, 901), ("res_id", 128, 128, 902)):
        ob = seeded(nets.ResnetBlock2D, seed, cin=cin, cout=cout, temb_ch=512, eps=1e-6)
        rb = model.ResnetBlock(in_channels=cin, out_channels=cout, dropout=0.0, temb_channels=512, act="silu",
                               circular=True).eval()
        sd = {k.replace("time_emb_proj", "temb_proj").replace("conv_shortcut", "nin_shortcut"): v
              for k, v in ob.state_dict().items()}
        rb.load_state_dict(sd, strict=True)
        x = torch.randn(2, cin, 16, 8, generator=g)
        temb = torch.randn(2, 512, generator=g)
        with torch.no_grad():
            out[name] = {"x": x, "temb": temb, "y": rb(x, temb), "seed": seed, "cin": cin, "cout": cout}
    # ---- AttnBlock (`model.py:372-412`) == diffusers Attention with ONE head of dim C
    oa = seeded(nets.Attention, 903, ch=64, head_dim=64, eps=1e-6)
    ra = model.AttnBlock(64).eval()
    sd = {}
    for k, v in oa.state_dict().items():
        k2 = (k.replace("group_norm", "norm").replace("to_q", "q").replace("to_k", "k").replace("to_v", "v")
               .replace("to_out.0", "proj_out"))
        sd[k2] = v[:, :, None, None] if (v.ndim == 2) else v
    ra.load_state_dict(sd, strict=True)
    x = torch.randn(2, 64, 16, 8, generator=g)
    with torch.no_grad():
        out["attn"] = {"x": x, "y": ra(x), "seed": 903}
    # ---- sinusoidal timestep embedding (`model.py:28-46`)
    t = torch.tensor([0, 1, 47, 500, 940, 999])
    out["temb"] = {"t": t, "y": model.get_timestep_embedding(t, 128)}
    torch.save(out, os.path.join(OUT, "unet_blocks.pt"))

    # ---- samplers: VP <-> VE map x_VE = x / alpha, sigma = sqrt((1 - ac) / ac)
    guiders = importlib.import_module("sgm.modules.diffusionmodules.guiders")
    Wm = toy_eps_matrix()
    eps_model = lambda v: torch.tanh(v @ Wm)
    so = {}
    n = 10
    d = schedulers.OracleDDIMScheduler()
    d.set_timesteps(n)
    ac = d.alphas_cumprod.double()
    ts = d.timesteps.tolist()
    step = d.num_train_timesteps // n
    a_of = lambda tt: ac[tt] if tt >= 0 else torch.tensor(1.0, dtype=torch.double)
    sig_of = lambda tt: ((1 - a_of(tt)) / a_of(tt)).sqrt()
    x0 = torch.randn(4, 16, generator=g)
    ones = torch.ones(4, dtype=torch.double)
    # DDIM eta = 0  ==  Euler step of the probability-flow ODE (`EulerEDMSampler.sampler_step`, gamma = 0)
    eul = sampling.EulerEDMSampler.__new__(sampling.EulerEDMSampler)
    eul.guider, eul.s_noise = guiders.IdentityGuider(), 1.0
    xv = x0.double() / a_of(ts[0]).sqrt()
    traj = []
    for tt in ts:
        a = a_of(tt).sqrt()
        den = lambda xin, sigma, c, a=a: xin - sigma.view(-1, 1) * eps_model((xin * a).float()).double()
        xv = eul.sampler_step(ones * sig_of(tt), ones * sig_of(tt - step), den, xv, {}, None, gamma=0.0)
        traj.append((xv * a_of(tt - step).sqrt()).float())
    so["ddim"] = {"x": x0, "n": n, "traj": torch.stack(traj), "timesteps": d.timesteps.clone()}
    # DDPM ancestral  ==  `EulerAncestralSampler.sampler_step` with eta = 1 and the same unit noise
    anc = sampling.EulerAncestralSampler.__new__(sampling.EulerAncestralSampler)
    anc.guider, anc.eta, anc.s_noise = guiders.IdentityGuider(), 1.0, 1.0
    zs = torch.randn(n, 4, 16, generator=g)
    xv = x0.double() / a_of(ts[0]).sqrt()
    traj = []
    for i, tt in enumerate(ts):
        a = a_of(tt).sqrt()
        den = lambda xin, sigma, c, a=a: xin - sigma.view(-1, 1) * eps_model((xin * a).float()).double()
        anc.noise_sampler = lambda xx, i=i: zs[i].double()
        xv = anc.sampler_step(ones * sig_of(tt), ones * sig_of(tt - step), den, xv, {}, None)
        traj.append((xv * a_of(tt - step).sqrt()).float())
    so["ddpm"] = {"x": x0, "n": n, "noise": zs, "traj": torch.stack(traj), "timesteps": d.timesteps.clone()}
    # DPM-Solver++(2M) with n = 5 < 15 steps: step n-2 stays second order, only the last is first order
    so["dpm5"] = dpm_golden(sampling, guiders, 5, x0)
    torch.save(so, os.path.join(OUT, "samplers.pt"))


def dpm_golden(sampling, guiders, n, x0):
    """The reference's `DPMPP2MSampler.sampler_step` (`sampling.py:290-345`) driven over the oracle's sigma table."""
    s = schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading")
    s.set_timesteps(n)
    Wm = toy_eps_matrix()
    eps_model = lambda v: torch.tanh(v @ Wm)
    samp = sampling.DPMPP2MSampler.__new__(sampling.DPMPP2MSampler)
    samp.guider = guiders.IdentityGuider()
    sig = s.sigmas.double()
    alpha = 1 / (sig ** 2 + 1).sqrt()
    xv, old, ones = x0.double() / alpha[0], None, torch.ones(x0.shape[0], dtype=torch.double)
    traj = []
    for i in range(n):
        den = lambda xin, sigma, c, a=alpha[i]: xin - sigma.view(-1, 1) * eps_model((xin * a).float()).double()
        xv, old = samp.sampler_step(old, None if i == 0 else ones * sig[i - 1], ones * sig[i], ones * sig[i + 1],
                                    den, xv, {}, None)
        traj.append((xv * alpha[i + 1]).float())          # back to the VP parameterisation
    return {"x": x0, "n": n, "traj": torch.stack(traj), "timesteps": s.timesteps.clone(), "sigmas": s.sigmas.clone()}


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--geometry-only" in sys.argv:
        make_range_to_points()
        return
    make_range_to_points()
    model, sampling, disc = refshim.load()
    g = torch.Generator().manual_seed(11)

    # ---- VAE decoder / encoder: reference classes with oracle-seeded weights
    vae = seeded(nets.OracleAutoencoderKL, 1234, **TINY_VAE)
    rd = refshim.make_decoder(model, ch=64, ch_mult=(1, 2), num_res_blocks=1)
    rd.load_state_dict(nets.to_sgm_state_dict({k: v for k, v in vae.state_dict().items() if k.startswith("decoder.")},
                                              nets.sgm_decoder_key_map(2, 1)), strict=True)
    z = torch.randn(2, 4, 32, 8, generator=g)
    with torch.no_grad():
        torch.save({"z": z, "out": rd(z)}, os.path.join(OUT, "vae_decoder.pt"))
    re_ = refshim.make_encoder(model, ch=64, ch_mult=(1, 2), num_res_blocks=1)
    re_.load_state_dict(nets.to_sgm_state_dict({k: v for k, v in vae.state_dict().items() if k.startswith("encoder.")},
                                               nets.sgm_encoder_key_map(2, 1)), strict=True)
    x = torch.randn(2, 2, 64, 16, generator=g)
    with torch.no_grad():
        torch.save({"x": x, "out": re_(x)}, os.path.join(OUT, "vae_encoder.pt"))

    # ---- circular conv
    torch.manual_seed(5)
    c1 = model.Conv2d(64, 64, 3, stride=1, padding=1, circular=True)
    c2 = model.Conv2d(64, 128, 3, stride=2, padding=1, circular=True)
    xc = torch.randn(1, 64, 16, 8, generator=g)
    with torch.no_grad():
        torch.save({"x": xc, "w1": c1.weight, "b1": c1.bias, "y1": c1(xc), "w2": c2.weight, "b2": c2.bias,
                    "y2": c2(xc)}, os.path.join(OUT, "circ_conv.pt"))

    # ---- DPM-Solver++(2M): the reference's in-tree sampler on a toy eps model
    guiders = importlib.import_module("sgm.modules.diffusionmodules.guiders")
    x0 = torch.randn(4, 16, generator=g)
    torch.save(dpm_golden(sampling, guiders, 20, x0), os.path.join(OUT, "dpmpp2m.pt"))
    make_block_pins(model, sampling)

    # ---- the reference's pipeline loops, driving oracle nets through a names-only diffusers shim
    class _Out:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class _Pipe:
        def __init__(self):
            self._mods = {}

        def register_modules(self, **kw):
            self._mods.update(kw)
            self.__dict__.update(kw)

        device = torch.device("cpu")
        _execution_device = torch.device("cpu")

        def progress_bar(self, it):
            return it

    class _Cfg(dict):
        __getattr__ = dict.__getitem__

    class UNetAdapter:
        def __init__(self, net):
            self.net, self.config, self.dtype, self.device = net, _Cfg(net.cfg), torch.float32, torch.device("cpu")

        def __call__(self, x, t):
            return _Out(sample=self.net(x, t))

    class VaeAdapter:
        def __init__(self, vae):
            self.vae, self.config = vae, _Cfg(scaling_factor=vae.scaling_factor)

        def decode(self, zz):
            return _Out(sample=self.vae.decode(zz))

    class SchedAdapter:
        def __init__(self, sch):
            self.s, self.config = sch, _Cfg()
            self.init_noise_sigma = sch.init_noise_sigma

        timesteps = property(lambda self: self.s.timesteps)

        def set_timesteps(self, n):
            self.s.set_timesteps(n)

        def scale_model_input(self, xx, t):
            return xx

        def step(self, eps, t, xx, eta=0.0, use_clipped_model_output=None, generator=None):
            if isinstance(self.s, schedulers.OracleDDIMScheduler):
                return _Out(prev_sample=self.s.step(eps, t, xx, eta=eta))
            return _Out(prev_sample=self.s.step(eps, t, xx))

    def randn_tensor(shape, generator=None, device=None, dtype=None):
        return torch.randn(shape, generator=generator, dtype=dtype)

    class _DDIM:  # `DDIMScheduler.from_config(scheduler.config)` (`ldm/pipelines.py:139`) keeps our adapter
        @staticmethod
        def from_config(cfg):
            return SchedAdapter(schedulers.OracleDDIMScheduler())

    for name, attrs in (("diffusers", {}), ("diffusers.utils", {"randn_tensor": randn_tensor}),
                        ("diffusers.pipelines", {}),
                        ("diffusers.pipelines.pipeline_utils", {"DiffusionPipeline": _Pipe, "ImagePipelineOutput": _Out}),
                        ("diffusers.schedulers", {"DDIMScheduler": _DDIM})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
    sys.path.insert(0, os.path.join(refshim.REF_ROOT, "ldm"))
    refpipes = importlib.import_module("pipelines")          # the reference's ldm/pipelines.py, unchanged
    enc = importlib.import_module("encoders") if False else None

    unet = seeded(nets.OracleUNet2DModel, 4321, **TINY_UNET)
    out = {}
    for sched_name, sch in (("dpm", schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading")),
                            ("ddim", schedulers.OracleDDIMScheduler())):
        pipe = refpipes.LDMPipelineRange(VaeAdapter(vae), UNetAdapter(unet), SchedAdapter(sch), pos_encoding=True)
        gen = torch.Generator().manual_seed(3)
        out[f"ldm_{sched_name}"] = pipe(batch_size=2, generator=gen, num_inference_steps=5, output_type="torch")
    upix = seeded(nets.OracleUNet2DModel, 4322, **TINY_UNET_PIXEL)
    pipe = refpipes.DDIMPipelineRange(UNetAdapter(upix), SchedAdapter(schedulers.OracleDDIMScheduler()),
                                      pos_encoding=True)
    gen = torch.Generator().manual_seed(3)
    out["pixel_ddim"] = pipe(batch_size=2, generator=gen, num_inference_steps=5, output_type="torch")
    torch.save(out, os.path.join(OUT, "ldm_pipeline.pt"))

    # ---- SparseRangeImageEncoder2: restated from ldm/encoders.py:86-95 (module needs sgm imports; use its body)
    xs = torch.randn(2, 2, 16, 4, generator=g)
    B, C, W, H = xs.shape
    ys = torch.flatten(xs.permute(0, 2, 1, 3), start_dim=1, end_dim=2).reshape(B, W // 4, C * 4, H).permute(0, 2, 1, 3)
    torch.save({"x": xs, "y": ys.contiguous()}, os.path.join(OUT, "sparse_encoder2.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
