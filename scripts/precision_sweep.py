"""GPU box: precision / throughput Pareto of the tensor-core operand formats.
For settings of (UNet levels 1..n, UNet full-resolution level, VAE) in {fp16x3, fp16x2, fp16}, against the fp32 CPU oracle
(metric max|a-b|/max|b|, C3 shapes, seeded random-init weights):
  * unet_forward  : ONE UNet forward (2 images, t = 500) -- nothing damps the operand rounding here;
  * latent / image: seed-matched 20-step DPM-Solver++ trajectory (4 images): final latent, decoded image;
  * decoder_n01   : the VAE decoder alone on an N(0,1) latent (its worst case: increments as large as the stream);
  * device time of one graphed UNet forward and one graphed KITTI decode at per-GPU batch 8.
    python scripts/precision_sweep.py
(Earlier sweeps of the round -- per VAE level, per VAE block, stream/branch split -- are profiles/precision_sweep_r2.json,
precision_blocks_r2.json and precision_pareto_r2.json, produced by this script at commits 683743f..b0c3f4e.)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import rangeldm_b200 as R
from rangeldm_b200 import engine
from oracle import nets, pipeline, schedulers
from oracle.make_golden import seeded
from test_models_gpu import make_unet, make_vae
from timeline import graphed_ms

NAME = {3: "fp16x3", 2: "fp16x2", 1: "fp16"}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


if __name__ == "__main__":
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C3)
    ov = seeded(nets.OracleAutoencoderKL, 1)
    u, v = make_unet(nets.UNET_C3, ou), make_vae(ov, [64, 128, 256], 2)
    NB = 4
    noise = torch.randn((NB, 4, 256, 16), generator=torch.Generator().manual_seed(0))
    ref_img, traj = pipeline.ldm_sample(ou, ov, schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"),
                                        noise, 20, return_latents=True)
    ref_lat = traj[-1]
    xf = torch.randn(2, 5, 256, 16, generator=torch.Generator().manual_seed(5))
    zd = torch.randn(2, 4, 256, 16, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ref_f = ou(xf, torch.tensor(500))
        ref_d = ov.decode(zd)
    out = []
    t_unet, t_dec, e_unet, e_dec = {}, {}, {}, {}
    for low in (3, 2, 1):
        for top in (3, 2, 1):
            engine.PRECISION, engine.PRECISION_TOP, engine.PRECISION_TOP_SAMPLER = low, top, top
            u.invalidate_plans()
            f = u(xf.cuda(), torch.tensor(500)).sample
            sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
            sch.set_timesteps(20)
            pe = R.pipelines.make_pos_encoding(NB, 256, 16, torch.device("cuda"))
            lat = R.FusedSampler(u, sch, None, NB, 1).run(noise.cuda(), pe)
            p = u.plan(8, 256, 16, 1)
            p.x_in.normal_(); p.t_buf.fill_(500.0)
            t_unet[(low, top)] = graphed_ms(p.prog) * 1e3
            e_unet[(low, top)] = (rel(f, ref_f), rel(lat, ref_lat), lat)
            u.invalidate_plans(); torch.cuda.empty_cache()
            print("unet", NAME[low], NAME[top], f"forward {e_unet[(low, top)][0]:.3e} latent {e_unet[(low, top)][1]:.3e} {t_unet[(low, top)]:.1f} us", flush=True)
    for vae in (3, 2, 1):
        engine.PRECISION_VAE = vae
        v.invalidate_plans()
        d = v.decode(zd.cuda()).sample
        d_lat = v.decode((ref_lat / ov.scaling_factor).cuda()).sample
        pl = v.decoder_plan(8, 256, 16)
        pl.z_in.normal_()
        t_dec[vae] = graphed_ms(pl.prog) * 1e3
        e_dec[vae] = (rel(d, ref_d), rel(d_lat, ref_img))
        v.invalidate_plans(); torch.cuda.empty_cache()
        print("vae", NAME[vae], f"decoder_n01 {e_dec[vae][0]:.3e} decoder_on_oracle_latent {e_dec[vae][1]:.3e} {t_dec[vae]:.1f} us", flush=True)
    # end-to-end image parity of every combination: decode each UNet setting's latent with each VAE setting
    for (low, top), (ef, el, lat) in e_unet.items():
        for vae in (3, 2, 1):
            engine.PRECISION_VAE = vae
            v.invalidate_plans()
            img = v.decode((lat / ov.scaling_factor)).sample
            row = {"unet_low": NAME[low], "unet_top": NAME[top], "vae": NAME[vae], "unet_forward_relerr": ef, "latent_relerr": el,
                   "image_relerr": rel(img, ref_img), "decoder_n01_relerr": e_dec[vae][0],
                   "unet_forward_us_b8": round(t_unet[(low, top)], 1), "decoder_us_b8": round(t_dec[vae], 1),
                   "step_ms_b8_est": round((20 * t_unet[(low, top)] + t_dec[vae]) / 1e3, 3)}
            out.append(row)
            print(json.dumps(row), flush=True)
    v.invalidate_plans()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precision_pareto2.json"), "w"), indent=1)
