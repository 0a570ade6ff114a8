"""GPU box: precision / throughput Pareto of the tensor-core operand formats.
For every setting of (branch convolutions, stream convolutions) in {fp16x3, fp16x2, fp16} (engine.py: a STREAM
convolution carries the residual stream through its operand -- resamplers, conv2 with a folded 1x1 shortcut -- a BRANCH
convolution produces an increment that is added to the fp32 stream):
  * parity of the seed-matched C3 20-step DPM-Solver++ trajectory (NB images) against the fp32 CPU oracle:
    max|a-b|/max|b| of the final latent (UNet + scheduler only), of the decoded image, and of the decoder alone;
  * device time of one graphed UNet forward and one graphed KITTI decode at per-GPU batch 8.
    python scripts/precision_sweep.py [quick]
(The per-level and per-block sweeps that led to the stream/branch split are profiles/precision_sweep_r2.json and
profiles/precision_blocks_r2.json, produced by this script at commits 683743f..6210bb0 of the round.)"""
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
    quick = "quick" in sys.argv
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C3)
    ov = seeded(nets.OracleAutoencoderKL, 1)
    u, v = make_unet(nets.UNET_C3, ou), make_vae(ov, [64, 128, 256], 2)
    NB = 2 if quick else 4
    noise = torch.randn((NB, 4, 256, 16), generator=torch.Generator().manual_seed(0))
    ref_img, traj = pipeline.ldm_sample(ou, ov, schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"),
                                        noise, 20, return_latents=True)
    ref_lat = traj[-1]
    settings = [(3, 3), (1, 3), (2, 3), (1, 2), (2, 2), (1, 1), (1, (1, 3)), (1, (2, 3))]      # (branch, stream | (UNet stream, VAE stream))
    if quick:
        settings = [(3, 3), (1, 3), (1, 1)]
    out = []
    for branch, stream in settings:
        su, sv = stream if isinstance(stream, tuple) else (stream, stream)
        engine.PRECISION, engine.PRECISION_STREAM, engine.PRECISION_STREAM_VAE = branch, su, sv
        u.invalidate_plans(); v.invalidate_plans()
        sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
        sch.set_timesteps(20)
        pe = R.pipelines.make_pos_encoding(NB, 256, 16, torch.device("cuda"))
        lat = R.FusedSampler(u, sch, None, NB, 1).run(noise.cuda(), pe)
        img = R.FusedSampler(u, sch, v, NB, 1).run(noise.cuda(), pe)
        dec_only = v.decode((ref_lat / ov.scaling_factor).cuda()).sample      # decoder alone on the ORACLE latent
        row = {"branch": NAME[branch], "stream_unet": NAME[su], "stream_vae": NAME[sv], "latent_relerr": rel(lat, ref_lat),
               "image_relerr": rel(img, ref_img), "decoder_only_relerr": rel(dec_only, ref_img)}
        p = u.plan(8, 256, 16, 1)
        p.x_in.normal_(); p.t_buf.fill_(500.0)
        tu = graphed_ms(p.prog) * 1e3
        d = v.decoder_plan(8, 256, 16)
        d.z_in.normal_()
        td = graphed_ms(d.prog) * 1e3
        row.update(unet_forward_us_b8=round(tu, 1), decoder_us_b8=round(td, 1), step_ms_b8_est=round((20 * tu + td) / 1e3, 3))
        out.append(row)
        print(json.dumps(row), flush=True)
        u.invalidate_plans(); v.invalidate_plans()
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precision_pareto.json"), "w"), indent=1)
