"""GPU box: precision / throughput Pareto of the tensor-core operand formats.
For every setting of (UNet lower levels, UNet full-resolution level, VAE decoder) in {fp16x3, fp16x2, fp16}:
  * parity of the seed-matched C3 20-step DPM-Solver++ trajectory (2 images) against the fp32 CPU oracle:
    max|a-b|/max|b| of the final latent (UNet + scheduler only) and of the decoded image;
  * device time of one graphed UNet forward and one graphed KITTI decode at per-GPU batch 8.
    python scripts/precision_sweep.py [quick]"""
import os, sys, json, itertools
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

class _N(dict):
    def __getitem__(self, k):
        return "/".join(dict.__getitem__(self, x) for x in k) if isinstance(k, tuple) else dict.__getitem__(self, k)


NAME = _N({3: "fp16x3", 2: "fp16x2", 1: "fp16"})


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


if __name__ == "__main__":
    quick = "quick" in sys.argv
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C3)
    ov = seeded(nets.OracleAutoencoderKL, 1)
    u, v = make_unet(nets.UNET_C3, ou), make_vae(ov, [64, 128, 256], 2)
    NB = 2
    noise = torch.randn((NB, 4, 256, 16), generator=torch.Generator().manual_seed(0))
    ref_img, traj = pipeline.ldm_sample(ou, ov, schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"),
                                        noise, 20, return_latents=True)
    ref_lat = traj[-1]
    settings = [(3, 3, 3), (3, 3, 2), (3, 3, 1), (3, 2, 3), (3, 1, 3), (2, 3, 3), (1, 3, 3), (2, 2, 3), (3, 2, 2), (3, 2, 1),
                (2, 2, 2), (2, 2, 1), (1, 1, 1)]
    if quick:
        settings = [(3, 3, 3), (3, 3, 1), (3, 2, 1)]
    if "dec" in sys.argv:      # per-level decoder sweep (latent 256x16 level, 512x32 level, 1024x64 level), UNet at fp16x2
        settings = [(2, 2, d) for d in ((3, 3, 3), (2, 3, 3), (3, 2, 3), (3, 3, 2), (1, 3, 3), (3, 1, 3), (3, 3, 1), (2, 2, 3),
                                        (1, 1, 3), (1, 2, 3), (2, 2, 2), (1, 1, 1))] + [(1, 1, (3, 3, 3)), (1, 1, (1, 1, 3))]
    if "blocks" in sys.argv:
        # decoder sensitivity by block (mid 0-1, up0 res 2-4 + upsampler 5, up1 res 6-8 + upsampler 9, up2 res 10-12):
        # everything plain fp16 except ONE block (or a tail of blocks) at fp16x3; decoder-only error on the oracle latent
        rows = []
        z = (ref_lat / ov.scaling_factor).cuda()

        def run(blocks, base):
            engine.PRECISION_DEC = [base]
            engine.PRECISION_BLOCKS.clear()
            engine.PRECISION_BLOCKS.update(blocks)
            v.invalidate_plans()
            e = rel(v.decode(z).sample, ref_img)
            d = v.decoder_plan(8, 256, 16)
            d.z_in.normal_()
            t = graphed_ms(d.prog) * 1e3
            v.invalidate_plans(); torch.cuda.empty_cache()
            return e, t
        for base, other in ((1, 3), (3, 1), (2, 3), (3, 2)):
            for k in [None] + list(range(13)):
                e, t = run({} if k is None else {k: other}, base)
                rows.append({"base": NAME[base], "block": k, "block_terms": NAME[other], "decoder_only_relerr": e, "decoder_us_b8": round(t, 1)})
                print(json.dumps(rows[-1]), flush=True)
        for n_tail in range(1, 8):
            e, t = run({k: 3 for k in range(13 - n_tail, 13)}, 1)
            rows.append({"base": "fp16", "tail_blocks_fp16x3": n_tail, "decoder_only_relerr": e, "decoder_us_b8": round(t, 1)})
            print(json.dumps(rows[-1]), flush=True)
        engine.PRECISION_BLOCKS.clear()
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "precision_blocks.json"), "w"), indent=1)
        sys.exit(0)
    out = []
    times = {}
    for low, top, dec in settings:
        engine.PRECISION, engine.PRECISION_TOP, engine.PRECISION_DEC = low, top, (list(dec) if isinstance(dec, tuple) else [dec])
        u.invalidate_plans(); v.invalidate_plans()
        # parity: final latent via the per-step API of a no-VAE pipeline is awkward; run the fused sampler without
        # and with the VAE
        sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
        sch.set_timesteps(20)
        pe = R.pipelines.make_pos_encoding(NB, 256, 16, torch.device("cuda"))
        lat = R.FusedSampler(u, sch, None, NB, 1).run(noise.cuda(), pe)
        img = R.FusedSampler(u, sch, v, NB, 1).run(noise.cuda(), pe)
        # decoder alone on the ORACLE latent: isolates the decoder's own error
        dec_only = v.decode((ref_lat / ov.scaling_factor).cuda()).sample
        row = {"unet_low": NAME[low], "unet_top": NAME[top], "vae": NAME[dec],
               "latent_relerr": rel(lat, ref_lat), "image_relerr": rel(img, ref_img),
               "decoder_only_relerr": rel(dec_only, ref_img)}
        # timing at batch 8 (cached per (low, top) and per dec)
        if (low, top) not in times:
            p = u.plan(8, 256, 16, 1)
            p.x_in.normal_(); p.t_buf.fill_(500.0)
            times[(low, top)] = graphed_ms(p.prog) * 1e3
        if ("dec", dec) not in times:
            d = v.decoder_plan(8, 256, 16)
            d.z_in.normal_()
            times[("dec", dec)] = graphed_ms(d.prog) * 1e3
        row["unet_forward_us_b8"] = round(times[(low, top)], 1)
        row["decoder_us_b8"] = round(times[("dec", dec)], 1)
        row["step_ms_b8_est"] = round((20 * times[(low, top)] + times[("dec", dec)]) / 1e3, 3)
        out.append(row)
        print(json.dumps(row), flush=True)
        u.invalidate_plans(); v.invalidate_plans()
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precision_sweep.json"), "w"), indent=1)
