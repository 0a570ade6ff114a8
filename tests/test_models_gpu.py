"""GPU parity tests of whole modules and pipelines against the CPU oracle and the golden vectors
produced by the reference's own code (tests/golden, oracle/make_golden.py).

Tolerance: BASELINE.json's north star -- max|a-b|/max|b| <= 1e-3 on seed-matched outputs; integer
timestep tables bit exact."""
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-3          # the north-star tolerance (BASELINE.json); the default precision policy lands 2-25x inside it (profiles/parity_r2.jsonl)
# Two runs of the SAME arithmetic in a different order (graph replay with atomics in another order, another batch
# tiling, per-step vs fused program) differ by ~1e-7 in fp32 -- and the branch convolutions round their activations to
# fp16, so such a difference occasionally flips one rounding (2^-11 of that element): run-to-run agreement of whole
# trajectories is ~1e-5..1e-4, not bit-level.
RUN_TO_RUN = 5e-4


@pytest.fixture
def x3(monkeypatch):
    """Every convolution at split-fp16 x3 (RLDM_PRECISION=RLDM_PRECISION_TOP=RLDM_PRECISION_VAE=fp16x3): for structural checks that
    need ~1e-6 agreement."""
    from rangeldm_b200 import engine
    monkeypatch.setattr(engine, "PRECISION", 3)
    monkeypatch.setattr(engine, "PRECISION_TOP", 3)
    monkeypatch.setattr(engine, "PRECISION_TOP_SAMPLER", 3)
    monkeypatch.setattr(engine, "PRECISION_VAE", 3)
    yield


def make_unet(cfg, oracle_net, dev="cuda"):
    import rangeldm_b200 as R
    u = R.UNet2DModel(**cfg)
    R.replace_down(u)
    R.replace_conv(u)
    u.load_state_dict(oracle_net.state_dict())
    return u.to(dev)


def make_vae(oracle_vae, boc, layers, dev="cuda"):
    import rangeldm_b200 as R
    n = len(boc)
    v = R.AutoencoderKL(in_channels=2, out_channels=2, down_block_types=["DownEncoderBlock2D"] * n,
                        up_block_types=["UpDecoderBlock2D"] * n, block_out_channels=boc, layers_per_block=layers,
                        latent_channels=4)
    v.quant_conv = torch.nn.Identity()
    v.post_quant_conv = torch.nn.Identity()
    R.replace_down(v)
    R.replace_conv(v)
    R.replace_attn(v)
    v.load_state_dict(oracle_vae.state_dict())
    return v.to(dev)


@pytest.fixture(scope="module")
def tiny():
    from oracle import nets
    from oracle.make_golden import TINY_UNET, TINY_UNET_PIXEL, TINY_VAE, seeded
    ou = seeded(nets.OracleUNet2DModel, 4321, **TINY_UNET)
    oup = seeded(nets.OracleUNet2DModel, 4322, **TINY_UNET_PIXEL)
    ov = seeded(nets.OracleAutoencoderKL, 1234, **TINY_VAE)
    return dict(ou=ou, oup=oup, ov=ov, u=make_unet(TINY_UNET, ou), up=make_unet(TINY_UNET_PIXEL, oup),
                v=make_vae(ov, [64, 128], 1))


def test_unet_forward_tiny(tiny):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 32, 8, generator=g)
    for t in (torch.tensor(940), 47, torch.tensor([3, 500])):
        with torch.no_grad():
            ref = tiny["ou"](x, t)
        out = tiny["u"](x.cuda(), t).sample
        assert out.shape == ref.shape
        assert relerr(out, ref, f"unet_tiny_t{t}") < TOL


def test_unet_forward_cuda_core_conv_path_agrees(tiny, x3):
    """Same program with the CUDA-core conv restatement: isolates tensor-core issues from precision."""
    from rangeldm_b200 import _lib, engine
    x = torch.randn(2, 5, 32, 8, generator=torch.Generator().manual_seed(1))
    tiny["u"].invalidate_plans()
    a = tiny["u"](x.cuda(), 500).sample
    engine.CONV_KIND = _lib.OP_CONV_REF
    try:
        tiny["u"].invalidate_plans()
        b = tiny["u"](x.cuda(), 500).sample
    finally:
        engine.CONV_KIND = _lib.OP_CONV_TC
        tiny["u"].invalidate_plans()
    assert relerr(a, b, "unet_tiny_tc_vs_cudacore") < 1e-4


def test_vae_decode_golden_from_reference_decoder(tiny, golden):
    g = golden("vae_decoder.pt")
    out = tiny["v"].decode(g["z"].cuda()).sample
    assert relerr(out, g["out"], "vae_decoder_golden") < TOL


def test_vae_encode_golden_from_reference_encoder(tiny, golden):
    g = golden("vae_encoder.pt")
    dist = tiny["v"].encode(g["x"].cuda()).latent_dist
    assert relerr(dist.parameters, g["out"], "vae_encoder_golden") < TOL
    s = dist.sample(generator=torch.Generator().manual_seed(0))
    assert s.shape == (2, 4, 32, 8)


def test_ldm_pipeline_golden_from_reference_pipeline(tiny, golden):
    """Our fused, graph-captured trajectory against the image the REFERENCE's ldm/pipelines.py loop
    produced on the oracle nets with the same generator seed."""
    import rangeldm_b200 as R
    g = golden("ldm_pipeline.pt")
    for name, sch in (("dpm", R.DPMSolverMultistepScheduler(timestep_spacing="leading")),
                      ("ddim", R.DDIMScheduler(clip_sample=False))):
        pipe = R.LDMPipelineRange(tiny["v"], tiny["u"], sch, pos_encoding=True)
        img = pipe(batch_size=2, generator=torch.Generator().manual_seed(3), num_inference_steps=5,
                   output_type="torch")
        assert relerr(img, g[f"ldm_{name}"], f"ldm_pipeline_golden_{name}") < TOL, name
        # replaying the captured graph with the same seed: only the double-precision GroupNorm-moment atomics
        # are order dependent (split-K is a deterministic cluster reduction)
        img2 = pipe(batch_size=2, generator=torch.Generator().manual_seed(3), num_inference_steps=5)
        assert relerr(img2, img, "graph_replay_stability") < RUN_TO_RUN


def test_pixel_pipeline_golden_from_reference_pipeline(tiny, golden):
    import rangeldm_b200 as R
    g = golden("ldm_pipeline.pt")
    pipe = R.DDIMPipelineRange(tiny["up"], R.DDPMScheduler(clip_sample=False), pos_encoding=True)
    assert isinstance(pipe.scheduler, R.DDIMScheduler)
    img = pipe(batch_size=2, generator=torch.Generator().manual_seed(3), num_inference_steps=5, output_type="torch")
    assert relerr(img, g["pixel_ddim"], "pixel_pipeline_golden") < TOL


def test_step_by_step_module_api_matches_fused(tiny):
    import rangeldm_b200 as R
    sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
    pipe = R.LDMPipelineRange(tiny["v"], tiny["u"], sch, pos_encoding=True)
    a = pipe(batch_size=1, generator=torch.Generator().manual_seed(5), num_inference_steps=4)
    imgs = pipe(batch_size=1, generator=torch.Generator().manual_seed(5), num_inference_steps=4, final_only=False)
    assert len(imgs) == 5
    assert relerr(imgs[-1], a, "stepwise_vs_fused") < RUN_TO_RUN


def test_upscale_pipeline_matches_oracle(tiny):
    import rangeldm_b200 as R
    from oracle import nets, pipeline, schedulers
    from oracle.make_golden import TINY_UNET, seeded
    cfg = dict(TINY_UNET, in_channels=12)
    ou = seeded(nets.OracleUNet2DModel, 77, **cfg)
    u = make_unet(cfg, ou)
    pipe = R.LDMUpscalePipelineRange(tiny["v"], u, R.DPMSolverMultistepScheduler(timestep_spacing="leading"))
    sparse = torch.randn(2, 2, 128, 8, generator=torch.Generator().manual_seed(8)).clamp(-1, 1)
    img = pipe(image=sparse.cuda(), condition_encoder=R.SparseRangeImageEncoder2(), batch_size=2,
               num_inference_steps=4, generator=torch.Generator().manual_seed(2))
    noise = torch.randn((2, 4, 32, 8), generator=torch.Generator().manual_seed(2))
    ref = pipeline.upscale_sample(ou, tiny["ov"], schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"),
                                  noise, pipeline.sparse_encoder2(sparse), 4)
    assert relerr(img, ref, "upscale_pipeline") < TOL
    with pytest.raises(ValueError):
        pipe(image=None)


def test_ddpm_stochastic_pipeline_matches_oracle_with_injected_noise(tiny):
    """`ldm/inference.py` drives LDMPipelineRange with a DDPMScheduler (stochastic): same noise stream."""
    import rangeldm_b200 as R
    from oracle import pipeline, schedulers
    n = 4
    pipe = R.LDMPipelineRange(tiny["v"], tiny["u"], R.DDPMScheduler(clip_sample=False), pos_encoding=True)
    gen = torch.Generator().manual_seed(21)
    img = pipe(batch_size=1, generator=gen, num_inference_steps=n)
    gen = torch.Generator().manual_seed(21)
    noise = torch.randn((1, 4, 32, 8), generator=gen)
    osch = schedulers.OracleDDPMScheduler()
    osch.set_timesteps(n)
    var = [torch.randn((1, 4, 32, 8), generator=gen) if int(t) > 0 else None for t in osch.timesteps]
    ref = pipeline.ldm_sample(tiny["ou"], tiny["ov"], osch, noise, n, pos_encoding=True, variance_noise=var)
    assert relerr(img, ref, "ddpm_pipeline") < TOL


def test_ddpm_pixel_pipeline_matches_reference_loop_with_injected_noise(tiny):
    """`DDPMPipelineRange.__call__` (`ldm/pipelines.py:34-117`): pixel-space ancestral sampling, NO pos-encoding
    (image channels == unet.in_channels), scheduler.step(..., generator=generator) draws in loop order.  Fused
    trajectory graph and the per-step fallback (long trajectories) against the oracle loop on the same noise stream."""
    import rangeldm_b200 as R
    from oracle import nets, schedulers
    from oracle.make_golden import TINY_UNET, seeded
    cfg = dict(TINY_UNET, in_channels=2, out_channels=2)
    ou = seeded(nets.OracleUNet2DModel, 91, **cfg)
    u = make_unet(cfg, ou)
    n = 6
    pipe = R.DDPMPipelineRange(u, R.DDPMScheduler(clip_sample=False))
    with pytest.raises(TypeError):
        R.DDPMPipelineRange(u, R.DDPMScheduler(clip_sample=False), pos_encoding=True)   # like the reference's __init__
    img = pipe(batch_size=2, generator=torch.Generator().manual_seed(33), num_inference_steps=n, output_type="torch")
    gen = torch.Generator().manual_seed(33)
    x = torch.randn((2, 2, 32, 8), generator=gen)
    osch = schedulers.OracleDDPMScheduler()
    osch.set_timesteps(n)
    with torch.no_grad():
        for t in osch.timesteps:
            eps = ou(x, t)
            z = torch.randn((2, 2, 32, 8), generator=gen) if int(t) > 0 else None
            x = osch.step(eps, t, x, variance_noise=z)
    assert relerr(img, x, "ddpm_pixel_pipeline") < TOL
    # per-step path (what 1000-step runs take): same stream, same result
    pipe.MAX_UNROLLED_STEPS = 1
    img2 = pipe(batch_size=2, generator=torch.Generator().manual_seed(33), num_inference_steps=n, output_type="torch")
    assert relerr(img2, x, "ddpm_pixel_pipeline_stepwise") < TOL
    out = pipe(batch_size=1, generator=torch.Generator().manual_seed(1), num_inference_steps=2, output_type="numpy")
    assert out.images.shape == (1, 32, 8, 2) and out.images.min() >= 0 and out.images.max() <= 1


def test_upscale_pipeline_inpainting_mask_branch_matches_oracle(tiny):
    """`LDMUpscalePipelineRange` with `mask=` (`ldm/pipelines.py:406-412`, `ldm/inference_conditional.py:137-152`):
    condition = vae.encode(masked image).latent_dist.sample() * scaling_factor ++ nearest-resized mask (5 channels,
    UNet in_channels 9).  The posterior sample draws on the device default generator: re-seeding it reproduces the draw."""
    import rangeldm_b200 as R
    from oracle import nets, pipeline, schedulers
    from oracle.make_golden import TINY_UNET, seeded
    cfg = dict(TINY_UNET, in_channels=9)
    ou = seeded(nets.OracleUNet2DModel, 78, **cfg)
    u = make_unet(cfg, ou)
    pipe = R.LDMUpscalePipelineRange(tiny["v"], u, R.DPMSolverMultistepScheduler(timestep_spacing="leading"))
    g = torch.Generator().manual_seed(12)
    masked = torch.randn(2, 2, 64, 16, generator=g).clamp(-1, 1)          # tiny VAE: 2x down -> (2,4,32,8)
    mask = (torch.rand(2, 1, 64, 16, generator=g) > 0.4).float()
    masked = masked * mask
    torch.cuda.manual_seed(555)
    img = pipe(image=masked, mask=mask, batch_size=2, num_inference_steps=4, generator=torch.Generator().manual_seed(2))
    torch.cuda.manual_seed(555)
    enc_noise = torch.randn((2, 4, 32, 8), device="cuda", dtype=torch.float32).cpu()    # the same device draw
    noise = torch.randn((2, 4, 32, 8), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        lat = tiny["ov"].encode_sample(masked, enc_noise) * tiny["ov"].scaling_factor
        cond = torch.cat([lat, torch.nn.functional.interpolate(mask, size=lat.shape[-2:])], dim=1)
        # the condition itself (our VAE encoder + DiagonalGaussianDistribution + nearest resize)
        torch.cuda.manual_seed(555)
        ours = pipe.encode_masked_image(masked, mask)
        assert ours.shape == (2, 5, 32, 8)
        assert relerr(ours, cond, "inpainting_condition") < TOL
        ref = pipeline.upscale_sample(ou, tiny["ov"], schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"),
                                      noise, cond, 4)
    assert relerr(img, ref, "upscale_pipeline_inpainting") < TOL
    with pytest.raises(AssertionError):          # channel contract of `ldm/pipelines.py:480`
        R.LDMUpscalePipelineRange(tiny["v"], tiny["u"], R.DDIMScheduler(clip_sample=False))(
            image=masked, mask=mask, batch_size=2, num_inference_steps=2)


def _trajectory_with_latents(u, v, noise, cond, steps):
    """Our per-step module API (unet(x, t).sample, scheduler.step) collecting every intermediate latent, plus the
    fused one-graph pipeline image for the same noise."""
    import rangeldm_b200 as R
    sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
    sch.set_timesteps(steps)
    lat = noise.cuda()
    lats = []
    for t in sch.timesteps:
        eps = u(torch.cat([lat, cond.cuda()], dim=1), t).sample
        lat = sch.step(eps, t, lat).prev_sample
        lats.append(lat.cpu())
    return lats


def test_c3_trajectory_batch_8_every_intermediate_latent():
    """SURVEY 8(d) parity protocol at the BENCH shape: C3, per-GPU batch 8, 20-step DPM-Solver++: EVERY intermediate
    latent of the per-step path and the final image of the fused graph against `oracle.pipeline.ldm_sample`."""
    import rangeldm_b200 as R
    from oracle import nets, pipeline, schedulers
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C3)
    ov = seeded(nets.OracleAutoencoderKL, 1)
    u, v = make_unet(nets.UNET_C3, ou), make_vae(ov, [64, 128, 256], 2)
    B = 8
    noise = torch.randn((B, 4, 256, 16), generator=torch.Generator().manual_seed(7))
    ref, traj = pipeline.ldm_sample(ou, ov, schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"),
                                    noise, 20, return_latents=True)
    lats = _trajectory_with_latents(u, v, noise, pipeline.pos_encoding_like(noise), 20)
    worst = max(relerr(a, b) for a, b in zip(lats, traj))
    relerr(lats[-1], traj[-1], "c3_b8_final_latent")
    assert worst < TOL, worst
    pipe = R.LDMPipelineRange(v, u, R.DPMSolverMultistepScheduler(timestep_spacing="leading"), pos_encoding=True)
    img = pipe(batch_size=B, generator=torch.Generator().manual_seed(7), num_inference_steps=20)
    assert img.shape == (B, 2, 1024, 64)
    assert relerr(img, ref, "c3_trajectory_batch8_fused_image") < TOL
    import json, os
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity.jsonl"), "a") as f:
        f.write(json.dumps({"name": "c3_b8_worst_intermediate_latent", "relerr": worst}) + "\n")


def test_c4_nuscenes_trajectory_batch_16():
    """BASELINE configs[3] at its per-GPU batch: nuScenes latent 4x256x8, 20-step DPM-Solver++, decode to (16,2,1024,32)."""
    import rangeldm_b200 as R
    from oracle import nets, pipeline, schedulers
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 3, **nets.UNET_C4)
    ov = seeded(nets.OracleAutoencoderKL, 2)
    u, v = make_unet(nets.UNET_C4, ou), make_vae(ov, [64, 128, 256], 2)
    pipe = R.LDMPipelineRange(v, u, R.DPMSolverMultistepScheduler(timestep_spacing="leading"), pos_encoding=True)
    img = pipe(batch_size=16, generator=torch.Generator().manual_seed(4), num_inference_steps=20)
    noise = torch.randn((16, 4, 256, 8), generator=torch.Generator().manual_seed(4))
    ref = pipeline.ldm_sample(ou, ov, schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"), noise, 20)
    assert img.shape == (16, 2, 1024, 32)
    assert relerr(img, ref, "c4_trajectory_batch16") < TOL


def test_c5_conditional_trajectory_batch_4():
    """BASELINE configs[4] at its per-GPU batch: sparse 16-beam condition -> SparseRangeImageEncoder2 (8 channels),
    12-channel UNet, 20-step DPM-Solver++, KITTI decode (`ldm/inference_conditional.py:125-170`)."""
    import rangeldm_b200 as R
    from oracle import nets, pipeline, schedulers
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 5, **nets.UNET_C5)
    ov = seeded(nets.OracleAutoencoderKL, 1)
    u, v = make_unet(nets.UNET_C5, ou), make_vae(ov, [64, 128, 256], 2)
    pipe = R.LDMUpscalePipelineRange(v, u, R.DPMSolverMultistepScheduler(timestep_spacing="leading"))
    sparse = torch.randn(4, 2, 1024, 16, generator=torch.Generator().manual_seed(8)).clamp(-1, 1)
    img = pipe(image=sparse.cuda(), condition_encoder=R.SparseRangeImageEncoder2(), batch_size=4,
               num_inference_steps=20, generator=torch.Generator().manual_seed(6))
    noise = torch.randn((4, 4, 256, 16), generator=torch.Generator().manual_seed(6))
    ref = pipeline.upscale_sample(ou, ov, schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"),
                                  noise, pipeline.sparse_encoder2(sparse), 20)
    assert img.shape == (4, 2, 1024, 64)
    assert relerr(img, ref, "c5_trajectory_batch4") < TOL


def test_c2_pixel_ddim_trajectory_10_steps_full_size():
    """BASELINE configs[1]: RangeDM pixel UNet (113.67 M params) on 2(+1)x1024x64, DDIM, batch 1 -- a 10-step
    seed-matched trajectory through `DDIMPipelineRange` (the 50-step run is the same loop; 10 keeps the CPU oracle
    at ~5 TFLOP)."""
    import rangeldm_b200 as R
    from oracle import nets, pipeline, schedulers
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C2)
    u = make_unet(nets.UNET_C2, ou)
    pipe = R.DDIMPipelineRange(u, R.DDPMScheduler(clip_sample=False), pos_encoding=True)
    img = pipe(batch_size=1, generator=torch.Generator().manual_seed(11), num_inference_steps=10, output_type="torch")
    noise = torch.randn((1, 2, 1024, 64), generator=torch.Generator().manual_seed(11))
    ref = pipeline.pixel_sample(ou, schedulers.OracleDDIMScheduler(), noise, 10, pos_encoding=True)
    assert relerr(img, ref, "c2_pixel_ddim_10step") < TOL


def test_c3_unet_full_size_one_forward():
    """BASELINE config C3 UNet (30.14 M params) on a (2,5,256,16) input against the fp32 oracle."""
    from oracle import nets
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C3)
    u = make_unet(nets.UNET_C3, ou)
    x = torch.randn(2, 5, 256, 16, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = ou(x, torch.tensor(500))
    out = u(x.cuda(), torch.tensor(500)).sample
    assert relerr(out, ref, "c3_unet_forward") < TOL


def test_c3_trajectory_20_step_dpm_solver_full_size():
    """Seed-matched 20-step DPM-Solver++ trajectory + KITTI VAE decode at full C3 size, batch 1."""
    import rangeldm_b200 as R
    from oracle import nets, pipeline, schedulers
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C3)
    ov = seeded(nets.OracleAutoencoderKL, 1)
    u = make_unet(nets.UNET_C3, ou)
    v = make_vae(ov, [64, 128, 256], 2)
    pipe = R.LDMPipelineRange(v, u, R.DPMSolverMultistepScheduler(timestep_spacing="leading"), pos_encoding=True)
    img = pipe(batch_size=1, generator=torch.Generator().manual_seed(0), num_inference_steps=20)
    noise = torch.randn((1, 4, 256, 16), generator=torch.Generator().manual_seed(0))
    ref = pipeline.ldm_sample(ou, ov, schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"), noise, 20)
    assert img.shape == (1, 2, 1024, 64)
    assert relerr(img, ref, "c3_trajectory_20step_dpm") < TOL


def test_c1_pixel_unet_forward_and_one_ddim_step_full_size():
    """BASELINE configs[0]: single UNet2DModel forward on one 2(+1)x1024x64 range image + 1 DDIM step
    (RangeDM pixel UNet, 113.67 M params, 6 resolution levels down to 32x2)."""
    import rangeldm_b200 as R
    from oracle import nets, schedulers
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C2)
    u = make_unet(nets.UNET_C2, ou)
    g = torch.Generator().manual_seed(0)
    img = torch.randn(1, 2, 1024, 64, generator=g)
    pe = torch.zeros(1, 1, 1024, 64)
    pe[:, :, 0, :] = 1
    x = torch.cat([img, pe], 1)
    osch, sch = schedulers.OracleDDIMScheduler(), R.DDIMScheduler(clip_sample=False)
    osch.set_timesteps(50)
    sch.set_timesteps(50)
    t = sch.timesteps[0]
    with torch.no_grad():
        eps_ref = ou(x, t)
    eps = u(x.cuda(), t).sample
    assert relerr(eps, eps_ref, "c1_pixel_unet_forward") < TOL
    nxt = sch.step(eps, t, img.cuda()).prev_sample
    assert relerr(nxt, osch.step(eps_ref, t, img), "c1_ddim_step") < TOL


def test_c4_nuscenes_latent_unet_forward():
    """BASELINE configs[3] geometry: nuScenes latent 256x8 (levels 8,4,2,1 beams: single-beam rows, 32-token
    attention on the CUDA-core kernel, GroupNorm statistics below the fused-epilogue size)."""
    from oracle import nets
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 3, **nets.UNET_C4)
    u = make_unet(nets.UNET_C4, ou)
    x = torch.randn(3, 5, 256, 8, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        ref = ou(x, torch.tensor(940))
    out = u(x.cuda(), torch.tensor(940)).sample
    assert relerr(out, ref, "c4_nuscenes_unet_forward") < TOL


def test_c5_conditional_unet_forward_full_size_batch_4():
    """BASELINE configs[4] geometry: conditional upsampling UNet (4 latent + 8 condition channels in, 12-channel
    conv_in) at 256x16, per-GPU batch 4 (`ldm/pipelines.py:498` concat each step)."""
    from oracle import nets
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 5, **nets.UNET_C5)
    u = make_unet(nets.UNET_C5, ou)
    x = torch.randn(4, 12, 256, 16, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ref = ou(x, torch.tensor(300))
    out = u(x.cuda(), torch.tensor(300)).sample
    assert relerr(out, ref, "c5_conditional_unet_forward") < TOL


def test_c4_nuscenes_decoder_full_size():
    """nuScenes VAE decode: (B,4,256,8) latent -> (B,2,1024,32) range image (BASELINE configs[3]); batch 3 exercises
    odd batch sizes on the persistent / two-tile kernels (M not a power of two)."""
    from oracle import nets
    from oracle.make_golden import seeded
    ov = seeded(nets.OracleAutoencoderKL, 2)
    v = make_vae(ov, [64, 128, 256], 2)
    z = torch.randn(3, 4, 256, 8, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        ref = ov.decode(z)
    out = v.decode(z.cuda()).sample
    assert out.shape == (3, 2, 1024, 32)
    assert relerr(out, ref, "c4_nuscenes_decoder") < TOL


@pytest.mark.parametrize("batch", [1, 3])
def test_c3_unet_odd_batches(batch):
    """Per-GPU batches other than 8: single image (every level under one wave) and 3 images (ragged tile counts)."""
    from oracle import nets
    from oracle.make_golden import seeded
    ou = seeded(nets.OracleUNet2DModel, 0, **nets.UNET_C3)
    u = make_unet(nets.UNET_C3, ou)
    x = torch.randn(batch, 5, 256, 16, generator=torch.Generator().manual_seed(10 + batch))
    t = torch.tensor([17, 500, 999][:batch])
    with torch.no_grad():
        ref = ou(x, t)
    out = u(x.cuda(), t).sample
    assert relerr(out, ref, f"c3_unet_batch{batch}") < TOL


def test_multi_stream_sampler_matches_single(tiny):
    """FusedSampler(streams=2): two sub-batch programs on parallel graph branches give the same images."""
    import rangeldm_b200 as R
    sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
    sch.set_timesteps(3)
    lat = torch.randn(4, 4, 32, 8, generator=torch.Generator().manual_seed(9)).cuda()
    pe = R.pipelines.make_pos_encoding(4, 32, 8, lat.device)
    a = R.FusedSampler(tiny["u"], sch, tiny["v"], 4, 1, streams=1).run(lat, pe)
    b = R.FusedSampler(tiny["u"], sch, tiny["v"], 4, 1, streams=2).run(lat, pe)
    assert relerr(b, a, "multi_stream_vs_single") < RUN_TO_RUN


def test_batch_sharding_is_index_stable(tiny):
    """Multi-GPU contract (`ldm/inference.py:159,174-183`): image of global index k depends only on
    its own seed -- generating it inside a batch of 2 or alone gives the same image."""
    import rangeldm_b200 as R
    pipe = R.LDMPipelineRange(tiny["v"], tiny["u"], R.DPMSolverMultistepScheduler(timestep_spacing="leading"),
                              pos_encoding=True)
    gens = [torch.Generator().manual_seed(100), torch.Generator().manual_seed(101)]
    both = pipe(batch_size=2, generator=gens, num_inference_steps=3)
    one = pipe(batch_size=1, generator=[torch.Generator().manual_seed(101)], num_inference_steps=3)
    assert relerr(both[1:], one, "batch_index_stability") < RUN_TO_RUN


def test_missing_library_fails_loudly(monkeypatch):
    from rangeldm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librldm.so")
    with pytest.raises(_lib.RldmError):
        _lib.lib()
