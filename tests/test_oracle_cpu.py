"""CPU tests: the oracle against the reference's own code (golden vectors in tests/golden generated
by oracle/make_golden.py; live comparison when /root/reference is present)."""
import numpy as np
import pytest
import torch

from conftest import relerr
from oracle import nets, pipeline, refshim, schedulers
from oracle.make_golden import TINY_UNET, TINY_UNET_PIXEL, TINY_VAE, seeded, toy_eps_matrix


def test_param_counts_match_survey_anchors():
    # 113.67 M (RangeDM), 30.14 M (RangeLDM), VAE enc 5.34 M + dec 7.99 M
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(nets.OracleUNet2DModel(**nets.UNET_C3)) == 30135684
    assert n(nets.OracleUNet2DModel(**nets.UNET_C2)) == 113672066
    vae = nets.OracleAutoencoderKL()
    assert n(vae.decoder) == 7989570 and n(vae.encoder) == 5341320


def test_vae_decoder_matches_reference_golden(golden):
    g = golden("vae_decoder.pt")
    vae = seeded(nets.OracleAutoencoderKL, 1234, **TINY_VAE)
    with torch.no_grad():
        assert relerr(vae.decode(g["z"]), g["out"]) < 1e-5


def test_vae_encoder_matches_reference_golden(golden):
    g = golden("vae_encoder.pt")
    vae = seeded(nets.OracleAutoencoderKL, 1234, **TINY_VAE)
    with torch.no_grad():
        assert relerr(vae.encode_moments(g["x"]), g["out"]) < 1e-5


def test_circular_conv_matches_reference_golden(golden):
    g = golden("circ_conv.pt")
    assert relerr(nets.circ_conv2d(g["x"], g["w1"], g["b1"], 1, 1), g["y1"]) < 1e-6
    assert relerr(nets.circ_conv2d(g["x"], g["w2"], g["b2"], 2, 1), g["y2"]) < 1e-6
    # the wrap is on dim 2 only: rolling the input along W rolls the output, along H it does not
    y = nets.circ_conv2d(g["x"], g["w1"], g["b1"], 1, 1)
    yr = nets.circ_conv2d(torch.roll(g["x"], 3, dims=2), g["w1"], g["b1"], 1, 1)
    assert torch.allclose(torch.roll(y, 3, dims=2), yr, atol=1e-5)


def test_dpmpp2m_matches_reference_sampler_golden(golden):
    g = golden("dpmpp2m.pt")
    s = schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading")
    s.set_timesteps(20)
    assert torch.equal(s.timesteps, g["timesteps"])          # integer table: bit exact
    assert torch.equal(s.sigmas, g["sigmas"])
    Wm = toy_eps_matrix()
    x = g["x"].clone()
    for i, t in enumerate(s.timesteps):
        x = s.step(torch.tanh(x @ Wm), t, x)
        assert relerr(x, g["traj"][i]) < 5e-6, i


def test_dpmpp2m_fewer_than_15_steps_matches_reference_sampler_golden(golden):
    """n = 5 < 15: with solver_order 2 diffusers keeps step n-2 SECOND order (`lower_order_second` only demotes a
    third-order solver) and makes only the final step first order -- what the in-tree DPMPP2MSampler does."""
    g = golden("samplers.pt")["dpm5"]
    s = schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading")
    s.set_timesteps(5)
    assert torch.equal(s.timesteps, g["timesteps"]) and torch.equal(s.sigmas, g["sigmas"])
    Wm = toy_eps_matrix()
    x = g["x"].clone()
    for i, t in enumerate(s.timesteps):
        x = s.step(torch.tanh(x @ Wm), t, x)
        assert relerr(x, g["traj"][i]) < 5e-6, i


def test_ddim_step_matches_reference_euler_sampler_golden(golden):
    """DDIM(eta=0) == the Euler step of the probability-flow ODE: pinned to the reference's `EulerEDMSampler`
    (`vae/sgm/modules/diffusionmodules/sampling.py:88-134`) through x_VE = x / sqrt(alphas_cumprod)."""
    g = golden("samplers.pt")["ddim"]
    d = schedulers.OracleDDIMScheduler()
    d.set_timesteps(g["n"])
    assert torch.equal(d.timesteps, g["timesteps"])
    Wm = toy_eps_matrix()
    x = g["x"].clone()
    for i, t in enumerate(d.timesteps):
        x = d.step(torch.tanh(x @ Wm), t, x)
        assert relerr(x, g["traj"][i]) < 5e-6, i


def test_ddpm_step_matches_reference_ancestral_sampler_golden(golden):
    """DDPM ancestral step (posterior mean + fixed_small variance) == `EulerAncestralSampler` with eta = 1
    (`sampling.py:136-155,240-247`, `sampling_utils.py:27-37`) on the same unit noise."""
    g = golden("samplers.pt")["ddpm"]
    d = schedulers.OracleDDPMScheduler()
    d.set_timesteps(g["n"])
    Wm = toy_eps_matrix()
    x = g["x"].clone()
    for i, t in enumerate(d.timesteps):
        x = d.step(torch.tanh(x @ Wm), t, x, variance_noise=g["noise"][i])
        assert relerr(x, g["traj"][i]) < 5e-6, i


def test_resnet_block_with_temb_matches_reference_golden(golden):
    """oracle ResnetBlock2D WITH the time-embedding projection == sgm `ResnetBlock(temb_channels=512)`
    (`vae/sgm/modules/diffusionmodules/model.py:301-362`), with and without the 1x1 shortcut."""
    g = golden("unet_blocks.pt")
    for name in ("res_sc", "res_id"):
        d = g[name]
        rb = seeded(nets.ResnetBlock2D, d["seed"], cin=d["cin"], cout=d["cout"], temb_ch=512, eps=1e-6)
        with torch.no_grad():
            assert relerr(rb(d["x"], d["temb"]), d["y"]) < 5e-6, name


def test_attention_single_head_matches_reference_attn_block_golden(golden):
    """oracle Attention (GN -> q,k,v -> softmax(QK^T/sqrt(d))V -> out + residual) with one head of dim C == sgm
    `AttnBlock` (`model.py:372-412`)."""
    d = golden("unet_blocks.pt")["attn"]
    at = seeded(nets.Attention, d["seed"], ch=64, head_dim=64, eps=1e-6)
    with torch.no_grad():
        assert relerr(at(d["x"]), d["y"]) < 5e-6


def test_timestep_embedding_matches_reference_golden(golden):
    """`sinusoidal_timestep` == sgm `get_timestep_embedding` (`model.py:28-46`) under (flip_sin_to_cos=False,
    freq_shift=1); the UNet2DModel flavour differs only by those two documented knobs (cos first, divisor half)."""
    d = golden("unet_blocks.pt")["temb"]
    y = nets.sinusoidal_timestep(d["t"], 128, flip_sin_to_cos=False, freq_shift=1)
    # arguments reach 999 rad, where one fp32 ulp is 6e-5: the two orders of forming `t * exp(-ln(1e4) i / d)` differ by that
    assert relerr(y, d["y"]) < 1e-4
    u = nets.sinusoidal_timestep(d["t"], 128)
    half = 64
    f = torch.exp(-np.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    a = d["t"].float()[:, None] * f[None]
    assert torch.allclose(u[:, :half], torch.cos(a), atol=1e-6) and torch.allclose(u[:, half:], torch.sin(a), atol=1e-6)


def test_scheduler_timestep_tables():
    # SURVEY.md App. A.4
    d = schedulers.OracleDDIMScheduler(); d.set_timesteps(50)
    assert d.timesteps.tolist() == list(range(980, -1, -20))
    s = schedulers.OracleDPMSolverMultistepScheduler(); s.set_timesteps(20)
    assert s.timesteps.tolist() == [999, 949, 899, 849, 799, 749, 699, 649, 599, 549, 500, 450, 400, 350, 300,
                                    250, 200, 150, 100, 50]
    s = schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading"); s.set_timesteps(20)
    assert s.timesteps.tolist() == list(range(940, 0, -47))


def test_pipeline_loops_match_reference_pipelines_golden(golden):
    """oracle/pipeline.py against the reference's ldm/pipelines.py run on the same oracle nets."""
    g = golden("ldm_pipeline.pt")
    vae = seeded(nets.OracleAutoencoderKL, 1234, **TINY_VAE)
    unet = seeded(nets.OracleUNet2DModel, 4321, **TINY_UNET)
    for name, sch in (("dpm", schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading")),
                      ("ddim", schedulers.OracleDDIMScheduler())):
        noise = torch.randn((2, 4, 32, 8), generator=torch.Generator().manual_seed(3))
        img = pipeline.ldm_sample(unet, vae, sch, noise, 5, pos_encoding=True)
        assert relerr(img, g[f"ldm_{name}"]) < 1e-5, name
    upix = seeded(nets.OracleUNet2DModel, 4322, **TINY_UNET_PIXEL)
    noise = torch.randn((2, 2, 32, 8), generator=torch.Generator().manual_seed(3))
    img = pipeline.pixel_sample(upix, schedulers.OracleDDIMScheduler(), noise, 5, pos_encoding=True)
    assert relerr(img, g["pixel_ddim"]) < 1e-5


def test_sparse_encoder2_golden(golden):
    g = golden("sparse_encoder2.pt")
    assert torch.equal(pipeline.sparse_encoder2(g["x"]), g["y"])


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_full_size_decoder_matches_live_reference():
    model, _, _ = refshim.load()
    vae = seeded(nets.OracleAutoencoderKL, 1)
    rd = refshim.make_decoder(model)
    rd.load_state_dict(nets.to_sgm_state_dict({k: v for k, v in vae.state_dict().items() if k.startswith("decoder.")},
                                              nets.sgm_decoder_key_map()), strict=True)
    z = torch.randn(1, 4, 64, 16, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        assert relerr(vae.decode(z), rd(z)) < 1e-5


def test_ddpm_ddim_consistency():
    """DDIM(eta=1) and DDPM share the posterior mean; DDIM(eta=0) x0 prediction inverts add-noise."""
    d = schedulers.OracleDDIMScheduler(); d.set_timesteps(50)
    x0 = torch.randn(2, 4, 8, 4)
    eps = torch.randn_like(x0)
    t = 500
    a = d.alphas_cumprod[t]
    xt = a.sqrt() * x0 + (1 - a).sqrt() * eps
    prev = d.step(eps, t, xt)
    ap = d.alphas_cumprod[t - 20]
    assert torch.allclose(prev, ap.sqrt() * x0 + (1 - ap).sqrt() * eps, atol=1e-5)


def test_geometry_oracle_matches_reference_to_pc_torch(golden):
    """oracle/geometry.py == the reference's `point_cloud_to_range_image_KITTI.to_pc_torch` (`ldm/dataset.py:228-276`)
    in its three range encodings, and == the depth-masked rows of the writer loop (`ldm/inference.py:175-179`)."""
    from oracle import geometry as G
    d = golden("range_to_points.pt")
    for name, mode in (("linear", G.MODE_LINEAR), ("log", G.MODE_LOG), ("inverse", G.MODE_INVERSE)):
        y = G.to_points(d["image"], d["incl"], d["height"], mode, d["mean"], d["std"], d["fill"])
        assert y.shape == d[name].shape and torch.equal(y, d[name])
    rows = G.depth_masked(G.to_points(d["image"], d["incl"], d["height"], 0, d["mean"], d["std"], d["fill"])[0])
    assert rows.dtype.name == "float32" and torch.equal(torch.from_numpy(rows), d["masked_rows"])
    # edge cases: single channel -> (x, y, z) only; empty batch
    assert G.to_points(d["image"][:, :1], d["incl"], d["height"]).shape == (2, 96 * 64, 3)
    assert G.to_points(d["image"][:0], d["incl"], d["height"]).shape == (0, 96 * 64, 4)


def test_geometry_oracle_matches_reference_to_voxel(golden):
    """oracle to_voxel == the reference's `to_voxel` / `_splat_points_to_volumes` (`ldm/dataset.py:13-132,278-294`)."""
    from oracle import geometry as G
    d = golden("range_to_points.pt")
    v = G.to_voxel(d["image"], d["incl"], d["height"], G.MODE_LINEAR, d["mean"], d["std"], d["fill"],
                   tuple(d["voxel_grid"]), tuple(d["voxel_range"]))
    assert v.shape == d["voxel"].shape and torch.allclose(v, d["voxel"], rtol=0, atol=1e-6)
    assert G.bev_image(v[0]).shape == (128, 96) and G.bev_image(v[0]).dtype.name == "uint8"


def test_geometry_oracle_matches_reference_projection(golden):
    """oracle points_to_range_image == the reference's projection + process_miss_value + normalize
    (`ldm/dataset.py:159-226`, `ldm/kitti360_range_image.py:51-61`) in the three range encodings."""
    from oracle import geometry as G
    d = golden("range_to_points.pt")
    for name, mode in (("linear", G.MODE_LINEAR), ("log", G.MODE_LOG), ("inverse", G.MODE_INVERSE)):
        img, m, c = G.points_to_range_image(d["proj_points"].numpy(), d["incl"], d["height"], 128, mode)
        assert torch.equal(img, d["proj_" + name])
        assert torch.equal(m, d["proj_mask_" + name]) and torch.equal(c, d["proj_car_" + name])
