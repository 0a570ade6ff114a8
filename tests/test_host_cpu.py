"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/rldm.h
declares (no compute without a GPU), the diffusers-compatible surface, the reference's own surgery
and import sites, dry-run plan compilation, scheduler tables, and the world_size-2 sharding contract."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from rangeldm_b200.build import build_library
    return build_library()


def test_library_exports_every_declared_symbol(built):
    from rangeldm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "rldm.h")).read()
    declared = set(re.findall(r"\b(rldm_[a-z_0-9]+)\s*\(", header))
    declared.discard("rldm_op")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.rldm_version() == 100


def test_compute_without_cuda_fails_loudly(built):
    import rangeldm_b200 as R
    if torch.cuda.is_available():
        pytest.skip("has CUDA")
    u = R.UNet2DModel(sample_size=[32, 8], in_channels=5, out_channels=4, layers_per_block=1,
                      block_out_channels=[64, 128], down_block_types=["DownBlock2D", "AttnDownBlock2D"],
                      up_block_types=["AttnUpBlock2D", "UpBlock2D"])
    with pytest.raises(RuntimeError, match="CUDA"):
        u(torch.zeros(1, 5, 32, 8), 10)
    s = R.DDIMScheduler(clip_sample=False)
    s.set_timesteps(10)
    with pytest.raises(RuntimeError, match="CUDA"):
        s.step(torch.zeros(1, 4, 8, 8), 0, torch.zeros(1, 4, 8, 8))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rangeldm_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_config_roundtrip_and_state_dict_keys(tmp_path):
    import rangeldm_b200 as R
    from oracle import nets
    u = R.UNet2DModel(**nets.UNET_C3)
    assert u.config.sample_size == [256, 16] and u.config.in_channels == 5 and u.config["out_channels"] == 4
    assert sum(p.numel() for p in u.parameters()) == 30135684            # SURVEY.md anchor
    assert set(u.state_dict()) == set(nets.OracleUNet2DModel(**nets.UNET_C3).state_dict())
    u2 = R.UNet2DModel(**nets.UNET_C2)
    assert sum(p.numel() for p in u2.parameters()) == 113672066
    u.save_pretrained(tmp_path / "unet")
    cfg = R.UNet2DModel.load_config(tmp_path / "unet")
    u3 = R.UNet2DModel.from_config(cfg)
    import safetensors.torch
    safetensors.torch.load_model(u3, str(tmp_path / "unet" / "diffusion_pytorch_model.safetensors"))
    assert torch.equal(u3.conv_in.weight, u.conv_in.weight)
    s = R.DDPMScheduler(clip_sample=False)
    s.save_config(tmp_path / "scheduler")
    s2 = R.DDPMScheduler.from_config(R.DDPMScheduler.load_config(tmp_path / "scheduler"))
    d = R.DDIMScheduler.from_config(s2.config)                           # `ldm/pipelines.py:139`
    assert d.config.clip_sample is False and d.config.timestep_spacing == "leading"
    with pytest.raises(NotImplementedError):
        R.UNet2DModel(attention_head_dim=16)


@pytest.mark.skipif(not os.path.isdir("/root/reference/ldm"), reason="reference tree not present")
def test_reference_surgery_and_pipelines_import_against_the_shim():
    """The reference's own ldm/utils.py and ldm/pipelines.py import unchanged and operate on our classes."""
    code = r'''
import sys, torch
sys.path.insert(0, %r)
import rangeldm_b200 as R
from rangeldm_b200 import diffusers_compat
diffusers_compat.install()
sys.path.insert(0, "/root/reference/ldm")
import utils as refutils, pipelines as refpipes
u = R.UNet2DModel(sample_size=[32, 8], in_channels=5, out_channels=4, layers_per_block=1, block_out_channels=[64, 128],
                  down_block_types=["DownBlock2D", "AttnDownBlock2D"], up_block_types=["AttnUpBlock2D", "UpBlock2D"])
refutils.replace_down(u); refutils.replace_conv(u)
assert type(u.conv_in) is refutils.Conv2d and u.conv_in.circular
assert type(u.down_blocks[0].downsamplers[0]) is refutils.Downsample2D
v = R.AutoencoderKL(in_channels=2, out_channels=2, down_block_types=["DownEncoderBlock2D"] * 2,
                    up_block_types=["UpDecoderBlock2D"] * 2, block_out_channels=[64, 128], layers_per_block=1)
v.quant_conv = torch.nn.Identity(); v.post_quant_conv = torch.nn.Identity()
refutils.replace_down(v); refutils.replace_conv(v); refutils.replace_attn(v)
assert type(v.decoder.mid_block.attentions[0]) is refutils.attn_identity
assert v.encoder.down_blocks[0].downsamplers[0].padding == 0
pipe = refpipes.LDMPipelineRange(v, u, R.DDPMScheduler(clip_sample=False), pos_encoding=True)
assert pipe.unet is u and isinstance(refpipes.DDIMPipelineRange(u, R.DDPMScheduler(clip_sample=False)).scheduler, R.DDIMScheduler)
import os
os.environ["RLDM_DRYRUN"] = "1"
plan = u.plan(2, 32, 8, 1)
dp = v.decoder_plan(2, 32, 8)
print("OK", len(plan.prog.ops), len(dp.prog.ops))
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_dry_run_plans_and_fused_program_launch_counts(monkeypatch):
    monkeypatch.setenv("RLDM_DRYRUN", "1")
    import rangeldm_b200 as R
    from rangeldm_b200 import _lib
    from oracle import nets
    u = R.UNet2DModel(**nets.UNET_C3)
    R.replace_down(u); R.replace_conv(u)
    plan = u.plan(8, 256, 16, 1)
    kinds = [op.kind for op in plan.prog.ops]
    # 95 convolutions / projections, 13 of them 1x1 conv_shortcuts that ride inside their block's conv2 launch
    assert kinds.count(_lib.OP_CONV_TC) == 82 and kinds.count(_lib.OP_ATTENTION) == 16
    assert kinds.count(_lib.OP_CONV_IN) == 1 and kinds.count(_lib.OP_NORM_CONV_OUT) == 1 and kinds[0] == _lib.OP_MEMSET
    assert kinds.count(_lib.OP_GN_STATS) == 0          # every GroupNorm reads moments fused into a producing epilogue
    assert [op.kind for op in plan.prog.exec_ops] == kinds and plan.prog.n_launch == len(kinds) + 2   # temb = 3 launches
    # 17 convolutions of level 3 (images of 64 pixels: whole images inside one K-split cluster) write the next GroupNorm's
    # operand themselves, 7 of them (the conv1 of the ResnetBlock2Ds) without an fp32 output: 151 graph nodes
    convs = [op for op in plan.prog.ops if op.kind == _lib.OP_CONV_TC]
    assert kinds.count(_lib.OP_PREP) == 49 and len(kinds) == 151
    assert sum(1 for op in convs if op.p[19]) == 17 and sum(1 for op in convs if not op.p[5]) == 7
    assert all(op.p[19] and op.p[20] and op.p[21] and op.i[20] == 32 for op in convs if not op.p[5])
    assert not any(op.p[11] or op.p[17] for op in convs)
    from rangeldm_b200 import engine
    # RLDM_EMIT_PREP=0: one rldm_prep launch per GroupNorm / resampler operand, 168 graph nodes
    monkeypatch.setattr(engine, "EMIT_PREP", False)
    u.invalidate_plans()
    plan0 = u.plan(8, 256, 16, 1)
    kinds0 = [op.kind for op in plan0.prog.ops]
    assert kinds0.count(_lib.OP_PREP) == 66 and len(kinds0) == 168
    assert not any(op.p[19] for op in plan0.prog.ops if op.kind == _lib.OP_CONV_TC)
    # opt-in experiment RLDM_FUSE_PREP=1: the 50 convolutions of levels 1-3 that run on the small-layer kernel produce
    # their own operand; only the full-resolution level and the three 192-tile qkv projections keep a rldm_prep launch
    monkeypatch.setattr(engine, "FUSE_PREP", True)
    u.invalidate_plans()
    plan1 = u.plan(8, 256, 16, 1)
    kinds1 = [op.kind for op in plan1.prog.ops]
    assert kinds1.count(_lib.OP_PREP) == 16 and len(kinds1) == 118
    assert sum(1 for op in plan1.prog.ops if op.kind == _lib.OP_CONV_TC and (op.p[11] or op.p[17])) == 50
    monkeypatch.setattr(engine, "FUSE_PREP", False)
    # experiment RLDM_EMIT_MAXCLM=8: the clusters may also span the M tiles of larger images -- 40 convolutions of levels
    # 1-3 emit (17 of them, the conv1 of the ResnetBlock2Ds, no longer write an fp32 output at all)
    monkeypatch.setenv("RLDM_EMIT_MAXCLM", "8")
    _lib.lib().rldm_reload_env()
    monkeypatch.setattr(engine, "EMIT_PREP", True)
    u.invalidate_plans()
    plan3 = u.plan(8, 256, 16, 1)
    kinds3 = [op.kind for op in plan3.prog.ops]
    convs3 = [op for op in plan3.prog.ops if op.kind == _lib.OP_CONV_TC]
    assert len(kinds3) == 128 and kinds3.count(_lib.OP_PREP) == 26
    assert sum(1 for op in convs3 if op.p[19]) == 40 and sum(1 for op in convs3 if not op.p[5]) == 17
    monkeypatch.delenv("RLDM_EMIT_MAXCLM")
    _lib.lib().rldm_reload_env()
    monkeypatch.setattr(engine, "EMIT_PREP", False)
    # opt-in experiment RLDM_FUSE_LEVELS=1: 135 small ops of levels 1-3 as 5 fused persistent launches between the five
    # N = 1024 attention kernels
    monkeypatch.setattr(engine, "FUSE_LEVELS", True)
    u.invalidate_plans()
    plan2 = u.plan(8, 256, 16, 1)
    assert len(plan2.prog.ops) == 168
    ex = [op.kind for op in plan2.prog.exec_ops]
    assert ex.count(_lib.OP_FUSED) == 5 and len(ex) <= 40 and sum(op.n for op in plan2.prog.exec_ops if op.kind == _lib.OP_FUSED) == 135
    assert plan2.prog.n_launch == len(ex) + 2
    monkeypatch.setattr(engine, "FUSE_LEVELS", False)
    u.invalidate_plans()
    plan = u.plan(8, 256, 16, 1)
    sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
    sch.set_timesteps(20)
    v = R.AutoencoderKL(in_channels=2, out_channels=2, down_block_types=["DownEncoderBlock2D"] * 3,
                        up_block_types=["UpDecoderBlock2D"] * 3, block_out_channels=[64, 128, 256], layers_per_block=2)
    v.quant_conv = torch.nn.Identity(); v.post_quant_conv = torch.nn.Identity()
    R.replace_down(v); R.replace_conv(v); R.replace_attn(v)
    fs = R.FusedSampler(u, sch, v, 8, 1, use_graph=False, streams=2)
    assert len(fs.parts) == 2 and fs.slices == [slice(0, 4), slice(4, 8)]       # two sub-batch trajectories
    for part in fs.parts:
        n_sched = sum(1 for op in part.prog.ops if op.kind == _lib.OP_SCHED_STEP)
        assert n_sched == 20 and part.image.shape == (4, 2, 1024, 64)
        # the time embedding of the whole timestep table is ONE op at the head of the (timed, graph-captured) program
        assert [k for k, op in enumerate(part.prog.ops) if op.kind == _lib.OP_TEMB] == [0]
        assert part.prog.ops[0].i[0] == 20
        convs_with_temb = [op for op in part.prog.ops if op.kind == _lib.OP_CONV_TC and op.p[3]]
        assert len(convs_with_temb) == 20 * 22 and all(op.i[0] == 0 for op in convs_with_temb)
    assert fs.parts[0].plan is not fs.parts[1].plan                               # own activation buffers ...
    assert len(u._packed) > 0 and len(fs.parts[0].plan.prog.ops) <= len(plan.prog.ops)   # ... shared weights (batch 4:
    # more layers fit the small-layer kernel and produce their own operand)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fs.parts[0].prog.run()
    v2 = R.AutoencoderKL(in_channels=2, out_channels=2, down_block_types=["DownEncoderBlock2D"],
                         up_block_types=["UpDecoderBlock2D"], block_out_channels=[64])
    with pytest.raises(NotImplementedError):
        v2.decoder_plan(1, 16, 8)                    # learned quant convs are not implemented


def test_sampler_cache_follows_weight_reloads_and_surgery(monkeypatch):
    """a cached FusedSampler must not outlive `load_state_dict` / `.to()` / surgery of its models."""
    monkeypatch.setenv("RLDM_DRYRUN", "1")
    import rangeldm_b200 as R
    from oracle.make_golden import TINY_UNET
    u = R.UNet2DModel(**TINY_UNET)
    R.replace_down(u); R.replace_conv(u)
    sch = R.DDIMScheduler(clip_sample=False)
    pipe = R.DDIMPipelineRange(u, sch, pos_encoding=True)
    pipe.scheduler.set_timesteps(3)
    monkeypatch.setattr(R.pipelines.FusedSampler, "_capture", lambda self: None)
    a = pipe._sampler(2, 1, None)
    assert pipe._sampler(2, 1, None) is a
    v0 = u._plan_version
    u.load_state_dict(u.state_dict())
    assert u._plan_version > v0
    b = pipe._sampler(2, 1, None)
    assert b is not a and pipe._sampler(2, 1, None) is b
    R.replace_conv(u)
    assert pipe._sampler(2, 1, None) is not b
    # replace_conv on a bare convolution is a no-op, like the reference's (`ldm/utils.py:125-146`)
    c = R.UNet2DModel(**TINY_UNET).conv_in
    R.replace_conv(c)
    assert not c.circular
    # long ancestral trajectories do not unroll (O(1) extra memory like the reference loop)
    ddpm = R.DDPMPipelineRange(u, R.DDPMScheduler(clip_sample=False))
    ddpm.scheduler.set_timesteps(1000)
    assert not ddpm._fusable(1000, 16 * 2 * 1024 * 64)
    ddpm.scheduler.set_timesteps(50)
    assert ddpm._fusable(50, 2 * 2 * 1024 * 64)


def test_scheduler_tables_bit_exact_with_oracle():
    import rangeldm_b200 as R
    from oracle import schedulers as O
    for n in (5, 20, 50):
        a, b = R.DDIMScheduler(clip_sample=False), O.OracleDDIMScheduler()
        a.set_timesteps(n); b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps)
        for sp in ("linspace", "leading", "trailing"):
            a = R.DPMSolverMultistepScheduler(timestep_spacing=sp)
            b = O.OracleDPMSolverMultistepScheduler(timestep_spacing=sp)
            a.set_timesteps(n); b.set_timesteps(n)
            assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.sigmas, b.sigmas)
    d = R.DDIMScheduler(clip_sample=False); d.set_timesteps(50)
    assert d.timesteps.tolist() == list(range(980, -1, -20))


def test_dpm_coefficient_table_reproduces_oracle_update_on_cpu():
    """The 7-coefficient affine form the fused step kernel evaluates == the oracle's DPM-Solver++ update, for the
    20-step benchmark table and for n < 15 (where only the FINAL step is first order)."""
    import rangeldm_b200 as R
    from oracle import schedulers as O
    for n in (20, 5, 3, 14):
        a = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
        b = O.OracleDPMSolverMultistepScheduler(timestep_spacing="leading")
        a.set_timesteps(n); b.set_timesteps(n)
        g = torch.Generator().manual_seed(0)
        x = torch.randn(64, generator=g)
        xa, xb, prev = x.clone(), x.clone(), torch.zeros(64)
        for i, t in enumerate(b.timesteps):
            eps = torch.randn(64, generator=g)
            k = a._coef_host[i]
            x0 = k[0] * xa + k[1] * eps
            xa = k[2] * xa + k[3] * x0 + k[4] * prev + k[5] * eps
            prev = x0
            xb = b.step(eps, t, xb)
            assert torch.allclose(xa, xb, rtol=2e-5, atol=2e-5), (n, i)
        # second-order everywhere except the first and the last step
        assert [bool(k[4] != 0) for k in a._coef_host] == [False] + [True] * (n - 2) + [False]


def test_scheduler_coefficient_tables_reproduce_reference_sampler_goldens(golden):
    """The PRODUCT's host coefficient tables (what `rldm_sched_step` evaluates) against trajectories of the reference's
    own samplers (`vae/sgm/modules/diffusionmodules/sampling.py`: DPMPP2MSampler n=5 and n=20, EulerEDMSampler == DDIM,
    EulerAncestralSampler == DDPM; fixtures made by oracle/make_golden.py)."""
    import rangeldm_b200 as R
    from oracle.make_golden import toy_eps_matrix
    Wm = toy_eps_matrix()
    sam = golden("samplers.pt")
    cases = [(R.DPMSolverMultistepScheduler(timestep_spacing="leading"), sam["dpm5"], None),
             (R.DPMSolverMultistepScheduler(timestep_spacing="leading"), golden("dpmpp2m.pt"), None),
             (R.DDIMScheduler(clip_sample=False), sam["ddim"], None),
             (R.DDPMScheduler(clip_sample=False), sam["ddpm"], sam["ddpm"]["noise"])]
    for sch, g, noise in cases:
        n = g["traj"].shape[0]
        sch.set_timesteps(n)
        assert torch.equal(sch.timesteps, g["timesteps"])
        x, prev = g["x"].clone(), torch.zeros_like(g["x"])
        for i in range(n):
            eps = torch.tanh(x @ Wm)
            k = sch._coef_host[i]
            x0 = k[0] * x + k[1] * eps
            x = k[2] * x + k[3] * x0 + k[4] * prev + k[5] * eps + (k[6] * noise[i] if noise is not None else 0)
            prev = x0
            e = ((x - g["traj"][i]).abs().max() / g["traj"][i].abs().max()).item()
            assert e < 2e-5, (type(sch).__name__, n, i, e)


def test_sparse_encoder2_golden(golden):
    import rangeldm_b200 as R
    g = golden("sparse_encoder2.pt")
    assert torch.equal(R.SparseRangeImageEncoder2()(g["x"]), g["y"])


def test_randn_tensor_semantics():
    from rangeldm_b200 import randn_tensor
    a = randn_tensor((2, 3), generator=torch.Generator().manual_seed(1))
    b = torch.randn((2, 3), generator=torch.Generator().manual_seed(1))
    assert torch.equal(a, b)
    gens = [torch.Generator().manual_seed(5), torch.Generator().manual_seed(6)]
    c = randn_tensor((2, 3), generator=gens)
    assert torch.equal(c[1:], torch.randn((1, 3), generator=torch.Generator().manual_seed(6)))


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from rangeldm_b200.sharding import shard_indices, gather_images
    samples, batch = 10, 2
    mine = shard_indices(samples, batch, rank, world)
    # a stand-in "image" that only depends on the global index (what seed-per-index sampling guarantees)
    imgs = torch.stack([torch.full((2, 4, 4), float(i)) for i in mine]) if mine else torch.zeros(0, 2, 4, 4)
    out = gather_images(imgs, mine, samples)
    q.put((rank, mine, None if out is None else out[:, 0, 0, 0].tolist()))
    dist.destroy_process_group()


def test_world_size_2_sharding_contract_gloo():
    """`ldm/inference.py:159,174-183`: rank r of P generates global indices (r + P*i)*B + j; the optional
    gather returns every index exactly once, in order, on rank 0."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == [0, 1, 4, 5, 8, 9] and res[1][1] == [2, 3, 6, 7]
    assert res[0][2] == [float(i) for i in range(10)] and res[1][2] is None


def test_documented_switches_exist_in_the_sources():
    """Every RLDM_* environment switch that DESIGN.md documents is read somewhere in the package (no doc rot)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "DESIGN.md")).read()
    documented = set(re.findall(r"`(RLDM_[A-Z0-9_]+)", text))
    assert documented, "DESIGN.md lists no switches?"
    sources = ""
    pkg = os.path.join(root, "rangeldm_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py")):
                sources += open(os.path.join(d, f)).read()
    missing = sorted(s for s in documented if s not in sources)
    assert not missing, f"documented but not read anywhere: {missing}"


def test_program_alloc_respects_the_op_that_writes_the_buffer(monkeypatch):
    """An operand EMITTED by an already recorded op (rldm_conv_tc_emit) must not land in a buffer that op still reads:
    `alloc(before_op=k)` only reuses buffers that were free before op k was appended."""
    monkeypatch.setenv("RLDM_DRYRUN", "1")
    from rangeldm_b200 import engine, _lib
    pg = engine.Program(torch.device("cpu"))
    a = pg.alloc((4, 8), torch.float16)                 # e.g. the operand conv k reads
    early = pg.alloc((4, 8), torch.float16)
    pg.free(early)                                      # free before op 0 exists
    pg.add(_lib.OP_MEMSET, p=(a,), n=0)                 # op 0 reads / writes `a`
    pg.free(a)                                          # ... and `a` is released right after it
    b = pg.alloc((4, 8), torch.float16, before_op=0)    # written BY op 0: may take `early`, never `a`
    assert b.data_ptr() == early.data_ptr() and b.data_ptr() != a.data_ptr()
    c = pg.alloc((4, 8), torch.float16, before_op=0)    # nothing else qualifies: a fresh buffer
    assert c.data_ptr() not in (a.data_ptr(), early.data_ptr())
    d = pg.alloc((4, 8), torch.float16)                 # an ordinary allocation may reuse `a`
    assert d.data_ptr() == a.data_ptr()


def test_pipeline_set_steps_reuses_tables_only_for_identical_arguments():
    """`_RangePipeline._set_steps`: the second call with the same step count keeps the installed tables (and resets the
    multistep state); another step count, another eta, or a direct `scheduler.set_timesteps` re-installs them."""
    import rangeldm_b200 as R
    from rangeldm_b200.pipelines import _RangePipeline
    p = _RangePipeline.__new__(_RangePipeline)
    sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")
    p.scheduler = sch
    p._set_steps(20)
    t20 = sch.timesteps
    sch._state["x0_prev"] = object()
    p._set_steps(20)
    assert sch.timesteps is t20 and sch._state == {}
    p._set_steps(10)
    assert sch.timesteps is not t20 and len(sch.timesteps) == 10
    sch.set_timesteps(20)                       # direct call by the user: the pipeline must not trust its key afterwards
    t20b = sch.timesteps
    p._set_steps(10)
    assert len(sch.timesteps) == 10 and sch.timesteps is not t20b
    d = R.DDIMScheduler(clip_sample=False)
    p.scheduler = d
    p._set_steps(10, eta=0.0)
    c0 = d._coef_host
    p._set_steps(10, eta=0.5)
    assert d._coef_host is not c0 and float(d._coef_host[:, 6].abs().max()) > 0


def test_folded_upsample_phase_weights_reproduce_the_upsampled_convolution(monkeypatch):
    """`Builder.pack_conv_up2` (host side of rldm_conv_tc_up2): the four 2x2 phase convolutions over the low-resolution
    input, with the pads librldm gives them (W: 1 - a, H: 1 - b; circular along W, zeros along H), reproduce
    conv3x3(nearest-2x(x)) -- checked here with PyTorch on the CPU, in fp64 on the packed fp16 hi + lo planes."""
    monkeypatch.setenv("RLDM_DRYRUN", "1")
    import torch.nn.functional as F
    from rangeldm_b200 import engine, models
    from oracle.nets import circ_conv2d
    g = torch.Generator().manual_seed(5)
    Cin, Cout, W, H = 8, 16, 6, 4
    conv = models.LoRACompatibleConv(Cin, Cout, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g))
        conv.bias.copy_(torch.randn(Cout, generator=g))
    pg = engine.Program(torch.device("cpu"))
    bd = engine.Builder(pg, 1, cache={})
    wt, bias = bd.pack_conv_up2(conv, 3)                     # [4 phases][2 planes * 4 taps][Cout][Cin] fp16
    assert wt.shape == (4, 8, Cout, Cin) and wt.dtype == torch.float16
    x = torch.randn(2, Cin, W, H, generator=g, dtype=torch.float64)
    with torch.no_grad():
        ref = circ_conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), conv.weight.double(), conv.bias.double(), 1, 1)
    out = torch.zeros_like(ref)
    for a in range(2):
        for b in range(2):
            w4 = (wt[2 * a + b, :4].double() + wt[2 * a + b, 4:].double())          # hi + lo planes, taps 2*ti + tj
            k = w4.reshape(2, 2, Cout, Cin).permute(2, 3, 0, 1)                     # (Cout, Cin, ti, tj)
            pw, ph = 1 - a, 1 - b                                                   # low-side pads of this phase
            xp = torch.cat([x[:, :, -1:], x, x[:, :, :1]], dim=2)                   # circular halo along W
            xp = xp[:, :, 1 - pw: 1 - pw + W + 1]                                   # columns w - pw .. w - pw + 1
            xp = F.pad(xp, (ph, 1 - ph))                                            # zeros along H
            out[:, :, a::2, b::2] = F.conv2d(xp, k) + conv.bias.detach().double()[None, :, None, None]
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-6             # fp16 hi + lo weights: ~22 bits
