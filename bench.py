#!/usr/bin/env python
"""bench.py -- range-images/sec (64x1024, 20-step DPM-Solver++) on N B200s, next to the CPU oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2], "C3"): RangeLDM KITTI-360 -- 5->4 channel latent UNet
[128,128,256,256] on 256x16 latents (30.14 M params), 20-step DPM-Solver++(2M), AutoencoderKL 4x
decoder to a (2,1024,64) range image; per-GPU batch 8 (global batch 64 on 8 GPUs, weak scaling);
synthetic N(0,1) noise, random-init weights.  One "step" = one batch: the whole trajectory
(20 x (time-embedding + UNet + scheduler step) + rescale + VAE decode), replayed as one CUDA graph.

  value : images/s, device-timed (CUDA events, max over ranks), noise already resident in HBM
  e2e   : the same through the public call `LDMPipelineRange.__call__` (CPU randn like the reference,
          H2D of the noise, sampling, D2H of the finished images into host memory)
  roofline : tcgen05 conv kernel: algorithmic FLOPs / CUDA-event time of its launches (per-op pass)
  cpu_baseline : the fp32 PyTorch oracle (restated diffusers modules, reference loop) on the host cores
  gpu_library_baseline : the SAME oracle modules (stock PyTorch ops: cuDNN convolutions, F.group_norm, SDPA -- what
          the reference's diffusers/sgm modules call) on this B200 at the same batch: strict fp32, PyTorch defaults
          (TF32 convolutions), bf16 autocast; eager and as one CUDA graph (the bar for the kernels)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_C3 = dict(sample_size=[256, 16], in_channels=5, out_channels=4, layers_per_block=2,
               block_out_channels=[128, 128, 256, 256],
               down_block_types=["DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"],
               up_block_types=["AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"])
VAE_KITTI = dict(in_channels=2, out_channels=2, down_block_types=["DownEncoderBlock2D"] * 3,
                 up_block_types=["UpDecoderBlock2D"] * 3, block_out_channels=[64, 128, 256], layers_per_block=2,
                 latent_channels=4, scaling_factor=0.18215)
STEPS = 20
PER_GPU_BATCH = 8
GFLOP_PER_IMAGE = 20 * 34.07 + 157.46
METRIC = "range-images/sec (64x1024, 20-step DPM-Solver)"
DTYPE = ("fp16 operands, fp32 accumulate on tcgen05 kind::f16; UNet inside the trajectory graph: plain fp16 (1 MMA per "
         "algorithmic MAC); VAE decoder: split-fp16 x3 (activations and weights as hi+lo fp16 pairs, Ah*Wh + Al*Wh + "
         "Ah*Wl, ~22-bit operands, 3 MMAs per MAC); fp32 residual stream / softmax / scheduler, GroupNorm moments in "
         "fp64 (rangeldm_b200/engine.py PRECISION*)")
WORKLOAD = ("C3 RangeLDM KITTI-360: latent 4x256x16 UNet[128,128,256,256] x 20-step DPM-Solver++(2M, leading) "
            "+ AutoencoderKL 4x decode -> 2x1024x64")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def conv_traffic():
    """DRAM bytes (read + write) per conv_tc launch from the committed ncu capture of the same workload
    (profiles/conv_tc_traffic_r2.json, written by scripts/conv_traffic.py; the round-1 file as a fallback); None when
    neither exists."""
    for name in ("conv_tc_traffic_r2.json", "conv_tc_traffic_r1.json"):
        try:
            return round(json.load(open(os.path.join(ROOT, "profiles", name)))["traffic_bytes_per_launch"])
        except Exception:
            continue
    return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "power_w_max": max(power), "samples": len(sm)}
        return out


def build_pipeline(dev):
    import rangeldm_b200 as R
    torch.manual_seed(0)
    unet = R.UNet2DModel(**UNET_C3)
    vae = R.AutoencoderKL(**VAE_KITTI)
    vae.quant_conv = torch.nn.Identity()
    vae.post_quant_conv = torch.nn.Identity()
    for m in (unet, vae):                       # the surgery every shipped config applies (all_circonv)
        R.replace_down(m)
        R.replace_conv(m)
    R.replace_attn(vae)
    sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading")   # from_config(training DDPM config) -> leading
    return R.LDMPipelineRange(vae, unet, sch, pos_encoding=True).to(dev)


def conv_flops(op):
    i = op.i
    B, W, H, Cin, Cout, ks, stride = i[1], i[2], i[3], i[4], i[5], i[6], i[7]
    sc_cin = i[11]                       # 1x1 conv_shortcut folded into this launch (0: none)
    return 2.0 * B * (W // stride) * (H // stride) * Cout * (Cin * ks * ks + sc_cin)


def up2_flops(op):
    """rldm_conv_tc_up2 (nearest-2x upsampling folded into the convolution): the multiply-adds it EXECUTES -- four phases of
    a 2x2 convolution over the low-resolution grid, 16/36 of the 3x3 convolution on the upsampled tensor it replaces."""
    B, W, H, Cin, Cout = op.i[0], op.i[1], op.i[2], op.i[3], op.i[4]
    return 2.0 * B * W * H * 16 * Cin * Cout


def up2_bytes(op):
    B, W, H, Cin, Cout, terms = op.i[0], op.i[1], op.i[2], op.i[3], op.i[4], op.i[6]
    xp, wp = (2 if terms == 3 else 1), (2 if terms >= 2 else 1)
    return float(xp * B * (W + 2) * H * Cin * 2 + wp * 16 * Cin * Cout * 2 + B * 4 * W * H * Cout * 4)


def conv_bytes(op):
    """Algorithmic HBM bytes of one conv_tc launch: every operand plane, weight plane, residual and output element moves
    once (fp16 operand planes as the layer's precision says, fp32 output / residual)."""
    i = op.i
    B, W, H, Cin, Cout, ks, stride, sc_cin, terms = i[1], i[2], i[3], i[4], i[5], i[6], i[7], i[11], i[12]
    xp, wp = (2 if terms == 3 else 1), (2 if terms >= 2 else 1)
    out_elems = B * (W // stride) * (H // stride) * Cout
    b = xp * B * (W + 2) * H * (Cin + sc_cin) * 2 + wp * (ks * ks * Cin + sc_cin) * Cout * 2 + out_elems * 4
    if op.p[4]:
        b += out_elems * 4
    return float(b)


def timed_profile(prog, reps=5):
    """Per-op durations (us) of a librldm program: `rldm_run_timed` puts a one-thread %globaltimer stamp kernel after
    every op; the stamped program is captured in a CUDA graph and replayed, so the numbers are device-side, cache-warm
    and free of host launch overhead (each includes one ~2 us stamp-kernel launch)."""
    from rangeldm_b200 import _lib
    if prog.arr is None:
        prog.finalize()
    n = len(prog.exec_ops)
    stamps = torch.zeros(n + 1, dtype=torch.int64, device=prog.device)
    lib = _lib.lib()
    run = lambda: _lib.check(lib.rldm_run_timed(prog.arr, n, stamps.data_ptr(), _lib.stream_ptr()))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    acc = torch.zeros(n, dtype=torch.float64)
    for _ in range(reps):
        g.replay()
        torch.cuda.synchronize()
        t = stamps.cpu().double()
        acc += (t[1:] - t[:-1]) / 1e3
    return (acc / reps).tolist()


OP_NAMES = {1: "gn_stats", 2: "prep", 3: "conv_tc", 4: "conv_in", 5: "conv_out", 6: "attention", 7: "temb",
            8: "sched_step", 9: "memset", 10: "conv_ref", 11: "scale", 12: "norm_conv_out", 13: "fused_levels", 14: "conv_tc"}


def per_op_profile(sampler):
    """One UNet forward + the decoder, every op timed on the device inside a CUDA graph (`timed_profile`), weighted as
    one step runs them (STEPS UNet forwards, one decode).  Returns per-kind milliseconds and launch counts per step, and
    the conv kernels' algorithmic FLOPs, issued FLOPs and algorithmic HBM bytes per step."""
    from rangeldm_b200 import _lib
    # (profiles the first sub-batch program: with RLDM_STREAMS=2 that is half of the per-GPU batch)
    progs = [(sampler.plan.prog, STEPS)] + ([(sampler.dec.prog, 1)] if sampler.dec is not None else [])
    ms, cnt, flops, fused_flops, hw_flops, nbytes = {}, {}, 0.0, 0.0, 0.0, 0.0
    set_bytes, set_n = 0.0, 0          # the launches of ONE UNet forward + one decode: the set the ncu traffic figure averages
    for prog, weight in progs:
        for op, (i, j), us in zip(prog.exec_ops, prog.exec_src, timed_profile(prog)):
            k = OP_NAMES.get(op.kind, str(op.kind))
            ms[k] = ms.get(k, 0.0) + weight * us / 1e3
            cnt[k] = cnt.get(k, 0) + weight
            if op.kind == _lib.OP_CONV_TC:
                flops += weight * conv_flops(op)
                hw_flops += weight * conv_flops(op) * op.i[12]       # MMAs actually issued: 1, 2 or 3 per algorithmic MAC
                nbytes += weight * conv_bytes(op)
                set_bytes += conv_bytes(op); set_n += 1
            if op.kind == _lib.OP_CONV_UP2:       # four role-swapped conv launches: counted with the conv kernels
                flops += weight * up2_flops(op)
                hw_flops += weight * up2_flops(op) * op.i[6]
                nbytes += weight * up2_bytes(op)
                set_bytes += up2_bytes(op); set_n += 4
                cnt[k] += 3 * weight            # four launches behind one op
            if op.kind == _lib.OP_FUSED:          # convolutions inside a fused run of small layers
                fused_flops += weight * sum(conv_flops(o) for o in prog.ops[i:j] if o.kind == _lib.OP_CONV_TC)
    return ms, cnt, flops, fused_flops, hw_flops, nbytes, set_bytes / max(set_n, 1)


def run_native(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE line (the JSON): library chatter written to fd 1 during the run (e.g. NCCL's
    # "NCCL version ..." banner at communicator creation) is sent to stderr until the result is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = PER_GPU_BATCH
    pipe = build_pipeline(dev)
    pipe.set_progress_bar_config(disable=True)
    pipe.scheduler.set_timesteps(STEPS)
    sampler = pipe._sampler(B, 1, pipe.vae)                 # compiles + captures the trajectory graph
    from rangeldm_b200.pipelines import make_pos_encoding
    pos = make_pos_encoding(B, 256, 16, dev)
    gen = torch.Generator(device=dev).manual_seed(rank)
    noise = torch.randn((B, 4, 256, 16), generator=gen, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    gathered = torch.empty((world, B, 2, 1024, 64), device=dev) if world > 1 else None

    def one_step():
        sampler.load(noise, pos)            # device -> device copies of the (already resident) inputs
        sampler.replay()
        if world > 1:       # the only collective: collect the finished range images (north star)
            dist.all_gather_into_tensor(gathered, sampler.result())

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local) if rank == 0 else None
    evs = []
    for _ in range(args.steps):
        flush.zero_()                                       # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop() if clocks else None
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * args.steps * B / (total_ms / 1e3)

    # ---- end to end through the public pipeline call: CPU randn -> H2D -> sample -> D2H ----
    host_img = torch.empty((B, 2, 1024, 64), pin_memory=True)
    g2 = torch.Generator().manual_seed(1000 + rank)
    for _ in range(2):
        host_img.copy_(pipe(batch_size=B, generator=g2, num_inference_steps=STEPS, output_type="torch"))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        img = pipe(batch_size=B, generator=g2, num_inference_steps=STEPS, output_type="torch")
        host_img.copy_(img, non_blocking=False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * B / float(t.item())

    out = None
    if rank == 0:
        burst, sustained, hbm, src = peaks()
        ms, cnt, flops, fused_flops, hw_flops, conv_nbytes, set_bytes = per_op_profile(sampler)
        conv_ms = ms.get("conv_tc", 0.0)
        all_ms = sum(ms.values())
        achieved = flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        launches_per_step = sampler.gpu_launches
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic N(0,1) noise, random-init weights",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world, "sampler_steps": STEPS,
                       "parallelism": f"dp{world} (batch-axis shards, no hot-path collective; "
                                      f"all_gather of finished images)" if world > 1 else "single GPU",
                       "l2": "256 MiB L2 flush between timed iterations; working set ~1 GB > 126 MB L2",
                       "gflop_per_image": GFLOP_PER_IMAGE},
            "e2e": {"value": round(e2e_value, 3), "unit": "images/s",
                    "h2d_bytes_per_step": int(noise.numel() * 4), "d2h_bytes_per_step": int(host_img.numel() * 4)},
            "gpu_launches": int(launches_per_step * args.steps),
            "achieved_tflops_whole_job": round(value / world * GFLOP_PER_IMAGE / 1e3, 2),
            "roofline": {"bound": "tensor", "kernel": "conv_tc kernels (tcgen05 implicit-GEMM circular conv: small-layer, persistent and role-swapped variants)",
                         "achieved": round(achieved, 2), "peak": burst, "unit": "TFLOP/s",
                         "frac": round(achieved / burst, 4), "peak_source": f"{src} bf16 burst (kernel timed alone: device-side stamps around every launch, in-graph)",
                         "frac_of_sustained": round(achieved / sustained, 4), "traffic": conv_traffic(),
                         "algorithmic_bytes": round(conv_nbytes / max(cnt.get("conv_tc", 1), 1)),
                         "traffic_set": {"what": "mean over the conv launches of ONE UNet forward + ONE decode (the set of the ncu capture, cold L2)",
                                         "traffic": conv_traffic(), "algorithmic_bytes": round(set_bytes),
                                         "ratio": round(conv_traffic() / set_bytes, 3) if conv_traffic() and set_bytes else None},
                         "issued_tflops": round(hw_flops / (conv_ms / 1e3) / 1e12, 2) if conv_ms > 0 else 0.0,
                         "launches": cnt.get("conv_tc", 0), "avg_launch_us": round(1e3 * conv_ms / max(cnt.get("conv_tc", 1), 1), 2),
                         "share_of_step": round(conv_ms / all_ms, 4) if all_ms else None,
                         "per_kind_ms_per_step": {k: round(v, 4) for k, v in sorted(ms.items())}},
            "clocks": clk,
        }
        if not args.no_library_baseline and world == 1:
            out["gpu_library_baseline"] = gpu_library_baseline(dev, B)
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_reference(samples=args.cpu_samples)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(out), flush=True)


def cpu_reference(samples=2, threads=None):
    """The fp32 PyTorch oracle of the reference path (restated diffusers UNet2DModel + DPM-Solver++,
    sgm-equivalent Decoder, `ldm/pipelines.py` loop) on the host cores: whole C3 trajectories, batch 1."""
    from oracle import nets, pipeline, schedulers
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    unet = nets.OracleUNet2DModel(**nets.UNET_C3).eval()
    vae = nets.OracleAutoencoderKL().eval()
    sch = schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading")
    g = torch.Generator().manual_seed(0)
    pipeline.ldm_sample(unet, vae, sch, torch.randn((1, 4, 256, 16), generator=g), 2)        # warm-up (2 steps)
    t0 = time.perf_counter()
    for _ in range(samples):
        pipeline.ldm_sample(unet, vae, sch, torch.randn((1, 4, 256, 16), generator=g), STEPS)
    dt = time.perf_counter() - t0
    return {"value": round(samples / dt, 4), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{samples} full C3 images (batch 1, 20 DPM-Solver++ steps + VAE decode), fp32 torch "
                      f"{torch.__version__}, {dt:.1f} s"}


def gpu_library_baseline(dev, B, reps=3):
    """What stock PyTorch 2.11 + cuDNN does with the reference's op graph on the same B200 (VERDICT r1
    item 3): the oracle modules (restated diffusers UNet2DModel / sgm Decoder / DPM-Solver++ loop, `ldm/pipelines.py:
    353-367`) moved to the device, C3 at per-GPU batch `B`, random-init weights.  Variants: strict fp32; PyTorch
    defaults (cudnn.allow_tf32 = True: TF32 convolutions, fp32 matmuls); bf16 autocast.  Each eager (the reference's
    Python loop, ~4000 launches per batch) and captured as ONE CUDA graph (its launch overhead removed)."""
    from oracle import nets, pipeline, schedulers
    torch.manual_seed(0)
    unet = nets.OracleUNet2DModel(**nets.UNET_C3).eval().to(dev)
    vae = nets.OracleAutoencoderKL().eval().to(dev)
    sch = schedulers.OracleDPMSolverMultistepScheduler(timestep_spacing="leading")
    noise = torch.randn((B, 4, 256, 16), device=dev)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    out = {"unit": "images/s", "per_gpu_batch": B, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return round(reps * B / (e0.elapsed_time(e1) / 1e3), 2)

    try:
        torch.backends.cudnn.benchmark = True
        for name, conv_tf32, mm_tf32, ac in (("fp32_strict", False, False, None), ("tf32_default", True, False, None),
                                             ("bf16_autocast", True, True, torch.bfloat16)):
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = conv_tf32, mm_tf32

            def sample():
                if ac is None:
                    return pipeline.ldm_sample(unet, vae, sch, noise, STEPS)
                with torch.autocast("cuda", dtype=ac):
                    return pipeline.ldm_sample(unet, vae, sch, noise, STEPS)
            try:
                out[name + "_eager"] = timed(sample)
            except Exception as e:          # noqa: BLE001
                out[name + "_eager"] = f"failed: {type(e).__name__}: {e}"[:200]
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    sample()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    sample()
                out[name + "_graph"] = timed(g.replay)
                del g
            except Exception as e:          # noqa: BLE001
                out[name + "_graph"] = f"failed: {type(e).__name__}: {e}"[:200]
                torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n = max(1, min(args.steps, 4))
    t_all = time.perf_counter()
    base = cpu_reference(samples=n)
    out = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/s", "n_gpus": args.gpus,
           "steps": n, "warmup": 1, "ms_per_step": round(1e3 / base["value"], 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic N(0,1) noise, random-init weights",
           "config": {"workload": WORKLOAD, "per_gpu_batch": 1, "note": "CPU oracle port of the reference path; "
                      "diffusers is not installable here (SURVEY.md 8c); one step = one full image"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": round(time.perf_counter() - t_all, 1)}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
