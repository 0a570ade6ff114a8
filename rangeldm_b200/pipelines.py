"""Sampling pipelines with the call surface of the reference's `ldm/pipelines.py`
(`DDPMPipelineRange`, `DDIMPipelineRange`, `LDMPipelineRange`, `LDMUpscalePipelineRange`) plus the
slice of `diffusers.DiffusionPipeline` they rely on.

What is different underneath: the reference runs a Python loop of ~200 library launches per step.
Here `FusedSampler` compiles the WHOLE trajectory -- N x (time-embedding, UNet, scheduler step),
the 1/scaling_factor rescale and the VAE decode -- into one flat librldm program over static HBM
buffers, captures it in a CUDA graph and replays it per batch; the host only uploads noise and
reads images.  The per-step module API (`unet(x, t).sample`, `scheduler.step(...)`) remains
available and is what `final_only=False` and the reference's own `ldm/pipelines.py` use.
"""
import inspect
import json
import os
from dataclasses import dataclass
from typing import List, Optional, Union

import numpy as np
import torch

from . import _lib
from ._lib import RldmOp
from .engine import Program
from .schedulers import DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.randn_tensor semantics: a CPU generator (or no device)
    draws on the CPU and moves; a list of generators draws per sample.  RNG stays in PyTorch."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    batch = shape[0]
    if generator is not None:
        gdev = generator[0].device.type if isinstance(generator, (list, tuple)) else generator.device.type
        if gdev != device.type and gdev == "cpu":
            rand_device = torch.device("cpu")
        elif gdev != device.type and gdev == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {gdev}.")
    if isinstance(generator, (list, tuple)) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, (list, tuple)):
        one = (1,) + tuple(shape[1:])
        latents = torch.cat([torch.randn(one, generator=generator[i], device=rand_device, dtype=dtype)
                             for i in range(batch)], dim=0).to(device)
    else:
        latents = torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)
    return latents


@dataclass
class ImagePipelineOutput:
    images: Union[List, np.ndarray]


class DiffusionPipeline:
    """The members of diffusers.DiffusionPipeline the reference touches (`ldm/pipelines.py:9,30-32,96,
    101,224`; `ldm/inference.py:139,155`; `ldm/train_unconditional.py:675`)."""
    config_name = "model_index.json"

    def __init__(self):
        self._modules_registered = {}
        self._progress_bar_config = {}

    def register_modules(self, **kwargs):
        if not hasattr(self, "_modules_registered"):
            DiffusionPipeline.__init__(self)
        for name, module in kwargs.items():
            self._modules_registered[name] = module
            setattr(self, name, module)

    @property
    def components(self):
        return dict(self._modules_registered)

    def to(self, device=None, dtype=None):
        for m in self._modules_registered.values():
            if isinstance(m, torch.nn.Module):
                m.to(device)
        return self

    @property
    def device(self):
        for m in self._modules_registered.values():
            if isinstance(m, torch.nn.Module):
                return next(m.parameters()).device
        return torch.device("cpu")

    @property
    def _execution_device(self):
        return self.device

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def progress_bar(self, iterable=None, total=None):
        cfg = dict(getattr(self, "_progress_bar_config", {}))
        try:
            from tqdm.auto import tqdm
        except Exception:  # pragma: no cover
            return iterable
        if iterable is not None:
            return tqdm(iterable, **cfg)
        return tqdm(total=total, **cfg)

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        if images.shape[-1] == 1:
            return [Image.fromarray(im.squeeze(), mode="L") for im in images]
        return [Image.fromarray(im) for im in images]

    def save_pretrained(self, save_directory, **kw):
        os.makedirs(save_directory, exist_ok=True)
        index = {"_class_name": type(self).__name__}
        for name, m in self._modules_registered.items():
            sub = os.path.join(save_directory, name)
            if hasattr(m, "save_pretrained"):
                m.save_pretrained(sub)
            elif hasattr(m, "save_config"):
                m.save_config(sub)
            index[name] = [type(m).__module__.split(".")[0], type(m).__name__]
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump(index, f, indent=2)


# ------------------------------------------------------------------------------------------------
def make_pos_encoding(batch, W, H, device):
    """`ldm/pipelines.py:229-232,346-349`: one extra channel, azimuth row w=0 set to one."""
    pe = torch.zeros([batch, 1, W, H], device=device)
    pe[:, :, 0, :] = 1
    return pe


class _Trajectory:
    """One whole trajectory program for a sub-batch: N x (UNet + scheduler.step) [+ rescale + vae.decode] as a
    flat librldm program over static buffers (`latents`, `cond`, `noise`, `image`)."""

    def __init__(self, unet, scheduler, vae, batch, cond_channels, replica=0, fuse=True):
        dev = unet.device
        cfg = unet.config
        W, H = cfg.sample_size if not isinstance(cfg.sample_size, int) else (cfg.sample_size, cfg.sample_size)
        self.B, self.steps = batch, len(scheduler.timesteps)
        self.plan = unet.plan(batch, W, H, cond_channels, replica=replica, sampler=True)
        self.latents, self.cond = self.plan.x_in, self.plan.cond
        pg = self.prog = Program(dev, fuse=fuse)
        pg.keep += [self.plan, scheduler]
        coef = pg.hold(scheduler.coef_table(dev).clone())
        n = self.latents.numel()
        x0buf = pg.hold(torch.zeros_like(self.latents))
        host_coef = scheduler._coef_host
        self.noise = None
        if bool((host_coef[:, 6] != 0).any()):
            self.noise = pg.hold(torch.zeros((self.steps,) + tuple(self.latents.shape), device=dev))
        uses_prev = bool((host_coef[:, 4] != 0).any())
        # The time-embedding MLP and all per-resnet projections depend only on the timestep table: the FIRST op of the
        # trajectory evaluates them for every step in one batched pass (rldm_temb over `steps` rows, inside the
        # captured graph and the timed step), and step i's conv epilogues read row i for every image of the batch
        # (temb_stride 0) -- instead of 20 x (MLP + 22 projections) spread over the loop.
        T = self.plan.temb_T
        tvec = pg.hold(scheduler.timesteps.to(dev, torch.float32).contiguous())
        temb_all = pg.hold(torch.zeros(self.steps, T, device=dev))
        temb_base = self.plan.temb_out.data_ptr()
        D0, D4 = self.plan.temb_dims
        temb_scratch = pg.hold(torch.zeros(2, self.steps, D4, device=dev))
        for op in self.plan.prog.ops:
            if op.kind == _lib.OP_TEMB:
                cp = RldmOp.from_buffer_copy(op)
                cp.i[0] = self.steps
                cp.p[0], cp.p[7], cp.p[8] = tvec.data_ptr(), temb_scratch.data_ptr(), temb_all.data_ptr()
                pg.append(cp, 3)
        for i in range(self.steps):
            for op, nl in zip(self.plan.prog.ops, self.plan.prog.launches):
                if op.kind == _lib.OP_TEMB:
                    continue
                cp = RldmOp.from_buffer_copy(op)
                if cp.kind in (_lib.OP_CONV_TC, _lib.OP_CONV_REF) and cp.p[3]:
                    cp.p[3] = temb_all[i].data_ptr() + (cp.p[3] - temb_base)
                    cp.i[0] = 0                       # every image of the batch shares the step's row
                pg.append(cp, nl)
            noise_i = self.noise[i] if (self.noise is not None and float(host_coef[i, 6]) != 0.0) else None
            pg.add(_lib.OP_SCHED_STEP, p=(coef[i], self.latents, self.plan.out,
                                          x0buf if (uses_prev and i > 0) else None, noise_i, self.latents,
                                          x0buf if uses_prev else None), n=n)
        if vae is not None:
            self.dec = vae.decoder_plan(batch, W, H, replica=replica)
            pg.keep.append(self.dec)
            pg.add(_lib.OP_AXPY, f=(1.0 / float(vae.config.scaling_factor),), p=(self.latents, self.dec.z_in), n=n)
            pg.extend(self.dec.prog)
            self.image = self.dec.out
        else:
            self.dec = None
            self.image = self.latents
        pg.finalize()


class FusedSampler:
    """The whole sampling job of one batch as ONE CUDA graph.

    The batch is split into `streams` independent sub-batches (images never interact), each with its own
    trajectory program and activation buffers (weights are shared); the programs are captured on parallel
    branches of the graph.  Measured on B200 (C3, batch 8): 1 stream 136 img/s, 2 streams 131, 4 streams 121 --
    the smaller tiles cost more than the overlap gains, so the default is ONE program; RLDM_STREAMS=n opts in."""

    def __init__(self, unet, scheduler, vae, batch, cond_channels, use_graph=True, streams=None):
        cfg = unet.config
        if cfg.in_channels != cfg.out_channels + cond_channels:
            raise AssertionError(f"unet.in_channels {cfg.in_channels} != out_channels {cfg.out_channels} + "
                                 f"condition channels {cond_channels}")
        if streams is None:
            streams = int(os.environ.get("RLDM_STREAMS", "1"))
        streams = max(1, min(streams, batch))
        while batch % streams:
            streams -= 1
        sub = batch // streams
        self.B, self.steps = batch, len(scheduler.timesteps)
        # sub-batch programs on parallel graph branches run CONCURRENTLY: no fused-levels launches there (each one
        # needs every SM)
        self.parts = [_Trajectory(unet, scheduler, vae, sub, cond_channels, replica=r, fuse=streams == 1)
                      for r in range(streams)]
        self.slices = [slice(r * sub, (r + 1) * sub) for r in range(streams)]
        self.noise = self.parts[0].noise          # None for deterministic schedulers
        self.has_cond = self.parts[0].cond is not None
        self.gpu_launches = sum(t.prog.n_launch for t in self.parts)
        self.graph = None
        self._side = [torch.cuda.Stream() for _ in range(streams - 1)] if unet.device.type == "cuda" else []
        if use_graph:
            self._capture()

    # -- single-program views kept for callers that only look at the first sub-batch (profiling) --
    @property
    def plan(self):
        return self.parts[0].plan

    @property
    def dec(self):
        return self.parts[0].dec

    def _launch_all(self):
        """Issue every sub-batch program: part 0 on the current stream, the others on side streams (fork/join)."""
        cur = torch.cuda.current_stream()
        for st in self._side:
            st.wait_stream(cur)
        for t, st in zip(self.parts[1:], self._side):
            with torch.cuda.stream(st):
                t.prog.run()
        self.parts[0].prog.run()
        for st in self._side:
            cur.wait_stream(st)

    def _capture(self):
        saved = [t.latents.clone() for t in self.parts]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._launch_all()                    # warm-up: sets function attributes, sizes workspaces
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch_all()
        self.graph = g
        for t, v in zip(self.parts, saved):
            t.latents.copy_(v)

    def load(self, latents, cond=None, noise=None):
        """Copy a batch of inputs into the static buffers of the sub-batch programs."""
        for t, sl in zip(self.parts, self.slices):
            t.latents.copy_(latents[sl])
            if t.cond is not None:
                t.cond.copy_(cond[sl])
            if t.noise is not None:
                t.noise.copy_(noise[:, sl])

    def replay(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch_all()

    def result(self):
        """Fresh (B, C, W, H) tensor with the finished images (or latents when there is no VAE)."""
        if len(self.parts) == 1:
            return self.parts[0].image.clone()
        return torch.cat([t.image for t in self.parts], dim=0)

    def run(self, latents, cond=None, noise=None):
        self.load(latents, cond, noise)
        self.replay()
        return self.result()


def _is_native_scheduler(s):
    return isinstance(s, (DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler))


class _RangePipeline(DiffusionPipeline):
    def _sample_size(self):
        ss = self.unet.config.sample_size
        return (ss, ss) if isinstance(ss, int) else tuple(ss)

    def _sampler(self, batch, cond_channels, vae):
        """FusedSampler cache: keyed by everything that is baked into the program."""
        sch = self.scheduler
        ver = lambda m: None if m is None else (id(m), getattr(m, "_plan_version", 0), str(m.device))
        key = (batch, cond_channels, type(sch).__name__, tuple(sch.timesteps.tolist()),
               tuple(sch._coef_host.flatten().tolist()), ver(self.unet), ver(vae))
        cache = self.__dict__.setdefault("_fused", {})
        if key not in cache:
            cache.clear()                     # one live trajectory program per pipeline
            cache[key] = FusedSampler(self.unet, sch, vae, batch, cond_channels)
        return cache[key]

    # A fused trajectory unrolls every step into one program / CUDA graph and, for stochastic schedulers, holds the
    # per-step noise of the whole trajectory in HBM.  Long ancestral runs (DDPMPipelineRange defaults to 1000 steps
    # on (B,2,1024,64) pixels: 8.4 GB of noise at B = 16 and several 10^5 graph nodes) take the per-step path, which
    # needs O(1) extra memory like the reference loop.
    MAX_UNROLLED_STEPS = 250
    MAX_UNROLLED_NOISE_BYTES = 1 << 30

    def _fusable(self, steps, numel):
        if not _is_native_scheduler(self.scheduler) or steps > self.MAX_UNROLLED_STEPS:
            return False
        stochastic = bool((self.scheduler._coef_host[:, 6] != 0).any())
        return not (stochastic and steps * numel * 4 > self.MAX_UNROLLED_NOISE_BYTES)

    def _stepwise(self, latents, cond, generator, extra=None):
        """The reference's loop body (`ldm/pipelines.py:101-106,234-246,353-362,496-502`) over the module API."""
        extra = dict(extra or {})
        if "generator" in set(inspect.signature(self.scheduler.step).parameters.keys()):
            extra["generator"] = generator
        for t in self.progress_bar(self.scheduler.timesteps):
            x = self.scheduler.scale_model_input(latents, t)
            if cond is not None:
                x = torch.cat([x, cond], dim=1)
            latents = self.scheduler.step(self.unet(x, t).sample, t, latents, **extra).prev_sample
        return latents

    def _draw_step_noise(self, sampler, generator, shape, device):
        if sampler.noise is None:
            return None
        noise = torch.zeros((sampler.steps,) + tuple(shape), device=device)
        host = self.scheduler._coef_host
        for i in range(sampler.steps):
            if float(host[i, 6]) != 0.0:      # same draw order as the reference's step-by-step loop
                noise[i] = randn_tensor(shape, generator=generator, device=device, dtype=torch.float32)
        return noise

    def _finish(self, image, output_type, return_dict):
        if output_type == "torch":
            return image
        image = (image / 2 + 0.5).clamp(0, 1)
        image = image.cpu().permute(0, 2, 3, 1).numpy()
        if output_type == "pil":
            image = self.numpy_to_pil(image)
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image)

    def _check_generators(self, generator, batch_size):
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(
                f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                f" size of {batch_size}. Make sure the batch size matches the length of the generators.")

    def _set_steps(self, n, eta=0.0):
        """`scheduler.set_timesteps(n)` as the reference loop does on every call (`ldm/pipelines.py:97,230,343,484`).  For
        the native schedulers a repeated call with the same arguments only resets the multistep state: the tables (and
        their device copy) are the ones of the previous call.  Any direct `set_timesteps` by the user re-installs them."""
        sch = self.scheduler
        key = (int(n), float(eta) if isinstance(sch, DDIMScheduler) else 0.0)
        if _is_native_scheduler(sch) and getattr(sch, "_steps_key", None) == key:
            sch._state = {}
            return
        if isinstance(sch, DDIMScheduler):
            sch.set_timesteps(n, eta=eta)
        else:
            sch.set_timesteps(n)
        if _is_native_scheduler(sch):
            sch._steps_key = key


class DDPMPipelineRange(_RangePipeline):
    """Pixel-space ancestral sampling (`ldm/pipelines.py:14-117`); like the reference, takes no
    `pos_encoding` argument."""
    model_cpu_offload_seq = "unet"

    def __init__(self, unet, scheduler):
        super().__init__()
        self.register_modules(unet=unet, scheduler=scheduler)

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, num_inference_steps: int = 1000,
                 output_type: Optional[str] = "torch", return_dict: bool = True):
        W, H = self._sample_size()
        shape = (batch_size, self.unet.config.in_channels, W, H)
        image = randn_tensor(shape, generator=generator, device=self.device)
        self._set_steps(num_inference_steps)
        if not self._fusable(num_inference_steps, image.numel()):
            return self._finish(self._stepwise(image, None, generator), output_type, return_dict)
        sampler = self._sampler(batch_size, 0, None)
        noise = self._draw_step_noise(sampler, generator, shape, self.device)
        return self._finish(sampler.run(image, None, noise), output_type, return_dict)


class DDIMPipelineRange(_RangePipeline):
    """RangeDM pixel pipeline (`ldm/pipelines.py:119-258`): converts the scheduler to DDIM."""
    model_cpu_offload_seq = "unet"

    def __init__(self, unet, scheduler, pos_encoding=False):
        super().__init__()
        if not isinstance(scheduler, DDIMScheduler):
            scheduler = DDIMScheduler.from_config(scheduler.config)     # `ldm/pipelines.py:139`
        self.register_modules(unet=unet, scheduler=scheduler)
        self.pos_encoding = pos_encoding

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, eta: float = 0.0, num_inference_steps: int = 50,
                 use_clipped_model_output: Optional[bool] = None, output_type: Optional[str] = "torch",
                 return_dict: bool = True):
        W, H = self._sample_size()
        shape = (batch_size, self.unet.config.out_channels, W, H)
        self._check_generators(generator, batch_size)
        image = randn_tensor(shape, generator=generator, device=self._execution_device, dtype=self.unet.dtype)
        self._set_steps(num_inference_steps, eta)
        cc = 1 if self.pos_encoding else 0
        cond = make_pos_encoding(batch_size, W, H, self.device) if cc else None
        if not self._fusable(num_inference_steps, image.numel()):
            return self._finish(self._stepwise(image, cond, generator, {"eta": eta}), output_type, return_dict)
        sampler = self._sampler(batch_size, cc, None)
        noise = self._draw_step_noise(sampler, generator, shape, self.device)
        return self._finish(sampler.run(image, cond, noise), output_type, return_dict)


class LDMPipelineRange(_RangePipeline):
    """RangeLDM latent pipeline (`ldm/pipelines.py:260-383`); scheduler-agnostic like the reference."""

    def __init__(self, vae, unet, scheduler, pos_encoding=False):
        super().__init__()
        self.register_modules(vae=vae, unet=unet, scheduler=scheduler)
        self.pos_encoding = pos_encoding

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, eta: float = 0.0, num_inference_steps: int = 50,
                 output_type: Optional[str] = "torch", return_dict: bool = True, final_only: bool = True, **kwargs):
        W, H = self._sample_size()
        shape = (batch_size, self.unet.config.out_channels, W, H)
        latents = randn_tensor(shape, generator=generator).to(self.device)       # CPU draw, then H2D (:329-333)
        latents = latents * self.scheduler.init_noise_sigma
        accepts_eta = "eta" in set(inspect.signature(self.scheduler.step).parameters.keys())
        self._set_steps(num_inference_steps, eta if accepts_eta else 0.0)
        cc = 1 if self.pos_encoding else 0
        cond = make_pos_encoding(batch_size, W, H, self.device) if cc else None
        if final_only and self._fusable(num_inference_steps, latents.numel()):
            sampler = self._sampler(batch_size, cc, self.vae)
            noise = self._draw_step_noise(sampler, generator, shape, self.device)
            return self._finish(sampler.run(latents, cond, noise), output_type, return_dict)
        # step-by-step path (intermediate decodes, or a foreign scheduler object)
        assert final_only or output_type == "torch"
        extra = {"eta": eta} if accepts_eta else {}
        if "generator" in set(inspect.signature(self.scheduler.step).parameters.keys()):
            extra["generator"] = generator
        image_list = []
        for t in self.progress_bar(self.scheduler.timesteps):
            if not final_only:
                image_list.append(self.vae.decode(latents / self.vae.config.scaling_factor).sample)
            x = self.scheduler.scale_model_input(latents, t)
            if cond is not None:
                x = torch.cat([x, cond], dim=1)
            eps = self.unet(x, t).sample
            latents = self.scheduler.step(eps, t, latents, **extra).prev_sample
        image = self.vae.decode(latents / self.vae.config.scaling_factor).sample
        if output_type == "torch" and not final_only:
            image_list.append(image)
            return image_list
        return self._finish(image, output_type, return_dict)


class LDMUpscalePipelineRange(_RangePipeline):
    """Conditional upsampling / inpainting (`ldm/pipelines.py:385-519`)."""

    def __init__(self, vae, unet, scheduler):
        super().__init__()
        self.register_modules(vae=vae, unet=unet, scheduler=scheduler)

    def encode_masked_image(self, image, mask):
        image = image.to(self.unet.device)
        image = self.vae.encode(image).latent_dist.sample()
        image = image * self.vae.config.scaling_factor
        mask = torch.nn.functional.interpolate(mask.to(self.unet.device), size=image.shape[-2:])
        return torch.cat([image, mask], dim=1)

    @torch.no_grad()
    def __call__(self, image=None, mask=None, condition_encoder=None, batch_size: int = 1,
                 num_inference_steps: int = 100, eta: float = 0.0, generator=None,
                 output_type: Optional[str] = "torch", return_dict: bool = True):
        if image is None:
            raise ValueError("`image` input cannot be undefined.")
        W, H = self._sample_size()
        shape = (batch_size, self.unet.config.out_channels, W, H)
        latents = randn_tensor(shape, generator=generator).to(self.unet.device)
        if mask is None:
            assert condition_encoder is not None
            cond = condition_encoder(image)
        else:
            cond = self.encode_masked_image(image, mask)
        cond = cond.to(dtype=latents.dtype, device=self.unet.device)
        assert self.unet.config.in_channels == self.unet.config.out_channels + cond.shape[1]
        assert cond.shape[2] == W
        assert cond.shape[3] == H
        latents = latents * self.scheduler.init_noise_sigma
        accepts_eta = "eta" in set(inspect.signature(self.scheduler.step).parameters.keys())
        self._set_steps(num_inference_steps, eta if accepts_eta else 0.0)
        if not self._fusable(num_inference_steps, latents.numel()):
            latents = self._stepwise(latents, cond, generator, {"eta": eta} if accepts_eta else {})
            return self._finish(self.vae.decode(latents / self.vae.config.scaling_factor).sample, output_type, return_dict)
        sampler = self._sampler(batch_size, cond.shape[1], self.vae)
        noise = self._draw_step_noise(sampler, generator, shape, self.unet.device)
        return self._finish(sampler.run(latents, cond, noise), output_type, return_dict)
