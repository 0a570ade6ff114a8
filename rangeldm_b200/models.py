"""`UNet2DModel` and `AutoencoderKL` with the diffusers call surface the reference uses
(SURVEY.md 8b), executing on hand-written sm_100a kernels through librldm.so.

The `nn.Module` trees below are PARAMETER HOLDERS with the diffusers state-dict key names
 so `safetensors.torch.load_model`, `load_state_dict`, and the reference's
surgery (`ldm/utils.py:125-203`: `replace_conv`, `replace_down`, `replace_attn`) work on them
unchanged.  None of the leaf modules is ever called: `forward` compiles the tree into a flat
kernel program (`rangeldm_b200/engine.py`) on first use and replays it.  There is no PyTorch or
CPU fallback -- without the CUDA library every forward raises.
"""
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from .configuration import ModelMixin


class _NoTorchForward:
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("rangeldm_b200 modules are parameter holders; the forward pass runs in librldm.so "
                           "through the owning UNet2DModel / AutoencoderKL (no PyTorch fallback)")


class LoRACompatibleConv(_NoTorchForward, nn.Conv2d):
    """Stand-in for `diffusers.models.lora.LoRACompatibleConv`, the type `replace_conv` looks for
    (`ldm/utils.py:134`).  `circular` is what the reference's `Conv2d` carries (`ldm/utils.py:37`)."""
    circular = False


class LoRACompatibleLinear(_NoTorchForward, nn.Linear):
    pass


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels=512, groups=32, eps=1e-5):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = LoRACompatibleConv(in_channels, out_channels, 3, stride=1, padding=1)
        self.time_emb_proj = LoRACompatibleLinear(temb_channels, out_channels) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = LoRACompatibleConv(out_channels, out_channels, 3, stride=1, padding=1)
        self.conv_shortcut = (LoRACompatibleConv(in_channels, out_channels, 1, stride=1, padding=0)
                              if in_channels != out_channels else None)


class Attention(_NoTorchForward, nn.Module):
    """`diffusers.models.attention_processor.Attention` as UNet2DModel builds it (App. A.1)."""

    def __init__(self, query_dim, heads, dim_head, norm_num_groups=32, eps=1e-5):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps)
        self.to_q = LoRACompatibleLinear(query_dim, heads * dim_head)
        self.to_k = LoRACompatibleLinear(query_dim, heads * dim_head)
        self.to_v = LoRACompatibleLinear(query_dim, heads * dim_head)
        self.to_out = nn.ModuleList([LoRACompatibleLinear(heads * dim_head, query_dim), nn.Dropout(0.0)])


class Downsample2D(_NoTorchForward, nn.Module):
    """`diffusers.models.resnet.Downsample2D` (attributes read by `replace_down`, `ldm/utils.py:182-187`)."""

    def __init__(self, channels, use_conv=True, out_channels=None, padding=1, name="conv"):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.padding, self.name = use_conv, padding, name
        if not use_conv:
            raise NotImplementedError("Downsample2D(use_conv=False) is not used by the reference configs")
        conv = LoRACompatibleConv(channels, self.out_channels, 3, stride=2, padding=padding)
        if name == "conv":
            self.Conv2d_0 = conv
        self.conv = conv


class Upsample2D(_NoTorchForward, nn.Module):
    def __init__(self, channels, use_conv=True, out_channels=None):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.conv = LoRACompatibleConv(channels, self.out_channels, 3, padding=1)


class DownBlock2D(nn.Module):
    has_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample, eps, groups,
                 attention_head_dim=8, downsample_padding=1):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels,
                                                    temb_channels, groups, eps) for i in range(num_layers)])
        if self.has_attention:
            self.attentions = nn.ModuleList([Attention(out_channels, out_channels // attention_head_dim,
                                                       attention_head_dim, groups, eps) for _ in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, downsample_padding, "op")])
                             if add_downsample else None)


class AttnDownBlock2D(DownBlock2D):
    has_attention = True


class UpBlock2D(nn.Module):
    has_attention = False

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers, add_upsample,
                 eps, groups, attention_head_dim=8):
        super().__init__()
        rs = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            rs.append(ResnetBlock2D(rin + skip, out_channels, temb_channels, groups, eps))
        self.resnets = nn.ModuleList(rs)
        if self.has_attention:
            self.attentions = nn.ModuleList([Attention(out_channels, out_channels // attention_head_dim,
                                                       attention_head_dim, groups, eps) for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None


class AttnUpBlock2D(UpBlock2D):
    has_attention = True


class UNetMidBlock2D(nn.Module):
    def __init__(self, in_channels, temb_channels, eps, groups, attention_head_dim=8, add_attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps)
                                      for _ in range(2)])
        hd = attention_head_dim if attention_head_dim is not None else in_channels
        self.attentions = nn.ModuleList([Attention(in_channels, in_channels // hd, hd, groups, eps)
                                         if add_attention else None])


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = LoRACompatibleLinear(in_channels, time_embed_dim)
        self.linear_2 = LoRACompatibleLinear(time_embed_dim, time_embed_dim)


@dataclass
class UNet2DOutput:
    sample: torch.Tensor


@dataclass
class DecoderOutput:
    sample: torch.Tensor


_DOWN = {"DownBlock2D": DownBlock2D, "AttnDownBlock2D": AttnDownBlock2D}
_UP = {"UpBlock2D": UpBlock2D, "AttnUpBlock2D": AttnUpBlock2D}


class _PlannedModel(ModelMixin, nn.Module):
    """Shared plan cache: plans are keyed by (batch, spatial size) and dropped whenever weights or
    the module tree may have changed (load_state_dict / .to / surgery before first use)."""

    def _init_plans(self):
        object.__setattr__(self, "_plans", {})
        object.__setattr__(self, "_packed", {})     # repacked weights shared by all plans of this model
        object.__setattr__(self, "_plan_version", 0)

    def invalidate_plans(self):
        """Drop every compiled plan and repacked weight copy, and bump `_plan_version` (part of the pipelines'
        FusedSampler cache key, so a captured trajectory graph is never replayed over stale weights).  Called by
        `load_state_dict`, `safetensors.torch.load_model`, `.to()` and the surgery helpers; IN-PLACE parameter edits
        (e.g. EMA `copy_to`) are invisible to it and need an explicit `invalidate_plans()`."""
        self._plans.clear()
        self._packed.clear()
        object.__setattr__(self, "_plan_version", self._plan_version + 1)

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate_plans()
        return r

    def _load_from_state_dict(self, *a, **k):   # reached by safetensors.torch.load_model as well
        self.invalidate_plans()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate_plans()
        return r

    def _require_cuda(self, x):
        if not x.is_cuda:
            raise RuntimeError("rangeldm_b200 runs on CUDA (sm_100a) only: input tensor is on %s; "
                               "there is no CPU fallback" % x.device)
        if self.device != x.device:
            raise RuntimeError(f"model is on {self.device}, input on {x.device}")


class UNet2DModel(_PlannedModel):
    """Drop-in for `diffusers.UNet2DModel` (constructor keys: `ldm/train_unconditional.py:237-289`)."""

    def __init__(self, sample_size: Optional[Union[int, Tuple[int, int]]] = None, in_channels: int = 3,
                 out_channels: int = 3, center_input_sample: bool = False, time_embedding_type: str = "positional",
                 freq_shift: int = 0, flip_sin_to_cos: bool = True,
                 down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
                 up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels=(224, 448, 672, 896), layers_per_block: int = 2,
                 mid_block_scale_factor: float = 1, downsample_padding: int = 1, downsample_type: str = "conv",
                 upsample_type: str = "conv", dropout: float = 0.0, act_fn: str = "silu",
                 attention_head_dim: Optional[int] = 8, norm_num_groups: int = 32, norm_eps: float = 1e-5,
                 resnet_time_scale_shift: str = "default", add_attention: bool = True,
                 class_embed_type: Optional[str] = None, num_class_embeds: Optional[int] = None):
        nn.Module.__init__(self)
        self._init_plans()
        self._capture_init(locals())
        unsupported = [(center_input_sample, False), (time_embedding_type, "positional"), (freq_shift, 0),
                       (flip_sin_to_cos, True), (mid_block_scale_factor, 1), (downsample_type, "conv"),
                       (upsample_type, "conv"), (act_fn, "silu"), (resnet_time_scale_shift, "default"),
                       (class_embed_type, None), (attention_head_dim, 8), (dropout, 0.0)]
        for got, want in unsupported:
            if got != want:
                raise NotImplementedError(f"UNet2DModel option {got!r} (only {want!r} is used by the reference "
                                          "configs and implemented by the sm_100a engine)")
        boc = list(block_out_channels)
        if len(down_block_types) != len(boc) or len(up_block_types) != len(boc):
            raise ValueError("down_block_types / up_block_types / block_out_channels must have equal length")
        temb_ch = boc[0] * 4
        self.conv_in = LoRACompatibleConv(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, typ in enumerate(down_block_types):
            cin, out = out, boc[i]
            self.down_blocks.append(_DOWN[typ](cin, out, temb_ch, layers_per_block, i < len(boc) - 1, norm_eps,
                                               norm_num_groups, attention_head_dim, downsample_padding))
        self.mid_block = UNetMidBlock2D(boc[-1], temb_ch, norm_eps, norm_num_groups, attention_head_dim,
                                        add_attention)
        self.up_blocks = nn.ModuleList()
        rb = boc[::-1]
        out = rb[0]
        for i, typ in enumerate(up_block_types):
            prev, out = out, rb[i]
            cin = rb[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(_UP[typ](cin, prev, out, temb_ch, layers_per_block + 1, i < len(boc) - 1,
                                           norm_eps, norm_num_groups, attention_head_dim))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = LoRACompatibleConv(boc[0], out_channels, 3, padding=1)

    # ------------------------------------------------------------------------------------------
    def plan(self, batch, W, H, cond_channels=0, replica=0, sampler=False):
        """Compiled kernel program for inputs (batch, in_channels, W, H) (see engine.UNetPlan).  `replica`
        selects an independent set of activation buffers (weights are shared) so several programs of the same
        shape can run concurrently on different streams.  `sampler`: the plan is one step of a fused trajectory
        (its own operand-precision profile, engine.PRECISION)."""
        from .engine import UNetPlan
        key = (batch, W, H, cond_channels, replica, bool(sampler))
        p = self._plans.get(key)
        if p is None:
            p = UNetPlan(self, batch, W, H, cond_channels, sampler=sampler)
            self._plans[key] = p
        return p

    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep, class_labels=None, return_dict: bool = True):
        """`unet(x, t).sample` (`ldm/pipelines.py:103,239,360,500`).  `timestep`: python number, 0-dim
        tensor (CPU or CUDA) or a (B,) tensor."""
        self._require_cuda(sample)
        if sample.dtype != torch.float32:
            raise TypeError("UNet2DModel runs in fp32 at the boundary (reference mixed_precision: 'no')")
        B, C, W, H = sample.shape
        if C != self.config.in_channels:
            raise ValueError(f"expected {self.config.in_channels} input channels, got {C}")
        p = self.plan(B, W, H)
        out = p.run(sample, timestep)
        return UNet2DOutput(sample=out) if return_dict else (out,)


class DiagonalGaussianDistribution:
    """`vae/sgm/modules/distributions/distributions.py:24-41` (== diffusers'): chunk, clamp logvar to
    [-30, 20], sample = mean + std * randn."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None):
        from .pipelines import randn_tensor
        noise = randn_tensor(self.mean.shape, generator=generator, device=self.parameters.device,
                             dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample, eps, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, 0,
                                                    groups, eps) for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, 0, "op")])
                             if add_downsample else None)


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample, eps, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, 0,
                                                    groups, eps) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups,
                 double_z=True, mid_block_add_attention=True):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = LoRACompatibleConv(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i in range(len(boc)):
            cin, out = out, boc[i]
            self.down_blocks.append(DownEncoderBlock2D(cin, out, layers_per_block, i < len(boc) - 1, 1e-6,
                                                       norm_num_groups))
        self.mid_block = UNetMidBlock2D(boc[-1], 0, 1e-6, norm_num_groups, None, mid_block_add_attention)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = LoRACompatibleConv(boc[-1], 2 * out_channels if double_z else out_channels, 3, padding=1)


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups,
                 mid_block_add_attention=True):
        super().__init__()
        rb = list(block_out_channels)[::-1]
        self.conv_in = LoRACompatibleConv(in_channels, rb[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(rb[0], 0, 1e-6, norm_num_groups, None, mid_block_add_attention)
        self.up_blocks = nn.ModuleList()
        out = rb[0]
        for i in range(len(rb)):
            prev, out = out, rb[i]
            self.up_blocks.append(UpDecoderBlock2D(prev, out, layers_per_block + 1, i < len(rb) - 1, 1e-6,
                                                   norm_num_groups))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, rb[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = LoRACompatibleConv(rb[-1], out_channels, 3, padding=1)


class AutoencoderKL(_PlannedModel):
    """Drop-in for `diffusers.AutoencoderKL` as built by `ldm/convert_vae.py:123-189`; arithmetic of
    `vae/sgm/modules/diffusionmodules/model.py` `Decoder` (:1024-1057) / `Encoder` (:852-896)."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, down_block_types=("DownEncoderBlock2D",),
                 up_block_types=("UpDecoderBlock2D",), block_out_channels=(64,), layers_per_block: int = 1,
                 act_fn: str = "silu", latent_channels: int = 4, norm_num_groups: int = 32, sample_size=32,
                 scaling_factor: float = 0.18215, force_upcast: bool = True):
        nn.Module.__init__(self)
        self._init_plans()
        self._capture_init(locals())
        if act_fn != "silu":
            raise NotImplementedError("AutoencoderKL act_fn must be 'silu'")
        if any(t != "DownEncoderBlock2D" for t in down_block_types) or any(t != "UpDecoderBlock2D" for t in up_block_types):
            raise NotImplementedError("only DownEncoderBlock2D / UpDecoderBlock2D VAEs are supported")
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = LoRACompatibleConv(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = LoRACompatibleConv(latent_channels, latent_channels, 1)

    def _plan(self, kind, batch, W, H, replica=0):
        from .engine import VaeDecoderPlan, VaeEncoderPlan
        key = (kind, batch, W, H, replica)
        p = self._plans.get(key)
        if p is None:
            p = (VaeDecoderPlan if kind == "dec" else VaeEncoderPlan)(self, batch, W, H)
            self._plans[key] = p
        return p

    def decoder_plan(self, batch, W, H, replica=0):
        return self._plan("dec", batch, W, H, replica)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        """`vae.decode(latents).sample` (`ldm/pipelines.py:355,367,507`)."""
        self._require_cuda(z)
        B, C, W, H = z.shape
        out = self._plan("dec", B, W, H).run(z.float())
        return DecoderOutput(sample=out) if return_dict else (out,)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """`vae.encode(image).latent_dist.sample()` (`ldm/pipelines.py:408`)."""
        self._require_cuda(x)
        B, C, W, H = x.shape
        moments = self._plan("enc", B, W, H).run(x.float())
        dist = DiagonalGaussianDistribution(moments)
        return AutoencoderKLOutput(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def forward(self, sample, sample_posterior=False, return_dict=True, generator=None):
        """`vae(x).sample` (`ldm/convert_vae.py:231`)."""
        post = self.encode(sample).latent_dist
        z = post.sample(generator=generator) if sample_posterior else post.mode()
        return self.decode(z, return_dict=return_dict)
