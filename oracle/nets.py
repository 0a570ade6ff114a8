"""Oracle networks (TEST INFRASTRUCTURE ONLY): fp32 PyTorch restatement of

* `diffusers.UNet2DModel` as configured by the reference
  (`ldm/train_unconditional.py:237-289`, `ldm/configs/*.yaml:model_config`) after
  `replace_down/replace_conv` surgery (`ldm/utils.py:125-203`) -- SURVEY.md App. A.1;
* `diffusers.AutoencoderKL` as built by `ldm/convert_vae.py:123-189`, whose arithmetic
  is the in-tree `vae/sgm/modules/diffusionmodules/model.py` `Decoder` (:899-1057),
  `Encoder` (:707-896), `ResnetBlock` (:301-362), `Upsample` (:110-125),
  `Downsample` (:148-175) -- SURVEY.md App. A.2.

Tensor layout is the reference's `(B, C, W, H)`: dim 2 = azimuth (circular), dim 3 = beams
(zero padded).  Parameter names are the diffusers state-dict keys.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def circ_conv2d(x, weight, bias, stride=1, padding=1, circular=True):
    """`ldm/utils.py:40-55` == `vae/sgm/.../model.py:93-108`: wrap-pad dim 2, zero-pad dim 3, conv pad 0."""
    if circular:
        p = padding
        if p > 0:
            x = F.pad(x, (0, 0, p, p), mode="circular")
            x = F.pad(x, (p, p, 0, 0), mode="constant")
        return F.conv2d(x, weight, bias, stride, 0)
    return F.conv2d(x, weight, bias, stride, padding)


class CircConv2d(nn.Conv2d):
    def __init__(self, cin, cout, k, stride=1, padding=0, circular=True):
        super().__init__(cin, cout, k, stride=stride, padding=padding)
        self.circular = circular

    def forward(self, x, scale=1.0):
        return circ_conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0], self.circular)


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D (App. A.1) == sgm ResnetBlock (`model.py:342-362`) when temb is None."""

    def __init__(self, cin, cout, temb_ch, groups=32, eps=1e-5, circular=True):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = CircConv2d(cin, cout, 3, padding=1, circular=circular)
        if temb_ch:
            self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = CircConv2d(cout, cout, 3, padding=1, circular=circular)
        if cin != cout:
            self.conv_shortcut = CircConv2d(cin, cout, 1, padding=0, circular=circular)
        self.has_temb = bool(temb_ch)
        self.has_shortcut = cin != cout

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.has_temb and temb is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.has_shortcut:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """diffusers Attention (deprecated-attn-block flavour) with AttnProcessor2_0 (App. A.1)."""

    def __init__(self, ch, head_dim=8, groups=32, eps=1e-5):
        super().__init__()
        self.heads = ch // head_dim
        self.group_norm = nn.GroupNorm(groups, ch, eps=eps)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x, temb=None):
        B, C, W, H = x.shape
        res = x
        h = self.group_norm(x.view(B, C, W * H)).transpose(1, 2)           # (B, N, C), n = w*H + h
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        sp = lambda t: t.view(B, -1, self.heads, C // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))            # scale 1/sqrt(d)
        o = o.transpose(1, 2).reshape(B, -1, C)
        o = self.to_out[0](o)
        o = o.transpose(-1, -2).reshape(B, C, W, H)
        return o + res


class Downsample2D(nn.Module):
    """Patched Downsample2D (`ldm/utils.py:60-116`): padding=1 -> circular conv s2; padding=0 -> asymmetric pad."""

    def __init__(self, ch, padding=1, circular=True):
        super().__init__()
        self.padding = padding
        self.conv = CircConv2d(ch, ch, 3, stride=2, padding=padding, circular=circular)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 0, 0, 1), mode="circular")
            x = F.pad(x, (0, 1, 0, 0), mode="constant")
        return self.conv(x)


class Upsample2D(nn.Module):
    """diffusers Upsample2D(use_conv=True) == sgm Upsample (`model.py:120-125`): nearest 2x then 3x3 conv."""

    def __init__(self, ch, circular=True):
        super().__init__()
        self.conv = CircConv2d(ch, ch, 3, padding=1, circular=circular)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb_ch, layers, attn, add_down, eps, head_dim, circular):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if j == 0 else cout, cout, temb_ch, eps=eps, circular=circular) for j in range(layers)])
        self.attentions = nn.ModuleList([Attention(cout, head_dim, eps=eps) for _ in range(layers)]) if attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 1, circular)]) if add_down else None

    def forward(self, h, temb):
        outs = ()
        for j, r in enumerate(self.resnets):
            h = r(h, temb)
            if self.attentions is not None:
                h = self.attentions[j](h)
            outs += (h,)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs += (h,)
        return h, outs


class UpBlock(nn.Module):
    def __init__(self, cin, prev, cout, temb_ch, layers, attn, add_up, eps, head_dim, circular):
        super().__init__()
        rs = []
        for j in range(layers):
            skip = cin if j == layers - 1 else cout
            rin = prev if j == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, temb_ch, eps=eps, circular=circular))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList([Attention(cout, head_dim, eps=eps) for _ in range(layers)]) if attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout, circular)]) if add_up else None

    def forward(self, h, skips, temb):
        for j, r in enumerate(self.resnets):
            h = r(torch.cat([h, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                h = self.attentions[j](h)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class MidBlock(nn.Module):
    def __init__(self, ch, temb_ch, eps, head_dim, circular, add_attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, eps=eps, circular=circular) for _ in range(2)])
        self.attentions = nn.ModuleList([Attention(ch, head_dim, eps=eps) if add_attention else None])

    def forward(self, h, temb=None):
        h = self.resnets[0](h, temb)
        if self.attentions[0] is not None:
            h = self.attentions[0](h)
        return self.resnets[1](h, temb)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


def sinusoidal_timestep(t, dim, flip_sin_to_cos=True, freq_shift=0.0):
    """diffusers `get_timestep_embedding` as `Timesteps(dim, flip_sin_to_cos=True, freq_shift=0)` calls it:
    cat([cos, sin]) with divisor `half` (App. A.1 step 1).  With flip_sin_to_cos=False, freq_shift=1 it is the
    in-tree `get_timestep_embedding` (`vae/sgm/modules/diffusionmodules/model.py:28-46`: cat([sin, cos]), divisor
    `half - 1`) -- the pin used by tests/test_oracle_cpu.py."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift))
    a = t.float()[:, None] * freqs[None]
    if flip_sin_to_cos:
        return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)
    return torch.cat([torch.sin(a), torch.cos(a)], dim=-1)


UNET_DEFAULTS = dict(
    sample_size=None, in_channels=3, out_channels=3, layers_per_block=2,
    block_out_channels=(224, 448, 672, 896),
    down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
    up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
    attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5, circular=True)

# BASELINE.json configs (SURVEY.md 8d "UNet configs")
UNET_C3 = dict(sample_size=[256, 16], in_channels=5, out_channels=4, layers_per_block=2,
               block_out_channels=[128, 128, 256, 256],
               down_block_types=["DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"],
               up_block_types=["AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"])
UNET_C4 = dict(UNET_C3, sample_size=[256, 8])
UNET_C5 = dict(UNET_C3, in_channels=12)
UNET_C2 = dict(sample_size=[1024, 64], in_channels=3, out_channels=2, layers_per_block=2,
               block_out_channels=[128, 128, 256, 256, 512, 512],
               down_block_types=["DownBlock2D"] * 4 + ["AttnDownBlock2D", "DownBlock2D"],
               up_block_types=["UpBlock2D", "AttnUpBlock2D"] + ["UpBlock2D"] * 4)


class OracleUNet2DModel(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        cfg = dict(UNET_DEFAULTS)
        cfg.update(kw)
        self.cfg = cfg
        boc = list(cfg["block_out_channels"])
        L, eps, hd, circ = cfg["layers_per_block"], cfg["norm_eps"], cfg["attention_head_dim"], cfg["circular"]
        temb_ch = boc[0] * 4
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.conv_in = CircConv2d(cfg["in_channels"], boc[0], 3, padding=1, circular=circ)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, typ in enumerate(cfg["down_block_types"]):
            cin, out = out, boc[i]
            self.down_blocks.append(DownBlock(cin, out, temb_ch, L, typ.startswith("Attn"), i < len(boc) - 1, eps, hd, circ))
        self.mid_block = MidBlock(boc[-1], temb_ch, eps, hd, circ)
        self.up_blocks = nn.ModuleList()
        rb = boc[::-1]
        out = rb[0]
        for i, typ in enumerate(cfg["up_block_types"]):
            prev, out = out, rb[i]
            cin = rb[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(cin, prev, out, temb_ch, L + 1, typ.startswith("Attn"), i < len(boc) - 1, eps, hd, circ))
        self.conv_norm_out = nn.GroupNorm(cfg["norm_num_groups"], boc[0], eps=eps)
        self.conv_out = CircConv2d(boc[0], cfg["out_channels"], 3, padding=1, circular=circ)

    def forward(self, sample, timestep):
        B = sample.shape[0]
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], dtype=torch.long)
        if t.ndim == 0:
            t = t[None]
        if t.device != sample.device:       # (bench.py's GPU library baseline runs these modules on the device)
            t = torch.full((t.numel(),), int(t[0]), dtype=t.dtype, device=sample.device) if t.numel() == 1 else t.to(sample.device)
        t = t * torch.ones(B, dtype=t.dtype, device=sample.device)
        emb = self.time_embedding(sinusoidal_timestep(t, self.cfg["block_out_channels"][0]))
        h = self.conv_in(sample)
        skips = (h,)
        for blk in self.down_blocks:
            h, outs = blk(h, emb)
            skips += outs
        h = self.mid_block(h, emb)
        skips = list(skips)
        for blk in self.up_blocks:
            h = blk(h, skips, emb)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


# ----------------------------------------------------------------------------------------------
# AutoencoderKL (diffusers naming; arithmetic == sgm Encoder/Decoder, attention-free)
VAE_KITTI = dict(in_channels=2, out_channels=2, latent_channels=4, block_out_channels=[64, 128, 256],
                 layers_per_block=2, scaling_factor=0.18215)


class VaeDecoder(nn.Module):
    """`vae/sgm/modules/diffusionmodules/model.py:1024-1057` under diffusers names (`ldm/convert_vae.py:14-121`)."""

    def __init__(self, out_ch, z_ch, boc, layers, eps=1e-6, circular=True):
        super().__init__()
        rb = boc[::-1]
        self.conv_in = CircConv2d(z_ch, rb[0], 3, padding=1, circular=circular)
        self.mid_block = MidBlock(rb[0], 0, eps, 8, circular, add_attention=False)
        self.up_blocks = nn.ModuleList()
        prev = rb[0]
        for i, c in enumerate(rb):
            blk = nn.Module()
            blk.resnets = nn.ModuleList([ResnetBlock2D(prev if j == 0 else c, c, 0, eps=eps, circular=circular)
                                         for j in range(layers + 1)])
            blk.upsamplers = nn.ModuleList([Upsample2D(c, circular)]) if i < len(rb) - 1 else None
            self.up_blocks.append(blk)
            prev = c
        self.conv_norm_out = nn.GroupNorm(32, rb[-1], eps=eps)
        self.conv_out = CircConv2d(rb[-1], out_ch, 3, padding=1, circular=circular)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            for r in blk.resnets:
                h = r(h)
            if blk.upsamplers is not None:
                h = blk.upsamplers[0](h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class VaeEncoder(nn.Module):
    """`vae/sgm/modules/diffusionmodules/model.py:852-896` under diffusers names; Downsample2D(padding=0)."""

    def __init__(self, in_ch, z_ch, boc, layers, eps=1e-6, circular=True):
        super().__init__()
        self.conv_in = CircConv2d(in_ch, boc[0], 3, padding=1, circular=circular)
        self.down_blocks = nn.ModuleList()
        prev = boc[0]
        for i, c in enumerate(boc):
            blk = nn.Module()
            blk.resnets = nn.ModuleList([ResnetBlock2D(prev if j == 0 else c, c, 0, eps=eps, circular=circular)
                                         for j in range(layers)])
            blk.downsamplers = nn.ModuleList([Downsample2D(c, 0, circular)]) if i < len(boc) - 1 else None
            self.down_blocks.append(blk)
            prev = c
        self.mid_block = MidBlock(boc[-1], 0, eps, 8, circular, add_attention=False)
        self.conv_norm_out = nn.GroupNorm(32, boc[-1], eps=eps)
        self.conv_out = CircConv2d(boc[-1], 2 * z_ch, 3, padding=1, circular=circular)

    def forward(self, x):
        h = self.conv_in(x)
        for blk in self.down_blocks:
            for r in blk.resnets:
                h = r(h)
            if blk.downsamplers is not None:
                h = blk.downsamplers[0](h)
        h = self.mid_block(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class OracleAutoencoderKL(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        cfg = dict(VAE_KITTI)
        cfg.update(kw)
        self.cfg = cfg
        boc = list(cfg["block_out_channels"])
        self.encoder = VaeEncoder(cfg["in_channels"], cfg["latent_channels"], boc, cfg["layers_per_block"])
        self.decoder = VaeDecoder(cfg["out_channels"], cfg["latent_channels"], boc, cfg["layers_per_block"])
        self.scaling_factor = cfg["scaling_factor"]

    def decode(self, z):
        return self.decoder(z)

    def encode_moments(self, x):
        return self.encoder(x)

    def encode_sample(self, x, noise):
        """DiagonalGaussianDistribution.sample (`vae/sgm/modules/distributions/distributions.py:24-41`)."""
        mean, logvar = torch.chunk(self.encoder(x), 2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
        return mean + torch.exp(0.5 * logvar) * noise


def sgm_decoder_key_map(n_levels=3, layers=2):
    """diffusers key prefix -> sgm key prefix for the decoder (`ldm/convert_vae.py:25-37,89-120`)."""
    m = {"decoder.conv_in": "conv_in", "decoder.conv_norm_out": "norm_out", "decoder.conv_out": "conv_out",
         "decoder.mid_block.resnets.0": "mid.block_1", "decoder.mid_block.resnets.1": "mid.block_2"}
    for i in range(n_levels):
        for j in range(layers + 1):
            m[f"decoder.up_blocks.{i}.resnets.{j}"] = f"up.{n_levels - 1 - i}.block.{j}"
        m[f"decoder.up_blocks.{i}.upsamplers.0.conv"] = f"up.{n_levels - 1 - i}.upsample.conv"
    return m


def sgm_encoder_key_map(n_levels=3, layers=2):
    m = {"encoder.conv_in": "conv_in", "encoder.conv_norm_out": "norm_out", "encoder.conv_out": "conv_out",
         "encoder.mid_block.resnets.0": "mid.block_1", "encoder.mid_block.resnets.1": "mid.block_2"}
    for i in range(n_levels):
        for j in range(layers):
            m[f"encoder.down_blocks.{i}.resnets.{j}"] = f"down.{i}.block.{j}"
        m[f"encoder.down_blocks.{i}.downsamplers.0.conv"] = f"down.{i}.downsample.conv"
    return m


def to_sgm_state_dict(sd, keymap):
    """Rename a diffusers-named state dict to the sgm names (inverse of `ldm/convert_vae.py`)."""
    out = {}
    for k, v in sd.items():
        for dk in sorted(keymap, key=len, reverse=True):
            if k.startswith(dk + "."):
                rest = k[len(dk) + 1:].replace("conv_shortcut", "nin_shortcut")
                out[keymap[dk] + "." + rest] = v
                break
    return out
