"""Oracle sampling loops (TEST INFRASTRUCTURE ONLY): restatement of `ldm/pipelines.py`.

* `ldm_sample`   == `LDMPipelineRange.__call__`        (`ldm/pipelines.py:329-373`)
* `pixel_sample` == `DDIMPipelineRange.__call__`       (`ldm/pipelines.py:224-248`)
* `upscale_sample` == `LDMUpscalePipelineRange.__call__` (`ldm/pipelines.py:466-509`)
* `sparse_encoder2` == `SparseRangeImageEncoder2.forward` (`ldm/encoders.py:86-95`)

All take the initial noise explicitly (the reference draws it with `randn_tensor` on the CPU and
moves it to the device, `ldm/pipelines.py:329-333`) so both sides see identical inputs.
"""
import torch


def pos_encoding_like(x):
    """`ldm/pipelines.py:346-349`: zeros(B,1,W,H) with azimuth row w=0 set to one."""
    pe = torch.zeros([x.shape[0], 1, x.shape[2], x.shape[3]], dtype=x.dtype, device=x.device)
    pe[:, :, 0, :] = 1
    return pe


@torch.no_grad()
def ldm_sample(unet, vae, scheduler, noise, num_inference_steps, pos_encoding=True, variance_noise=None,
               return_latents=False):
    latents = noise * scheduler.init_noise_sigma
    scheduler.set_timesteps(num_inference_steps)
    pe = pos_encoding_like(latents) if pos_encoding else None
    traj = []
    for i, t in enumerate(scheduler.timesteps):
        x = scheduler.scale_model_input(latents, t)
        if pe is not None:
            x = torch.cat([x, pe], dim=1)
        eps = unet(x, t)
        if variance_noise is not None:
            latents = scheduler.step(eps, t, latents, variance_noise=variance_noise[i])
        else:
            latents = scheduler.step(eps, t, latents)
        traj.append(latents)
    latents = latents / vae.scaling_factor
    image = vae.decode(latents)
    return (image, traj) if return_latents else image


@torch.no_grad()
def pixel_sample(unet, scheduler, noise, num_inference_steps, pos_encoding=True, eta=0.0):
    image = noise
    scheduler.set_timesteps(num_inference_steps)
    pe = pos_encoding_like(image) if pos_encoding else None
    for t in scheduler.timesteps:
        x = torch.cat([image, pe], dim=1) if pe is not None else image
        eps = unet(x, t)
        image = scheduler.step(eps, t, image, eta=eta) if eta is not None else scheduler.step(eps, t, image)
    return image


def sparse_encoder2(x, factor=4):
    """(B,C,W,H) -> (B,factor*C,W/factor,H): neighbouring azimuth columns become channels."""
    B, C, W, H = x.shape
    x = torch.flatten(x.permute(0, 2, 1, 3), start_dim=1, end_dim=2)          # (B, W*C, H): index w*C + c
    return x.reshape(B, W // factor, C * factor, H).permute(0, 2, 1, 3)       # channel = (w % factor)*C + c


@torch.no_grad()
def upscale_sample(unet, vae, scheduler, noise, condition, num_inference_steps):
    latents = noise * scheduler.init_noise_sigma
    assert unet.cfg["in_channels"] == latents.shape[1] + condition.shape[1]
    scheduler.set_timesteps(num_inference_steps)
    for t in scheduler.timesteps:
        x = torch.cat([scheduler.scale_model_input(latents, t), condition], dim=1)
        latents = scheduler.step(unet(x, t), t, latents)
    return vae.decode(latents / vae.scaling_factor)
