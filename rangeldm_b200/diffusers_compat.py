"""Install `rangeldm_b200` under the import names the reference uses, so `ldm/inference.py`,
`ldm/inference_conditional.py`, `ldm/pipelines.py`, `ldm/utils.py` and `ldm/convert_vae.py` import
unchanged when the real `diffusers` is absent:

    import rangeldm_b200.diffusers_compat as dc; dc.install()
    from diffusers import UNet2DModel, AutoencoderKL, DDPMScheduler          # -> rangeldm_b200 classes
    from diffusers.pipelines.pipeline_utils import DiffusionPipeline, ImagePipelineOutput
    from diffusers.utils import randn_tensor
    diffusers.models.lora.LoRACompatibleConv ...                             # types `replace_*` test for

Import sites covered: `ldm/inference.py:3,17`, `ldm/pipelines.py:5-10`, `ldm/utils.py:1,134,157,182`,
`ldm/train_unconditional.py:30-34` (sampling parts).
"""
import sys
import types

from . import models, pipelines, schedulers


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []          # behave as a package so dotted imports resolve through sys.modules
    sys.modules[name] = m
    return m


def install(force=False):
    """Register the shim as `diffusers` (no-op if a real diffusers is already imported, unless force)."""
    if "diffusers" in sys.modules and not force and not getattr(sys.modules["diffusers"], "__rldm_shim__", False):
        return sys.modules["diffusers"]
    lora = _mod("diffusers.models.lora", LoRACompatibleConv=models.LoRACompatibleConv,
                LoRACompatibleLinear=models.LoRACompatibleLinear)
    resnet = _mod("diffusers.models.resnet", Downsample2D=models.Downsample2D, Upsample2D=models.Upsample2D,
                  ResnetBlock2D=models.ResnetBlock2D)
    attn = _mod("diffusers.models.attention_processor", Attention=models.Attention)
    mdl = _mod("diffusers.models", lora=lora, resnet=resnet, attention_processor=attn,
               UNet2DModel=models.UNet2DModel, AutoencoderKL=models.AutoencoderKL)
    torch_utils = _mod("diffusers.utils.torch_utils", randn_tensor=pipelines.randn_tensor)
    utils = _mod("diffusers.utils", randn_tensor=pipelines.randn_tensor, torch_utils=torch_utils,
                 check_min_version=lambda v: None, is_accelerate_version=lambda *a: False,
                 is_tensorboard_available=lambda: False, is_wandb_available=lambda: False)
    pu = _mod("diffusers.pipelines.pipeline_utils", DiffusionPipeline=pipelines.DiffusionPipeline,
              ImagePipelineOutput=pipelines.ImagePipelineOutput)
    pl = _mod("diffusers.pipelines", pipeline_utils=pu, DiffusionPipeline=pipelines.DiffusionPipeline,
              ImagePipelineOutput=pipelines.ImagePipelineOutput)
    sch = _mod("diffusers.schedulers", DDIMScheduler=schedulers.DDIMScheduler, DDPMScheduler=schedulers.DDPMScheduler,
               DPMSolverMultistepScheduler=schedulers.DPMSolverMultistepScheduler)
    top = _mod("diffusers", models=mdl, utils=utils, pipelines=pl, schedulers=sch,
               UNet2DModel=models.UNet2DModel, AutoencoderKL=models.AutoencoderKL,
               DDIMScheduler=schedulers.DDIMScheduler, DDPMScheduler=schedulers.DDPMScheduler,
               DPMSolverMultistepScheduler=schedulers.DPMSolverMultistepScheduler,
               DiffusionPipeline=pipelines.DiffusionPipeline, ImagePipelineOutput=pipelines.ImagePipelineOutput,
               __version__="0.26.0+rldm", __rldm_shim__=True)
    return top
