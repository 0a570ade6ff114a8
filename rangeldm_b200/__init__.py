"""rangeldm_b200 -- B200-native (sm_100a) range-image latent-diffusion sampler behind the call
surface WoodwindHu/RangeLDM uses (UNet2DModel / AutoencoderKL / schedulers / ldm pipelines).
See DESIGN.md for the path, the boundary and the kernels; include/rldm.h for the C ABI."""
from .models import UNet2DModel, AutoencoderKL, UNet2DOutput, DecoderOutput, DiagonalGaussianDistribution  # noqa: F401
from .schedulers import DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler, SchedulerOutput  # noqa: F401
from .pipelines import (DiffusionPipeline, ImagePipelineOutput, DDPMPipelineRange, DDIMPipelineRange,  # noqa: F401
                        LDMPipelineRange, LDMUpscalePipelineRange, FusedSampler, randn_tensor)
from .utils import replace_conv, replace_down, replace_attn, attn_identity  # noqa: F401
from .encoders import SparseRangeImageEncoder2  # noqa: F401
from .geometry import RangeImageGeometry  # noqa: F401

__version__ = "0.1.0"
