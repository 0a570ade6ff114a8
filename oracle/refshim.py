"""Read-only import of the reference's own VAE / sampler code (TEST INFRASTRUCTURE ONLY).

`/root/reference` exists only in the build container, never on the GPU box: nothing under
`-m gpu` tests, `smoke()` or `bench.py` may call this at run time.  It is used by
`oracle/make_golden.py` (fixture generation) and by CPU tests that are skipped when the
reference tree is absent.

Recipe: pre-register empty `sgm`, `sgm.modules`,
`sgm.modules.diffusionmodules` packages (bypassing their `__init__.py`, which pull in
pytorch_lightning) and a names-only `omegaconf` stub, then import
`vae/sgm/modules/diffusionmodules/{model,sampling,discretizer}.py` unchanged.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("RLDM_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "vae", "sgm"))


def load():
    """Return (model, sampling, discretizer) modules of the reference."""
    root = os.path.join(REF_ROOT, "vae", "sgm")
    for name, path in (("sgm", root), ("sgm.modules", root + "/modules"),
                       ("sgm.modules.diffusionmodules", root + "/modules/diffusionmodules")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        oc.ListConfig = list
        oc.OmegaConf = dict
        sys.modules["omegaconf"] = oc
    model = importlib.import_module("sgm.modules.diffusionmodules.model")
    sampling = importlib.import_module("sgm.modules.diffusionmodules.sampling")
    disc = importlib.import_module("sgm.modules.diffusionmodules.discretizer")
    return model, sampling, disc


def make_decoder(model, ch=64, ch_mult=(1, 2, 4), z_channels=4, num_res_blocks=2):
    """Decoder as configured by `vae/configs/kitti360.yaml:47-62`."""
    return model.Decoder(attn_type="none", double_z=True, z_channels=z_channels, resolution=256, in_channels=2,
                         out_ch=2, ch=ch, ch_mult=list(ch_mult), num_res_blocks=num_res_blocks,
                         attn_resolutions=[], dropout=0.0, act="silu", circular=True)


def make_encoder(model, ch=64, ch_mult=(1, 2, 4), z_channels=4, num_res_blocks=2):
    """Encoder as configured by `vae/configs/kitti360.yaml:30-45`."""
    return model.Encoder(attn_type="none", double_z=True, z_channels=z_channels, resolution=256, in_channels=2,
                         out_ch=2, ch=ch, ch_mult=list(ch_mult), num_res_blocks=num_res_blocks,
                         attn_resolutions=[], dropout=0.0, act="silu", circular=True)
