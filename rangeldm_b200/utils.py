"""Model surgery with the reference's API (`ldm/utils.py:118-203`): `replace_conv`,
`replace_down`, `replace_attn`, `attn_identity`.

The reference swaps every Conv2d for a circular twin (re-initialising weights, loaded afterwards).
Here circularity is a FLAG the sm_100a conv kernel honours (wrap on azimuth via a modular TMA column
coordinate), so surgery only marks modules in place and weights are preserved.  The reference's
own `ldm/utils.py` functions also work unchanged on these models (see diffusers_compat).
"""
import torch

from . import models


class attn_identity(torch.nn.Identity):
    def forward(self, input, **kwargs):
        return input


def replace_conv(module, name="Conv2d"):
    """Mark every convolution under `module` horizontally circular (`ldm/utils.py:125-146`).  Like the reference --
    which swaps conv-typed ATTRIBUTES and children of the module it is given -- a bare convolution passed as the root
    is left alone: `replace_conv(model.conv_in)` in the `sub_circonv` path (`ldm/inference.py:110-119`) is a no-op."""
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d) and m is not module:
            m.circular = True
    if hasattr(module, "invalidate_plans"):
        module.invalidate_plans()


def replace_down(module, name="down"):
    """`ldm/utils.py:173-203`: Downsample2D convs become circular stride-2 convs; `padding=0`
    (VAE encoder) selects the asymmetric wrap/zero pad of `ldm/utils.py:109-111`."""
    for m in module.modules():
        if isinstance(m, models.Downsample2D):
            m.conv.circular = True
    if hasattr(module, "invalidate_plans"):
        module.invalidate_plans()


def replace_attn(module, name="Attention"):
    """`ldm/utils.py:148-170`: every Attention becomes an identity (attention-free VAE)."""
    for parent in list(module.modules()):
        for child_name, child in list(parent.named_children()):
            if isinstance(child, models.Attention):
                setattr(parent, child_name, attn_identity())
            elif isinstance(child, torch.nn.ModuleList):
                for i, c in enumerate(child):
                    if isinstance(c, models.Attention):
                        child[i] = attn_identity()
    if hasattr(module, "invalidate_plans"):
        module.invalidate_plans()
