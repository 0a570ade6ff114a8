"""CPU restatement of the reference's range-image -> point-cloud geometry (TEST INFRASTRUCTURE ONLY).

Follows `ldm/dataset.py:228-276` (`point_cloud_to_range_image.to_pc_torch`) and the writer loop of
`ldm/inference.py:174-179` (depth mask < 90 m, float32 N x 4 `.bin`).  Pinned against the reference's own
`point_cloud_to_range_image_KITTI.to_pc_torch` by `tests/golden/range_to_points.pt` (made by oracle/make_golden.py).
"""
import math

import numpy as np
import torch

MODE_LINEAR, MODE_LOG, MODE_INVERSE = 0, 1, 2


def to_points(range_images, incl, height, mode=MODE_LINEAR, mean=20.0, std=40.0, fill=100.0):
    """range_images (B, C, W, H) fp32 -> (B, W*H, 3 or 4): x, y, z [, remission]   (`ldm/dataset.py:228-276`)."""
    B, C, W, H = range_images.shape
    incl_t = torch.as_tensor(incl, dtype=torch.float32)
    height_t = torch.as_tensor(height, dtype=torch.float32)
    r0 = range_images[:, 0].to(torch.float32)
    if mode == MODE_LOG:                                  # `:241-242`
        r = 2 ** (r0 * 6) - 1
    elif mode == MODE_INVERSE:                            # `:243-244`
        r = 1 / torch.max(r0, torch.tensor([0.0001]))
    else:                                                 # `:245-246`
        r = r0 * std + mean
    r = torch.where(r < 0, torch.full_like(r, fill), r)   # `:256`  r_true[r_true<0] = range_fill_value[0]
    z = (height_t[None, None, :] - r * torch.sin(incl_t[None, None, :])).reshape(B, W * H)      # `:259`
    xy = r * torch.cos(incl_t[None, None, :])                                                 # `:262`
    azi = (W - 0.5 - torch.arange(0, W)) / W * 2.0 * math.pi - math.pi                        # `:266`
    x = (xy * torch.cos(azi[None, :, None])).reshape(B, W * H)                                   # `:269`
    y = (xy * torch.sin(azi[None, :, None])).reshape(B, W * H)                                   # `:270`
    cols = [x, y, z]
    if C > 1:
        cols.append(range_images[:, 1].reshape(B, W * H).to(torch.float32))                      # `:248,273`
    return torch.stack(cols, dim=2)


def depth_masked(points, max_depth=90.0):
    """One sample (N, 3|4) -> the rows the reference writes to `<index>.bin` (`ldm/inference.py:175-179`)."""
    pc = points.detach().cpu().numpy()
    depth = np.linalg.norm(pc[:, :3], 2, axis=1)
    return pc[depth < max_depth, :]
