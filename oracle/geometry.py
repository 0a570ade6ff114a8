"""CPU restatement of the reference's range-image -> point-cloud geometry (TEST INFRASTRUCTURE ONLY).

Follows `ldm/dataset.py:228-276` (`point_cloud_to_range_image.to_pc_torch`), `:278-294` (`to_voxel`) with
`_splat_points_to_volumes` (`:13-132`), and the writer loop of `ldm/inference.py:174-179` (depth mask < 90 m,
float32 N x 4 `.bin`).  Pinned against the reference's own
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


def to_voxel(range_images, incl, height, mode=MODE_LINEAR, mean=20.0, std=40.0, fill=100.0,
             grid_sizes=(1, 1024, 1024), pc_range=(-25.6, -25.6, -3.0, 25.6, 25.6, 1.0), normalize=True,
             min_weight=1e-4):
    """Bird's-eye-view volume of `to_voxel` (`ldm/dataset.py:278-294`): trilinear splat of every point into the
    (D, H, W) = grid_sizes volume (`_splat_points_to_volumes`, `:13-132`), features = remission divided by the
    clamped vote weight, densities -> log(d + 1).  Returns (B, 2*D, H, W): [densities, features]."""
    pc = to_points(range_images, incl, height, mode, mean, std, fill).double()
    B, N, _ = pc.shape
    D, Hh, Ww = grid_sizes
    lo = torch.tensor(pc_range[:3], dtype=torch.float32).double()
    hi = torch.tensor(pc_range[3:], dtype=torch.float32).double()
    # the reference does this in fp32 (`:282-283`); the oracle keeps fp32 semantics for the index computation
    xyz = ((pc[..., :3].float() - ((hi + lo) / 2).float()) / ((hi - lo) / 2).float())
    feat = pc[..., 3].float()
    gxyz = torch.tensor([Ww, Hh, D], dtype=torch.float32)
    idx3 = ((xyz + 1) * 0.5) * (gxyz - 1)                                     # `:66-68`
    base = idx3.floor()
    r = idx3 - base                                                           # `:70`
    base = base.long()
    n_vox = D * Hh * Ww
    dens = torch.zeros(B, n_vox, dtype=torch.float32)
    feats = torch.zeros(B, n_vox, dtype=torch.float32)
    for xd in (0, 1):                                                         # `:81-123`
        X_ = base[..., 0] + xd
        wX = (1 - xd) + (2 * xd - 1) * r[..., 0]
        for yd in (0, 1):
            Y_ = base[..., 1] + yd
            wY = (1 - yd) + (2 * yd - 1) * r[..., 1]
            for zd in (0, 1):
                Z_ = base[..., 2] + zd
                wZ = (1 - zd) + (2 * zd - 1) * r[..., 2]
                w = wX * wY * wZ
                valid = (0 <= X_) & (X_ < Ww) & (0 <= Y_) & (Y_ < Hh) & (0 <= Z_) & (Z_ < D)
                idx = ((Z_ * Hh + Y_) * Ww + X_) * valid                     # invalid votes: weight 0 into voxel 0
                wv = w * valid
                dens.scatter_add_(1, idx, wv)
                feats.scatter_add_(1, idx, wv * feat)
    feats = feats / dens.clamp(min_weight)                                    # `:126-128`
    if normalize:
        dens = torch.log(dens + 1)                                            # `:288-289`
    return torch.cat([dens.view(B, D, Hh, Ww), feats.view(B, D, Hh, Ww)], dim=1)


def bev_image(voxel_j):
    """uint8 (W, H) preview the reference saves as `<index>.png` (`ldm/inference.py:180-181`)."""
    return (voxel_j.permute(2, 1, 0).cpu().detach().numpy().clip(0, 1) * 255.0).astype(np.uint8)[:, :, 0]
