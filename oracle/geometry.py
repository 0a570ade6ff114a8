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


def kitti_row_inds(pc, incl, height):
    """`point_cloud_to_range_image_KITTI.get_row_inds` (`ldm/kitti360_range_image.py:51-61`): beam whose inclination is
    closest to the elevation of the point seen from that beam's height."""
    xy_norm = np.linalg.norm(pc[:, :2], ord=2, axis=1)
    err = [np.abs(incl[i] - np.arctan2(height[i] - pc[:, 2], xy_norm)) for i in range(len(incl))]
    return np.argmin(np.stack(err, axis=-1), axis=-1)


def points_to_range_image(pc, incl, height, width=1024, mode=MODE_LINEAR, mean=20.0, std=40.0, fill=(100.0, 0.0)):
    """Point cloud (N,4) float32 -> the dataset sample of `RangeDataset.__getitem__` (`ldm/dataset.py:327-336`):
    projection `__call__` (`:159-183`, nearest return wins), `process_miss_value` (`:193-221`), `normalize`
    (`:223-226`).  Returns (image (2, W, H), mask (W, H), car_window_mask (W, H))."""
    pc = np.array(pc, dtype=np.float32, copy=True)
    incl = np.asarray(incl, dtype=np.float32)
    height = np.asarray(height, dtype=np.float32)
    fill = np.asarray(fill)
    H = len(incl)
    row = kitti_row_inds(pc, incl, height)                                                  # `:160`
    azi = np.arctan2(pc[:, 1], pc[:, 0])
    col = width - 1.0 + 0.5 - (azi + np.pi) / (2.0 * np.pi) * width                         # `:163`
    col = np.round(col).astype(np.int32)
    col[col == width] = width - 1
    col[col < 0] = 0
    img = np.full((H, width, 2), -1, dtype=np.float32)
    pc[:, 2] -= height[row]                                                                 # `:168`
    rng = np.linalg.norm(pc[:, :3], axis=1, ord=2)
    rng[rng > fill[0]] = fill[0]
    order = np.argsort(-rng, kind="stable")                                                 # `:172` (ties: see tests)
    if mode == MODE_LOG:
        val = np.log2(rng[order] + 1) / 6
    elif mode == MODE_INVERSE:
        val = 1 / rng[order]
    else:
        val = rng[order]
    img[row[order], col[order], :] = np.concatenate([val[:, None], pc[order][:, 3:4]], axis=1)   # `:183` last wins
    # ---- process_miss_value (`:193-221`)
    mask = img[..., 0] > 0
    miss = img[:, :, 0] == -1
    shift = list(range(1, width)) + [0]
    img1 = img.copy()
    img1[miss, :] = img[:, shift, :][miss, :]
    mask1 = mask.copy()
    mask1[miss] = mask[:, shift][miss]
    still = img1[:, :, 0] == -1
    r0 = img1[:, :, 0]
    down = r0[[H - 2, H - 1] + list(range(H - 2)), :]
    top = r0[list(range(2, H)) + [0, 1], :]
    right = r0[:, [width - 2, width - 1] + list(range(width - 2))]
    left = r0[:, list(range(2, width)) + [0, 1]]
    car = still & ((down != -1) | (top != -1) | (right != -1) | (left != -1))
    if mode == MODE_LOG:
        img1[still, :] = np.log2(fill + 1) / 6
    elif mode == MODE_INVERSE:
        img1[still, :] = np.array([1 / fill[0], fill[1]])
    else:
        img1[still, :] = fill
    if mode == MODE_LINEAR:                                                                 # `:223-226`
        img1[..., 0] = (img1[..., 0] - mean) / std
    return (torch.from_numpy(img1).permute(2, 1, 0).contiguous(), torch.from_numpy(mask1).permute(1, 0).contiguous(),
            torch.from_numpy(car).permute(1, 0).contiguous())
