"""Range image -> point cloud on the GPU: the host-side mirror of the reference's
`point_cloud_to_range_image.to_pc_torch` / `to_voxel` (`ldm/dataset.py:228-294`) and of the `.bin` writer loop of
`ldm/inference.py:174-179`.

The sensor tables (`incl`, `height`, e.g. `ldm/kitti360_range_image.py:19-48`) are DATA supplied by the caller:
`RangeImageGeometry.from_reference(to_range)` copies them -- and `log`, `inverse`, `mean`, `std`,
`range_fill_value` -- from a reference `point_cloud_to_range_image*` object, so `ldm/inference.py:171`
(`pc_all = to_range.to_pc_torch(image)`) becomes `pc_all = geom.to_pc_torch(image)`.
"""
import numpy as np
import torch

from . import _lib

MODE_LINEAR, MODE_LOG, MODE_INVERSE = 0, 1, 2


class RangeImageGeometry:
    def __init__(self, incl, height, log=False, inverse=False, mean=20.0, std=40.0, range_fill_value=(100, 0),
                 grid_sizes=(1, 1024, 1024), pc_range=(-25.6, -25.6, -3.0, 25.6, 25.6, 1.0),
                 normalize_volume_densities=True):
        self.incl = np.ascontiguousarray(np.asarray(incl, dtype=np.float32))
        self.height = np.ascontiguousarray(np.asarray(height, dtype=np.float32))
        if self.incl.shape != self.height.shape or self.incl.ndim != 1:
            raise ValueError("incl and height must be 1-D tables of the same length (one entry per beam)")
        self.log, self.inverse = bool(log), bool(inverse)
        self.mean, self.std = float(mean), float(std)
        self.range_fill_value = np.asarray(range_fill_value)
        self.grid_sizes = [int(g) for g in grid_sizes]                       # (D, H, W) as in `ldm/dataset.py:137`
        self.pc_range = [float(v) for v in pc_range]
        self.normalize_volume_densities = bool(normalize_volume_densities)
        self._dev = {}

    @classmethod
    def from_reference(cls, to_range):
        """Build from a reference `point_cloud_to_range_image*` instance (`ldm/inference.py:81`)."""
        return cls(to_range.incl, to_range.height, log=to_range.log, inverse=to_range.inverse, mean=to_range.mean,
                   std=to_range.std, range_fill_value=to_range.range_fill_value, grid_sizes=to_range.grid_sizes,
                   pc_range=to_range.pc_range, normalize_volume_densities=to_range.normalize_volume_densities)

    @property
    def mode(self):
        return MODE_LOG if self.log else (MODE_INVERSE if self.inverse else MODE_LINEAR)     # `:241-246` precedence

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.incl).to(device), torch.from_numpy(self.height).to(device))
        return self._dev[key]

    def to_pc_torch(self, range_images, return_depth=False):
        """range_images (B, C, W, H) fp32 on CUDA -> (B, W*H, 3|4) [and depth (B, W*H)].  No CPU fallback."""
        if not range_images.is_cuda:
            raise RuntimeError("RangeImageGeometry.to_pc_torch needs a CUDA tensor: rangeldm_b200 has no CPU fallback")
        B, C, W, H = range_images.shape
        if H != self.incl.shape[0]:
            raise AssertionError(f"range image has {H} beams, the sensor tables {self.incl.shape[0]}")
        x = range_images.to(torch.float32).contiguous()
        incl, height = self._tables(x.device)
        pts = torch.empty((B, W * H, 4 if C > 1 else 3), device=x.device)
        depth = torch.empty((B, W * H), device=x.device) if return_depth else None
        _lib.call("rldm_range_to_points", _lib.ptr(x), B, C, W, H, _lib.ptr(incl), _lib.ptr(height), self.mode, self.mean,
                  self.std, float(self.range_fill_value[0]), _lib.ptr(pts), _lib.ptr(depth))
        return (pts, depth) if return_depth else pts

    def from_points(self, pc, width=1024):
        """Point cloud (N, 4) fp32 CUDA tensor [x, y, z, remission] -> the sample `RangeDataset.__getitem__` builds
        (`ldm/dataset.py:327-336`): {'jpg': (2, W, H) range image, 'mask': (W, H) bool, 'car_window_mask': (W, H) bool}
        = projection (`:159-183`, KITTI beam assignment `ldm/kitti360_range_image.py:51-61`) + `process_miss_value`
        (`:193-221`) + `normalize` (`:223-226`)."""
        if not pc.is_cuda:
            raise RuntimeError("RangeImageGeometry.from_points needs a CUDA tensor: rangeldm_b200 has no CPU fallback")
        if pc.ndim != 2 or pc.shape[1] != 4:
            raise ValueError("point cloud must be (N, 4): x, y, z, remission")
        pts = pc.to(torch.float32).contiguous()
        H = self.incl.shape[0]
        incl, height = self._tables(pts.device)
        keys = torch.empty(H * width, dtype=torch.int64, device=pts.device)
        image = torch.empty((2, width, H), device=pts.device)
        mask = torch.empty((width, H), dtype=torch.uint8, device=pts.device)
        car = torch.empty((width, H), dtype=torch.uint8, device=pts.device)
        _lib.call("rldm_points_to_range", _lib.ptr(pts), pts.shape[0], _lib.ptr(incl), _lib.ptr(height), H, width, self.mode,
                  self.mean, self.std, float(self.range_fill_value[0]), float(self.range_fill_value[1]), _lib.ptr(keys),
                  _lib.ptr(image), _lib.ptr(mask), _lib.ptr(car))
        return {"jpg": image, "mask": mask.bool(), "car_window_mask": car.bool()}

    def to_voxel(self, range_images):
        """`to_voxel` (`ldm/dataset.py:278-294`): (B, C, W, H) range images -> (B, 2*D, Hg, Wg) bird's-eye-view volume
        [log-densities, remission features] by trilinear splatting of the point cloud (`ldm/inference.py:172`)."""
        import ctypes
        pts = self.to_pc_torch(range_images)
        B, N, P = pts.shape
        D, Hg, Wg = self.grid_sizes
        scratch = torch.empty((2, B, D * Hg * Wg), device=pts.device)
        voxel = torch.empty((B, 2 * D, Hg, Wg), device=pts.device)
        rng = (ctypes.c_float * 6)(*self.pc_range)
        _lib.call("rldm_points_to_voxel", _lib.ptr(pts), B, N, P, rng, D, Hg, Wg, int(self.normalize_volume_densities),
                  _lib.ptr(scratch), _lib.ptr(voxel))
        return voxel

    @staticmethod
    def bev_image(voxel_j):
        """uint8 (Wg, Hg) preview image the reference saves as `<index>.png` (`ldm/inference.py:180-181`)."""
        return (voxel_j.permute(2, 1, 0).cpu().detach().numpy().clip(0, 1) * 255.0).astype(np.uint8)[:, :, 0]

    def masked_points(self, range_images, max_depth=90.0):
        """Per sample the float32 (N_i, 3|4) arrays the reference writes (`ldm/inference.py:175-179`): rows with
        |xyz| < max_depth, original order.  One device pass + one D2H copy for the whole batch."""
        pts, depth = self.to_pc_torch(range_images, return_depth=True)
        pts_h, keep = pts.cpu().numpy(), (depth < max_depth).cpu().numpy()
        return [pts_h[j][keep[j], :] for j in range(pts_h.shape[0])]

    def write_bins(self, range_images, out_dir, first_index=0, max_depth=90.0, limit=None):
        """`pc[mask, :].tofile(f'{out}/{index}.bin')` for every sample (`ldm/inference.py:174-179`)."""
        import os
        os.makedirs(out_dir, exist_ok=True)
        paths = []
        for j, pc in enumerate(self.masked_points(range_images, max_depth)):
            if limit is not None and first_index + j >= limit:
                break
            path = os.path.join(out_dir, f"{first_index + j}.bin")
            pc.tofile(path)
            paths.append(path)
        return paths
