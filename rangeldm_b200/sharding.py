"""Batch-axis sharding across the GPUs of one box -- the reference's own scheme
(`ldm/inference.py:159,174-183`; seed per global index `ldm/inference_conditional.py:161`):
rank r of P runs iterations i = 0,1,... and owns global sample indices (r + P*i)*B + j.
Every image is independent (GroupNorm and attention are per sample), so there is NO collective on
the hot path; `gather_images` is the optional single all-gather of finished range images."""
import torch


def shard_indices(samples, batch, rank, world):
    out, i = [], 0
    while True:
        base = (rank + world * i) * batch
        if base >= samples:
            return out
        out.extend(k for k in range(base, min(base + batch, samples)))
        i += 1


def gather_images(images, indices, samples):
    """All-gather per-rank images (n_r, C, W, H) with their global indices; rank 0 returns the
    (samples, C, W, H) tensor ordered by global index, other ranks return None.  NCCL over NVLink
    on GPUs (one ncclAllGather on padded buffers); gloo in the CPU tests."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = max(1, -(-samples // world) + 8)        # per-rank capacity (same on every rank)
    idx = torch.full((cap,), -1, dtype=torch.int64, device=images.device)
    idx[:len(indices)] = torch.tensor(indices, dtype=torch.int64, device=images.device)
    buf = torch.zeros((cap,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
    buf[:len(indices)] = images
    all_idx = torch.empty((world * cap,), dtype=torch.int64, device=images.device)
    all_buf = torch.empty((world * cap,) + tuple(buf.shape[1:]), dtype=images.dtype, device=images.device)
    dist.all_gather_into_tensor(all_idx, idx)
    dist.all_gather_into_tensor(all_buf, buf)
    if rank != 0:
        return None
    out = torch.zeros((samples,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
    keep = all_idx >= 0
    out[all_idx[keep]] = all_buf[keep]
    return out
