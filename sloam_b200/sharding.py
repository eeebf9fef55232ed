"""Keyframe sharding across ranks.

Keyframes are independent once firstScan_/prevGPlanes_/submap are explicit inputs, so rank r
of W takes the contiguous block [r*K/W, (r+1)*K/W); the only exchange is the gather of the
fixed-size per-keyframe result records (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_keyframes, rank, world):
    """Contiguous block of keyframes owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_keyframes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(local, n_keyframes, rank, world, dist, device="cpu"):
    """All-gather ragged per-rank record arrays (numpy structured, one per keyframe) into the
    global keyframe order.  `dist` is torch.distributed (already initialised)."""
    import torch
    cap = max(shard_range(n_keyframes, r, world)[1] - shard_range(n_keyframes, r, world)[0]
              for r in range(world))
    item = local.dtype.itemsize
    buf = np.zeros(cap * item, np.uint8)
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    buf[:raw.size] = raw
    send = torch.from_numpy(buf).to(device)
    recv = torch.empty(world * cap * item, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy()
    out = np.zeros(n_keyframes, local.dtype)
    for r in range(world):
        lo, hi = shard_range(n_keyframes, r, world)
        out[lo:hi] = recv[r * cap * item:(r * cap + hi - lo) * item].view(local.dtype)
    return out
