"""Collocation-point sampling with the signatures of base/sampling.py:4-64.

These stay PyTorch on purpose: they define the input distributions and the
RNG stream that trajectory parity with the reference depends on; their cost is negligible
next to the field evaluation.  ``shard`` is the only addition: the contiguous slice of a
global sample set owned by one data-parallel rank.
"""
from __future__ import annotations

import torch


def sample_uniform(resolution, sdim=1, device="cpu", flatten=True):
    """cell-centred grid in [-1, 1]^sdim (base/sampling.py:4-11)"""
    axis = torch.linspace(0.5, resolution - 0.5, resolution, device=device) / resolution * 2 - 1
    coords = torch.stack(torch.meshgrid([axis] * sdim, indexing="ij"), dim=-1)
    return coords.reshape(resolution ** sdim, sdim) if flatten else coords


def sample_random(N, sdim=1, device="cpu"):
    """U[-1, 1]^sdim (base/sampling.py:14-18)"""
    return torch.rand(N, sdim, device=device) * 2 - 1


def _strip(n, x_range, y_range, device):
    pts = torch.empty(n, 2, device=device)
    pts[:, 0] = torch.rand(n, device=device) * (x_range[1] - x_range[0]) + x_range[0]
    pts[:, 1] = torch.rand(n, device=device) * (y_range[1] - y_range[0]) + y_range[0]
    return pts


def sample_boundary(N, sdim, epsilon=1e-4, device="cpu"):
    """random points in an epsilon band around the boundary (base/sampling.py:21-42)"""
    if sdim == 1:
        left = (torch.rand(N // 2, 1, device=device) * 2 - 1) * epsilon - 1.
        right = (torch.rand(N // 2, 1, device=device) * 2 - 1) * epsilon + 1.
        return torch.cat([left, right], dim=0)
    if sdim == 2:
        inner, lo, hi = [-1, 1], [-1 - epsilon, -1 + epsilon], [1 - epsilon, 1 + epsilon]
        strips = [(inner, lo), (inner, hi), (lo, inner), (hi, inner)]
        return torch.cat([_strip(N // 4, xr, yr, device) for xr, yr in strips], dim=0)
    raise NotImplementedError


def sample_boundary2D_separate(N, side, epsilon=1e-4, device="cpu"):
    """left/right ('horizontal') or bottom/top ('vertical') bands (base/sampling.py:45-64)"""
    inner, lo, hi = [-1, 1], [-1 - epsilon, -1 + epsilon], [1 - epsilon, 1 + epsilon]
    if side == "horizontal":
        strips = [(lo, inner), (hi, inner)]
    elif side == "vertical":
        strips = [(inner, lo), (inner, hi)]
    else:
        raise RuntimeError
    return torch.cat([_strip(N // 2, xr, yr, device) for xr, yr in strips], dim=0)


def shard(points, rank, world_size):
    """contiguous slice [r*N/G, (r+1)*N/G) of a global point set for data-parallel rank r."""
    n = points.shape[0]
    lo = (n * rank) // world_size
    hi = (n * (rank + 1)) // world_size
    return points[lo:hi]
