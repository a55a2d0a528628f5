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


class BoxSampler:
    """Several uniform-in-a-box point sets per call from ONE kernel (insr_sample_boxes): the graph-friendly
    replacement of the torch.rand / scale / shift / cat sequences above when the reference's exact random stream is
    not required (same distributions, Philox4x32-10 keyed by seed, point index and a device iteration counter that
    the kernel bumps itself -- a CUDA-graph replay draws fresh points).

    ``sets``: list of point sets, each a list of boxes ``(count, lo, hi)`` with ``lo`` / ``hi`` of length dim;
    ``sample()`` returns one (sum of counts, dim) tensor per set (views of one buffer, fresh values on every call)."""

    def __init__(self, sets, dim, seed=0, device="cuda", point_offset=0):
        from . import _lib
        self.dim, self.seed, self.point_offset = dim, int(seed), int(point_offset)
        self.counts, self.lo, self.hi, self.sizes = [], [], [], []
        for boxes in sets:
            self.sizes.append(sum(int(b[0]) for b in boxes))
            for count, lo, hi in boxes:
                self.counts.append(int(count)); self.lo.append([float(v) for v in lo]); self.hi.append([float(v) for v in hi])
        self.device = torch.device(device)
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._lib = _lib

    def sample(self, out=None):
        """``out``: a contiguous (sum of counts, dim) fp32 view to draw into (e.g. a slice of a persistent batch buffer)"""
        lib = self._lib.get_lib()
        if out is None:
            out = torch.empty(sum(self.sizes), self.dim, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != (sum(self.sizes), self.dim) or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError("BoxSampler.sample: out must be a contiguous fp32 tensor of shape (sum of counts, dim)")
        with torch.cuda.device(self.device):
            lib.sample_boxes(self.counts, self.lo, self.hi, self.dim, self.seed, self.counter.data_ptr(), self.ticket.data_ptr(),
                             self.point_offset, out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        return list(torch.split(out, self.sizes, dim=0))


def element_measures(V, F):
    """triangle areas (torchgp/area_weighted_distribution.py:40-42 via per_face_normals.py:33-38) or tetrahedron
    volumes (torchgp/per_tet_volumes.py:12-19): the weights of the reference's Categorical over elements"""
    P = V[F.long()]
    if F.shape[1] == 3:
        return torch.linalg.norm(torch.linalg.cross(P[:, 0] - P[:, 1], P[:, 1] - P[:, 2]), dim=1) * 0.5
    if F.shape[1] == 4:
        a, b, c = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0], P[:, 3] - P[:, 0]
        return torch.abs(torch.sum(c * torch.linalg.cross(a, b), dim=-1)) / 6
    raise ValueError("elements must be triangles (n, 3) or tetrahedra (n, 4)")


class MeshSampler:
    """``sample_mesh(V, F, N, distrib)[:, 0:dim]`` (elasticity/sampling.py:4-9, elasticity/model.py:203) from one kernel
    (insr_sample_mesh): the element by binary search in the cumulative area / volume table, barycentric weights as the
    reference draws them (triangles) or as normalised exponentials = Dirichlet(1,1,1,1) (tetrahedra), all on the device
    -- the reference's tetrahedron path calls numpy on the host and copies.  Same distribution, Philox stream."""

    def __init__(self, V, F, dim_out=3, seed=0, point_offset=0):
        from . import _lib
        self.V = V.detach().to(torch.float32).contiguous()
        if self.V.shape[1] < 3:                                   # 2-D meshes carry a zero third coordinate
            self.V = torch.cat([self.V, self.V.new_zeros(self.V.shape[0], 3 - self.V.shape[1])], dim=1).contiguous()
        self.F = F.detach().to(torch.int32).contiguous()
        measure = element_measures(self.V, self.F).double()
        if not bool((measure > 0).all()):
            raise ValueError("mesh has degenerate elements (zero area / volume)")
        self.cdf = (torch.cumsum(measure, 0) / measure.sum()).to(torch.float32).contiguous()
        self.dim_out, self.seed, self.point_offset = int(dim_out), int(seed), int(point_offset)
        self.device = self.V.device
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._lib = _lib

    def sample(self, n, out=None):
        lib = self._lib.get_lib()
        if out is None:
            out = torch.empty(n, self.dim_out, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != (n, self.dim_out) or not out.is_contiguous() or out.dtype != torch.float32:
            raise ValueError("MeshSampler.sample: out must be a contiguous fp32 tensor of shape (n, dim_out)")
        with torch.cuda.device(self.device):
            lib.sample_mesh(self.V.data_ptr(), self.F.data_ptr(), self.cdf.data_ptr(), self.F.shape[0], self.F.shape[1], n,
                            self.dim_out, self.seed, self.counter.data_ptr(), self.ticket.data_ptr(), self.point_offset,
                            out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        return out


def fluid_sets(n_interior, n_boundary, epsilon=1e-4):
    """the three sets of a fluid iteration: sample_random(n, 2), sample_boundary2D_separate(nb, 'horizontal'),
    sample_boundary2D_separate(nb, 'vertical')  (fluid/model.py:75-77 etc., base/sampling.py:45-64)"""
    inner, lo, hi = (-1.0, 1.0), (-1.0 - epsilon, -1.0 + epsilon), (1.0 - epsilon, 1.0 + epsilon)
    box = lambda n, xr, yr: (n, (xr[0], yr[0]), (xr[1], yr[1]))
    return [[box(n_interior, inner, inner)],
            [box(n_boundary // 2, lo, inner), box(n_boundary // 2, hi, inner)],
            [box(n_boundary // 2, inner, lo), box(n_boundary // 2, inner, hi)]]


def shard(points, rank, world_size):
    """contiguous slice [r*N/G, (r+1)*N/G) of a global point set for data-parallel rank r."""
    n = points.shape[0]
    lo = (n * rank) // world_size
    hi = (n * (rank + 1)) // world_size
    return points[lo:hi]
