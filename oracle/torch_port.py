"""Restatement of the reference's ALGORITHM for the hot path, in plain PyTorch.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is what the reference runs:
a stack of ``nn.Linear`` + ``sin(30 x)`` evaluated with ATen ops, and derivatives obtained
by nested ``torch.autograd.grad(create_graph=True)`` sweeps.  It is the ``cpu_baseline``
("port") timed by ``bench.py`` and the fp32 checker the GPU parity tests compare with.

Follows (relative to /root/reference):
  base/networks.py:12-17   get_network
  base/networks.py:21-27   Sine
  base/networks.py:30-71   MLP (only the sine / outermost_linear=True variant the
                           configs ever request, config.py:97-100)
  base/networks.py:80-93   sine_init / first_layer_sine_init
  base/diff_ops.py:6-82    hessian / laplace / divergence / gradient / jacobian
  base/sampling.py:4-64    sample_uniform / sample_random / sample_boundary /
                           sample_boundary2D_separate
"""
from __future__ import annotations

import math

import torch
from torch import nn

OMEGA = 30.0


class SineAct(nn.Module):
    def forward(self, t):
        return torch.sin(OMEGA * t)


class RefMLP(nn.Module):
    """Same module tree as the reference MLP so that state_dict keys are ``net.{0,2,..}.*``."""

    def __init__(self, in_features, out_features, num_hidden_layers, hidden_features):
        super().__init__()
        widths = [in_features] + [hidden_features] * (num_hidden_layers + 1)
        mods = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            mods += [nn.Linear(fan_in, fan_out), SineAct()]
        mods.append(nn.Linear(hidden_features, out_features))
        self.net = nn.Sequential(*mods)
        with torch.no_grad():
            for idx, m in enumerate(self.net):
                if isinstance(m, nn.Linear):
                    fan_in = m.weight.shape[1]
                    bound = 1.0 / fan_in if idx == 0 else math.sqrt(6.0 / fan_in) / OMEGA
                    m.weight.uniform_(-bound, bound)

    def forward(self, coords, weights=None):
        out = self.net(coords)
        return out if weights is None else out * weights

    def flat_theta(self):
        return torch.cat([p.detach().reshape(-1) for p in self.parameters()])

    def load_flat_theta(self, theta):
        theta = torch.as_tensor(theta)
        p0 = 0
        with torch.no_grad():
            for p in self.parameters():
                n = p.numel()
                p.copy_(theta[p0:p0 + n].reshape(p.shape).to(p.dtype))
                p0 += n
        assert p0 == theta.numel()
        return self


def get_network(cfg, in_features, out_features):
    if cfg.network != "siren":
        raise NotImplementedError
    return RefMLP(in_features, out_features, cfg.num_hidden_layers, cfg.hidden_features)


# ---------------------------------------------------------------- diff_ops
def gradient(y, x, grad_outputs=None):
    if grad_outputs is None:
        grad_outputs = torch.ones_like(y)
    return torch.autograd.grad(y, [x], grad_outputs=grad_outputs, create_graph=True)[0]


def divergence(y, x):
    total = 0.0
    for i in range(y.shape[-1]):
        col = y[..., i]
        total = total + torch.autograd.grad(col, x, torch.ones_like(col), create_graph=True)[0][..., i:i + 1]
    return total


def laplace(y, x, normalize=False, eps=0.0, return_grad=False):
    g = gradient(y, x)
    if normalize:
        g = g / (g.norm(dim=-1, keepdim=True) + eps)
    lap = divergence(g, x)
    return (lap, g) if return_grad else lap


def jacobian(y, x):
    rows = []
    for i in range(y.shape[-1]):
        col = y[..., i]
        rows.append(torch.autograd.grad(col, x, torch.ones_like(col), create_graph=True)[0])
    jac = torch.stack(rows, dim=-2)
    status = -1 if bool(torch.isnan(jac).any()) else 0
    return jac, status


def hessian(y, x):
    """y: (M, N, O), x: (M, N, D)  ->  (M, N, O, D, D)"""
    ones = torch.ones_like(y[..., 0])
    per_out = []
    for i in range(y.shape[-1]):
        dydx = torch.autograd.grad(y[..., i], x, ones, create_graph=True)[0]
        per_out.append(torch.stack(
            [torch.autograd.grad(dydx[..., j], x, ones, create_graph=True)[0] for j in range(x.shape[-1])],
            dim=-2))
    h = torch.stack(per_out, dim=-3)
    status = -1 if bool(torch.isnan(h).any()) else 0
    return h, status


# ---------------------------------------------------------------- sampling
def sample_uniform(resolution, sdim=1, device="cpu", flatten=True):
    c = (torch.arange(resolution, device=device, dtype=torch.float32) + 0.5) / resolution * 2 - 1
    grid = torch.stack(torch.meshgrid([c] * sdim, indexing="ij"), dim=-1)
    return grid.reshape(resolution ** sdim, sdim) if flatten else grid


def sample_random(N, sdim=1, device="cpu"):
    return torch.rand(N, sdim, device=device) * 2 - 1


def _band(n, xr, yr, device):
    pts = torch.empty(n, 2, device=device)
    pts[:, 0] = torch.rand(n, device=device) * (xr[1] - xr[0]) + xr[0]
    pts[:, 1] = torch.rand(n, device=device) * (yr[1] - yr[0]) + yr[0]
    return pts


def sample_boundary(N, sdim, epsilon=1e-4, device="cpu"):
    if sdim == 1:
        left = (torch.rand(N // 2, 1, device=device) * 2 - 1) * epsilon - 1.0
        right = (torch.rand(N // 2, 1, device=device) * 2 - 1) * epsilon + 1.0
        return torch.cat([left, right], dim=0)
    if sdim == 2:
        lo, hi = (-1 - epsilon, -1 + epsilon), (1 - epsilon, 1 + epsilon)
        full = (-1, 1)
        bands = [(full, lo), (full, hi), (lo, full), (hi, full)]
        return torch.cat([_band(N // 4, xr, yr, device) for xr, yr in bands], dim=0)
    raise NotImplementedError


def sample_boundary2D_separate(N, side, epsilon=1e-4, device="cpu"):
    lo, hi = (-1 - epsilon, -1 + epsilon), (1 - epsilon, 1 + epsilon)
    full = (-1, 1)
    if side == "horizontal":
        bands = [(lo, full), (hi, full)]
    elif side == "vertical":
        bands = [(full, lo), (full, hi)]
    else:
        raise RuntimeError(side)
    return torch.cat([_band(N // 2, xr, yr, device) for xr, yr in bands], dim=0)
