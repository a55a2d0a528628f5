"""Small-matrix linear algebra of the elasticity closure on the device.

``svd``           drop-in for ``torch.svd`` on (..., 2, 2) / (..., 3, 3) fp32 CUDA matrices
                  (elasticity/model.py:144: ``U_x, S_x, V_x = torch.svd(jac_x)``): one kernel, one thread per matrix
                  (one-sided Jacobi), differentiable through S (d sigma_k / dF = u_k v_k^T); U and V are returned
                  detached -- the reference only ever uses S_x.
``elastic_energy``  the ARAP + volume energies of elasticity/model.py:146-147 and their adjoint in ONE kernel:
                  ``E = r_a sum (S - 1)^2 + r_v sum (prod(S) - 1)^2``; no (N, D) singular-value round trip, no SVD
                  backward graph.
"""
from __future__ import annotations

import torch

from . import _ops


class _SvdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, F):
        U, S, V = _ops.svd_small(F.detach().contiguous())
        ctx.save_for_backward(U, V)
        ctx.mark_non_differentiable(U, V)
        return U, S, V

    @staticmethod
    def backward(ctx, gU, gS, gV):
        U, V = ctx.saved_tensors
        return torch.einsum("...ik,...k,...jk->...ij", U, gS, V)


def svd(A, some=True, compute_uv=True):
    """``torch.svd`` semantics (A = U diag(S) V^T, S descending) for batches of 2x2 / 3x3 matrices"""
    U, S, V = _SvdFn.apply(A)
    return U, S, V


def supports(A):
    return (torch.is_tensor(A) and A.is_cuda and A.dtype == torch.float32 and A.dim() >= 2
            and A.shape[-1] == A.shape[-2] and A.shape[-1] in (2, 3))


class _EnergyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, F, ratio_arap, ratio_volume):
        E, gF = _ops.elastic_energy(F.detach().contiguous(), ratio_arap, ratio_volume, need_grad=True)
        ctx.save_for_backward(gF)
        return E[0]

    @staticmethod
    def backward(ctx, gE):
        (gF,) = ctx.saved_tensors
        return gE * gF, None, None


def elastic_energy(F, ratio_arap, ratio_volume):
    """ratio_arap * sum((S - 1)^2) + ratio_volume * sum((prod(S, -1) - 1)^2) with S = singular values of F (..., d, d)"""
    return _EnergyFn.apply(F, float(ratio_arap), float(ratio_volume))
