"""The drop-in autograd boundary: one ``torch.autograd.Function`` family in front of the fused
sm_100a kernels, so that the reference's ``model.py`` / ``diff_ops.py`` / ``loss.backward()``
(base/baseModel.py:77) run unchanged on top of them.

Mechanism.  ``SirenFn.apply(x, net, order, *params)`` evaluates the field at
derivative ``order`` with ONE kernel and returns ``(y[, J[, H2]])``.  Its ``backward`` has two
modes:

* ordinary backward (``loss.backward()``; grad mode off): ONE fused reverse kernel takes the
  cotangents of every output and produces the flat parameter gradient (and the gradient
  w.r.t. the points).
* differentiable backward (called from ``torch.autograd.grad(..., create_graph=True)``, which
  is how base/diff_ops.py:47-49,56-57,76 obtain spatial derivatives): the gradient w.r.t.
  the points is expressed through a *higher-order evaluation of the same field* -- J for a
  first derivative, the full Hessian for a second -- obtained from one more kernel launch,
  cached on the node, and itself an autograd output.  The reference's nested autograd sweeps
  therefore collapse to: fwd[y], fwd[y,J] (, fwd[y,J,H]) and one reverse kernel per node
  that actually received a cotangent.

Limitations (none are exercised by the reference): gradients w.r.t. the parameters are not
themselves differentiable, and a backward under ``create_graph=True`` delivers gradients w.r.t.
the POINTS only -- the direct parameter gradient of the node is not produced in that mode (it
raises when the points need no gradient, i.e. when the parameter gradient is all that can be
meant; with points that require grad use ``loss.backward()`` without create_graph, as
base/baseModel.py:77 does); third spatial derivatives through autograd are not available.
"""
from __future__ import annotations

import torch

from . import _ops
from ._ops import ORDER_HESS, ORDER_JAC, ORDER_LAP, ORDER_VALUE


class SirenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, order, *params):
        theta = net.flat_theta()
        # a forward that will be differentiated w.r.t. the parameters keeps its tape (the 32 < H <= 512 family would
        # otherwise recompute every hidden layer in the reverse sweep); frozen nets and no_grad calls do not
        want_tape = any(ctx.needs_input_grad[3:])
        if want_tape:
            outs, ctx.tape = _ops.siren_forward(net.desc, theta, x, order, keep_tape=True)
        else:
            outs, ctx.tape = _ops.siren_forward(net.desc, theta, x, order), None
        ctx.net, ctx.order, ctx.theta = net, order, theta
        ctx.save_for_backward(x, *params)
        ctx.set_materialize_grads(False)
        ctx.higher = None
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        x, *params = ctx.saved_tensors
        net, order = ctx.net, ctx.order
        gouts = list(gouts) + [None] * (3 - len(gouts))
        gy, gjac, gh2 = gouts
        n_extra = 2  # net, order
        if gy is None and gjac is None and gh2 is None:
            return (None,) * (1 + n_extra + len(params))

        if torch.is_grad_enabled():
            # ---- differentiable backward: express dL/dx through higher-order outputs
            if not ctx.needs_input_grad[0]:
                # no gradient w.r.t. the points can be wanted, so this call is after d/d theta with create_graph=True
                raise RuntimeError(
                    "insr_pde_b200: parameter gradients under create_graph=True are not supported by the fused path "
                    "(they would be silently missing); call backward() without create_graph, as base/baseModel.py:77 does")
            if gh2 is not None:
                raise RuntimeError(
                    "insr_pde_b200: a differentiable (create_graph=True) backward through second "
                    "derivatives needs third spatial derivatives, which the fused path does not "
                    "provide (the reference never requests them)")
            want = ORDER_JAC if gjac is None else ORDER_HESS
            if ctx.higher is None or ctx.higher[0] < want:
                ctx.higher = (want, SirenFn.apply(x, net, want, *params))
            outs = ctx.higher[1]
            gx = None
            if gy is not None:
                gx = torch.einsum("nod,no->nd", outs[1], gy)
            if gjac is not None:
                term = torch.einsum("node,nod->ne", outs[2], gjac)
                gx = term if gx is None else gx + term
            return (gx, None, None) + (None,) * len(params)

        # ---- ordinary backward: one fused reverse kernel
        need_gx = ctx.needs_input_grad[0]
        gtheta, gx = _ops.siren_backward(net.desc, ctx.theta, x, order, gy, gjac, gh2, need_gx=need_gx, tape=ctx.tape)
        ctx.tape = None                                   # the reverse kernels overwrite parts of it
        grads = []
        for (off, numel, shape), need in zip(net.param_slices(), ctx.needs_input_grad[1 + n_extra:]):
            grads.append(gtheta[off:off + numel].view(shape) if need else None)
        return (gx, None, None, *grads)


def evaluate(net, x2d, order):
    """(y[, J[, H2]]) of ``net`` at the (N, D) points ``x2d`` with autograd connectivity."""
    return SirenFn.apply(x2d, net, order, *net.parameters())


class FieldSource:
    """Provenance tag attached to the tensors returned by MLP.forward so that the drop-in
    diff_ops can fuse ``laplace(net(x), x)`` & co. into one kernel.  Holds the *original*
    coordinates tensor (any leading shape) and caches evaluations by order."""

    __slots__ = ("net", "coords", "x2d", "cache")

    def __init__(self, net, coords, x2d):
        self.net, self.coords, self.x2d, self.cache = net, coords, x2d, {}

    def outputs(self, order):
        """evaluation at >= ``order`` (cached; a higher cached order is reused)."""
        for have in sorted(self.cache):
            if have >= order and not (order == ORDER_LAP and have == ORDER_HESS):
                return have, self.cache[have]
        outs = evaluate(self.net, self.x2d, order)
        self.cache[order] = outs
        return order, outs
