"""Independent fp64 restatement of the SIREN field and its spatial derivatives.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

What it restates (citations relative to /root/reference):
  * ``base/networks.py:21-27``  Sine: ``sin(30 * x)``
  * ``base/networks.py:30-71``  MLP: Linear(D,H)+sine, L x [Linear(H,H)+sine], Linear(H,O)
  * ``base/diff_ops.py:53-58``  gradient  -> J^T . grad_outputs
  * ``base/diff_ops.py:44-50``  divergence -> trace of J
  * ``base/diff_ops.py:33-41``  laplace   -> sum_d d2y/dx_d^2
  * ``base/diff_ops.py:61-82``  jacobian  -> J[n, o, d]
  * ``base/diff_ops.py:6-30``   hessian   -> H[n, o, d, e]

The reference obtains the derivatives by nested reverse sweeps of torch.autograd.  This
file instead propagates *forward-mode* streams through every layer (value, D tangents,
second-order streams), and the parameter gradient by a hand-derived reverse sweep over
that forward-mode program.  It shares no code path with torch.autograd, which is what
makes it useful as an arbiter between the fp32 reference and the fp32 CUDA kernels.

Stream recurrences (omega = 30, z = W a + b):
    a0 = sin(omega z)
    a_d = omega cos(omega z) * zdot_d
    a_q = omega cos(omega z) * zddot_q - omega^2 sin(omega z) * Q_q(zdot)
with Q_q = sum_d zdot_d^2 for the Laplacian-trace stream and zdot_d * zdot_e for the
Hessian stream (d, e).  Layer 1 has zdot_d = W1[:, d], zddot = 0.

theta layout (flat, nn.Module.parameters() order): W1 (H,D), b1 (H), W2 (H,H), b2 (H),
..., Wout (O,H), bout (O).
"""
from __future__ import annotations

import numpy as np

OMEGA = 30.0

ORDER_VALUE = 0      # y
ORDER_JAC = 1        # y, J
ORDER_LAP = 2        # y, J, per-output Laplacian (trace stream)
ORDER_HESS = 3       # y, J, full Hessian


def layer_shapes(D, O, H, L):
    """[(out, in)] for the L+2 Linear layers of MLP (base/networks.py:50-60)."""
    shapes = [(H, D)]
    shapes += [(H, H)] * L
    shapes += [(O, H)]
    return shapes


def theta_size(D, O, H, L):
    return sum(o * i + o for o, i in layer_shapes(D, O, H, L))


def unpack_theta(theta, D, O, H, L):
    """flat theta -> [(W, b)] views."""
    out, p = [], 0
    for (o, i) in layer_shapes(D, O, H, L):
        W = theta[p:p + o * i].reshape(o, i); p += o * i
        b = theta[p:p + o]; p += o
        out.append((W, b))
    assert p == theta.size
    return out


def second_order_pairs(D, order):
    """list of second-order streams; each is a list of (d, e, weight) products."""
    if order == ORDER_LAP:
        return [[(d, d, 1.0) for d in range(D)]]
    if order == ORDER_HESS:
        return [[(d, e, 1.0)] for d in range(D) for e in range(d, D)]
    return []


def forward(theta, x, D, O, H, L, order, dtype=np.float64, keep=False):
    """Returns dict(y (N,O), jac (N,O,D), lap (N,O) | hess (N,O,D,D)); keep=True also
    returns the per-layer tape used by ``backward``."""
    theta = np.asarray(theta, dtype=dtype)
    x = np.asarray(x, dtype=dtype).reshape(-1, D)
    N = x.shape[0]
    layers = unpack_theta(theta, D, O, H, L)
    Q = second_order_pairs(D, order)
    nd = D if order >= ORDER_JAC else 0
    w = dtype(OMEGA)

    tape = []
    # ---- first sine layer: tangents are the columns of W1, second order is zero
    W1, b1 = layers[0]
    z = w * (x @ W1.T + b1)                              # (N,H) pre-activation (omega folded)
    zd = [np.broadcast_to(w * W1[:, d], (N, H)) for d in range(nd)]
    zq = [np.zeros((N, H), dtype=dtype) for _ in Q]
    a_in = None
    for li in range(L + 1):
        if li > 0:
            W, b = layers[li]
            z = w * (a[0] @ W.T + b)
            zd = [w * (a[1 + d] @ W.T) for d in range(nd)]
            zq = [w * (a[1 + nd + q] @ W.T) for q in range(len(Q))]
        s, c = np.sin(z), np.cos(z)
        a_new = [s] + [c * zd[d] for d in range(nd)]
        for q, prods in enumerate(Q):
            quad = sum(wt * zd[d] * zd[e] for (d, e, wt) in prods)
            a_new.append(c * zq[q] - s * quad)
        if keep:
            tape.append(dict(a_in=a_in, s=s, c=c, zd=zd, zq=zq))
        a_in = a_new
        a = a_new
    Wo, bo = layers[-1]
    out = {"y": a[0] @ Wo.T + bo}
    if nd:
        out["jac"] = np.stack([a[1 + d] @ Wo.T for d in range(D)], axis=-1)     # (N,O,D)
    if order == ORDER_LAP:
        out["lap"] = a[1 + nd] @ Wo.T                                            # (N,O)
    if order == ORDER_HESS:
        hess = np.zeros((N, O, D, D), dtype=dtype)
        for q, prods in enumerate(Q):
            d, e, _ = prods[0]
            v = a[1 + nd + q] @ Wo.T
            hess[:, :, d, e] = v
            hess[:, :, e, d] = v
        out["hess"] = hess
    if keep:
        out["_tape"] = tape
        out["_a_last"] = a
    return out


def backward(theta, x, D, O, H, L, order, gy=None, gjac=None, glap=None, ghess=None,
             dtype=np.float64):
    """Reverse sweep over the forward-mode program.

    Returns (gtheta flat, gx (N,D)) for the scalar  sum(gy*y) + sum(gjac*jac) + sum(glap*lap)
    + sum(ghess*hess).  This is what ``loss.backward()`` (base/baseModel.py:77) produces
    for the parameters when the loss consumed y / J / lap with those cotangents."""
    theta = np.asarray(theta, dtype=dtype)
    x = np.asarray(x, dtype=dtype).reshape(-1, D)
    N = x.shape[0]
    fw = forward(theta, x, D, O, H, L, order, dtype=dtype, keep=True)
    tape, a_last = fw["_tape"], fw["_a_last"]
    layers = unpack_theta(theta, D, O, H, L)
    Q = second_order_pairs(D, order)
    nd = D if order >= ORDER_JAC else 0
    S = 1 + nd + len(Q)
    w = dtype(OMEGA)

    # cotangents of the output-layer streams: g[s] has shape (N,O)
    g = [np.zeros((N, O), dtype=dtype) for _ in range(S)]
    if gy is not None:
        g[0] = g[0] + np.asarray(gy, dtype=dtype).reshape(N, O)
    if gjac is not None and nd:
        gj = np.asarray(gjac, dtype=dtype).reshape(N, O, D)
        for d in range(D):
            g[1 + d] = g[1 + d] + gj[:, :, d]
    if glap is not None and order == ORDER_LAP:
        g[1 + nd] = g[1 + nd] + np.asarray(glap, dtype=dtype).reshape(N, O)
    if ghess is not None and order == ORDER_HESS:
        gh = np.asarray(ghess, dtype=dtype).reshape(N, O, D, D)
        for q, prods in enumerate(Q):
            d, e, _ = prods[0]
            g[1 + nd + q] = g[1 + nd + q] + (gh[:, :, d, e] if d == e else gh[:, :, d, e] + gh[:, :, e, d])

    grads = [None] * (L + 2)
    Wo, bo = layers[-1]
    gWo = sum(g[s].T @ a_last[s] for s in range(S))      # (O,H)
    gbo = g[0].sum(axis=0)
    grads[L + 1] = (gWo, gbo)
    abar = [g[s] @ Wo for s in range(S)]                  # (N,H) per stream

    gx = np.zeros((N, D), dtype=dtype)
    for li in range(L, -1, -1):
        t = tape[li]
        s_, c_, zd, zq = t["s"], t["c"], t["zd"], t["zq"]
        # elementwise adjoint of the sine layer
        zqbar = [c_ * abar[1 + nd + q] for q in range(len(Q))]
        zdbar = [c_ * abar[1 + d] for d in range(nd)]
        zbar = c_ * abar[0]
        for d in range(nd):
            zbar = zbar - s_ * zd[d] * abar[1 + d]
        for q, prods in enumerate(Q):
            aq = abar[1 + nd + q]
            quad = sum(wt * zd[d] * zd[e] for (d, e, wt) in prods)
            zbar = zbar + aq * (-s_ * zq[q] - c_ * quad)
            for (d, e, wt) in prods:
                if d == e:
                    zdbar[d] = zdbar[d] - s_ * aq * (2.0 * wt) * zd[d]
                else:
                    zdbar[d] = zdbar[d] - s_ * aq * wt * zd[e]
                    zdbar[e] = zdbar[e] - s_ * aq * wt * zd[d]
        W, b = layers[li]
        if li == 0:
            gW = w * (zbar.T @ x)                                          # (H,D)
            for d in range(nd):
                gW[:, d] += w * zdbar[d].sum(axis=0)
            gb = w * zbar.sum(axis=0)
            gx = w * (zbar @ W)
            grads[0] = (gW, gb)
        else:
            a_in = t["a_in"]
            zb_all = [zbar] + zdbar + zqbar
            gW = w * sum(zb_all[s].T @ a_in[s] for s in range(S))
            gb = w * zbar.sum(axis=0)
            grads[li] = (gW, gb)
            abar = [w * (zb_all[s] @ W) for s in range(S)]
    gtheta = np.concatenate([np.concatenate([gW.ravel(), gb.ravel()]) for gW, gb in grads])
    return gtheta, gx


# ----------------------------------------------------------------------------------
# diff_ops restated on the explicit outputs (shapes as the reference returns them)
# ----------------------------------------------------------------------------------
def gradient_from_jac(jac, grad_outputs=None):
    """base/diff_ops.py:53-58 -> (N,D): sum_o grad_outputs[n,o] * J[n,o,d]."""
    if grad_outputs is None:
        return jac.sum(axis=-2)
    return np.einsum("...od,...o->...d", jac, grad_outputs)


def divergence_from_jac(jac):
    """base/diff_ops.py:44-50 -> (N,1): sum_i dy_i/dx_i  (needs O == D)."""
    return np.trace(jac, axis1=-2, axis2=-1)[..., None]


def laplace_from_lap(lap):
    """base/diff_ops.py:33-41 -> (N,1): div(grad(y)) with grad = sum_o dy_o/dx."""
    return lap.sum(axis=-1, keepdims=True)
