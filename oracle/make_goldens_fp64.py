"""fp64 companions of the elasticity closure goldens: the REAL reference's ``_solve_deformation`` closure
(elasticity/model.py:127-189, elasticity/losses.py) re-run in double precision on exactly the weights and the sample stream
recorded in tests/golden/closure_elasticity_<tag>.npz, writing tests/golden/closure_elasticity_<tag>_fp64.npz (loss and
parameter gradient).

TEST INFRASTRUCTURE ONLY.  Why: the reference differentiates through torch.svd, whose backward divides by
(sigma_i^2 - sigma_j^2); at near-coincident singular values its fp32 gradient carries ~1e-3 relative error, so the fp32
goldens cannot pin a gradient to 1e-4.  In fp64 the same formula is accurate to ~1e-10, which makes it the arbiter.

    python oracle/make_goldens_fp64.py           (build container: needs /root/reference)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import make_goldens as mg, ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class Replayer:
    """feeds the recorded outputs of the sampling functions back, in order, as float64"""

    def __init__(self, module, names, log):
        self.module, self.names, self.log, self.pos, self.orig = module, names, log, 0, {}

    def __enter__(self):
        for n in self.names:
            if hasattr(self.module, n):
                self.orig[n] = getattr(self.module, n)

                def wrap(*a, __n=n, **k):
                    name, arr = self.log[self.pos]
                    assert name == __n, (name, __n)
                    self.pos += 1
                    return torch.from_numpy(np.asarray(arr, dtype=np.float64)).clone()

                setattr(self.module, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, fn in self.orig.items():
            setattr(self.module, n, fn)


def load_flat(net, theta):
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            n = p.numel()
            p.copy_(torch.from_numpy(theta[off:off + n]).reshape(p.shape).to(p.dtype))
            off += n


def main():
    ref = ref_loader.load(cpu=True)
    torch.set_num_threads(1)
    import elasticity.model as mod
    for tag, over in mg.ELASTICITY_CASES:
        g = dict(np.load(os.path.join(OUT, f"closure_elasticity_{tag}.npz")))
        torch.set_default_dtype(torch.float64)          # base/diff_ops.py:70 allocates the Jacobian in the default dtype
        try:
            torch.manual_seed(13)
            cfg = ref_loader.make_cfg("elasticity", **over)
            m = ref.elasticity.ElasticityModel(cfg)
            for net, key in ((m.deformation_field, "theta.deformation"), (m.deformation_field_prev, "theta.prev"),
                             (m.deformation_field_prev_prev, "theta.prev_prev")):
                net.double()
                load_flat(net, g[key].astype(np.float64))
            for name in ("external_force", "constraint_offset_right", "circle_center"):
                setattr(m, name, getattr(m, name).double())
            m.timestep = 1
            keys = sorted(k for k in g if k.startswith("solve_deformation.samples"))
            log = [(k.split(".")[-1], g[k]) for k in sorted(keys, key=lambda s: int(s.split(".samples")[1].split(".")[0]))]
            fn = mg.raw_closure(type(m)._solve_deformation)
            m.deformation_field.zero_grad()
            with Replayer(mod, mg.SAMPLERS, log) as r:
                loss = fn(m)
            assert r.pos == len(log)
            total = sum(loss.values())
            total.backward()
            grad = torch.cat([p.grad.reshape(-1) for p in m.deformation_field.parameters()]).numpy()
        finally:
            torch.set_default_dtype(torch.float32)
        g32 = g["solve_deformation.grad.deformation"]
        err32 = np.abs(g32 - grad).max() / np.abs(grad).max()
        print(f"{tag}: loss fp64 {float(total):.10g} (fp32 golden {float(g['solve_deformation.loss.main']):.10g}); the reference's own "
              f"fp32 gradient is {err32:.2e} (max-abs / max-abs) from its fp64 gradient")
        np.savez_compressed(os.path.join(OUT, f"closure_elasticity_{tag}_fp64.npz"),
                            **{"solve_deformation.loss64.main": np.array(float(total)),
                               "solve_deformation.grad64.deformation": grad, "reference_fp32_grad_error": np.array(err32)})


if __name__ == "__main__":
    main()
