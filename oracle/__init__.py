"""CPU oracle for the INSR-PDE hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``insr_pde_b200``) imports this package.  The only
allowed callers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker or as the
CPU baseline, never as the thing shipped or measured as "ours".

Contents
--------
``siren_fwdmode``  independent fp64 numpy restatement (forward-mode value / Jacobian /
                   Laplacian / Hessian streams + hand-derived reverse sweep) of
                   ``base/networks.py:21-71`` + ``base/diff_ops.py:6-82``.
``torch_port``     restatement of the reference's *algorithm* (nn.Linear + sin(30x) +
                   ``torch.autograd.grad(create_graph=True)``), i.e. what the reference
                   executes on CPU; this is the ``cpu_baseline`` ("port").
``closures``       restated loss closures of ``advection/model.py:43-91``,
                   ``fluid/model.py:43-151``, ``elasticity/model.py:109-189`` on explicit
                   sample tensors (the reference samples inside the closure).
``ref_loader``     imports the *real* reference from ``/root/reference`` with stub
                   modules (only possible in the build container; never on the GPU box).
``make_goldens``   script that ran the real reference here and wrote ``tests/golden/``.

Parity pinning: the reference has no tests / golden vectors of its own,
so the oracle is pinned against outputs of the reference itself, generated in the build
container by ``make_goldens.py`` and committed under ``tests/golden/``.
"""
