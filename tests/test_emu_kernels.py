"""Kernel sources executed under the host SIMT emulator (tests/emu) and compared with the
fp64 oracle: catches indexing / layout / synchronisation mistakes in the CUDA code without a
GPU.  The real parity tests run on the B200 (test_gpu_parity.py)."""
import ctypes

import numpy as np
import pytest

from insr_pde_b200 import _lib
from oracle import siren_fwdmode as fm


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def make_theta(rng, D, O, H, L):
    parts = []
    for li, (o, i) in enumerate(fm.layer_shapes(D, O, H, L)):
        bound = 1.0 / i if li == 0 else np.sqrt(6.0 / i) / 30.0
        parts.append(rng.uniform(-bound, bound, o * i))
        parts.append(rng.uniform(-1, 1, o) / np.sqrt(i))
    return np.concatenate(parts).astype(np.float32)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run_case(lib, D, O, H, L, N, order, flags=0, seed=0):
    rng = np.random.default_rng(seed)
    desc = _lib.make_desc(D, O, H, L, flags=flags)
    theta = make_theta(rng, D, O, H, L)
    assert lib.theta_size(desc) == theta.size
    x = rng.uniform(-1, 1, (N, D)).astype(np.float32)
    y = np.full((N, O), np.nan, np.float32)
    jac = np.full((N, O, D), np.nan, np.float32)
    h2 = np.full((N, O) if order == 2 else (N, O, D, D), np.nan, np.float32)
    nb = lib.workspace_bytes(desc, N, order, False)
    ws = np.zeros(nb // 4 + 8, np.float32)
    lib.forward(desc, ptr(theta), ptr(x), N, order, ptr(y), ptr(jac) if order >= 1 else None,
                ptr(h2) if order >= 2 else None, ptr(ws), nb, None)
    ref = fm.forward(theta.astype(np.float64), x.astype(np.float64), D, O, H, L, order)
    errs = {"y": rel(y, ref["y"])}
    if order >= 1:
        errs["jac"] = rel(jac, ref["jac"])
    if order == 2:
        errs["lap"] = rel(h2, ref["lap"])
    if order == 3:
        errs["hess"] = rel(h2, ref["hess"])
    gy = rng.standard_normal((N, O)).astype(np.float32)
    gj = rng.standard_normal((N, O, D)).astype(np.float32)
    gh = rng.standard_normal(h2.shape).astype(np.float32)
    gth = np.zeros(theta.size, np.float32)
    gx = np.full((N, D), np.nan, np.float32)
    nb = lib.workspace_bytes(desc, N, order, True)
    ws = np.zeros(nb // 4 + 8, np.float32)
    lib.backward(desc, ptr(theta), ptr(x), N, order, ptr(gy), ptr(gj) if order >= 1 else None,
                 ptr(gh) if order >= 2 else None, ptr(gth), ptr(gx), ptr(ws), nb, None)
    kw = dict(gy=gy)
    if order >= 1:
        kw["gjac"] = gj
    if order == 2:
        kw["glap"] = gh
    if order == 3:
        kw["ghess"] = gh
    gref, gxref = fm.backward(theta.astype(np.float64), x.astype(np.float64), D, O, H, L, order, **kw)
    errs["gtheta"] = rel(gth, gref)
    errs["gx"] = rel(gx, gxref)
    assert max(errs.values()) < 1e-4, errs
    return errs


GENERIC_CASES = [
    (2, 1, 32, 3, 150, 2), (1, 1, 20, 2, 70, 1), (2, 2, 32, 3, 130, 0), (3, 3, 66, 3, 40, 1),
    (2, 2, 68, 3, 40, 3), (3, 1, 24, 5, 33, 3), (2, 1, 8, 0, 50, 2), (2, 1, 130, 1, 40, 2),
]


@pytest.mark.parametrize("case", GENERIC_CASES)
def test_generic_family_under_emulation(emu_library, case):
    run_case(emu_library, *case, flags=_lib.FLAG_FORCE_GENERIC)


def test_backward_accumulates_and_null_cotangents(emu_library):
    lib = emu_library
    rng = np.random.default_rng(3)
    D, O, H, L, N = 2, 2, 16, 1, 40
    desc = _lib.make_desc(D, O, H, L, flags=_lib.FLAG_FORCE_GENERIC)
    theta = make_theta(rng, D, O, H, L)
    x = rng.uniform(-1, 1, (N, D)).astype(np.float32)
    gj = rng.standard_normal((N, O, D)).astype(np.float32)
    nb = lib.workspace_bytes(desc, N, 1, True)
    ws = np.zeros(nb // 4 + 8, np.float32)
    g1 = np.zeros(theta.size, np.float32)
    lib.backward(desc, ptr(theta), ptr(x), N, 1, None, ptr(gj), None, ptr(g1), None, ptr(ws), nb, None)
    g2 = g1.copy()
    lib.backward(desc, ptr(theta), ptr(x), N, 1, None, ptr(gj), None, ptr(g2), None, ptr(ws), nb, None)
    assert rel(g2, 2 * g1) < 1e-5
    gref, _ = fm.backward(theta.astype(np.float64), x.astype(np.float64), D, O, H, L, 1, gjac=gj)
    assert rel(g1, gref) < 1e-4


FUSED_CASES = [
    (2, 1, 32, 3, 70, 2), (2, 1, 32, 3, 8, 0), (1, 1, 20, 2, 100, 1), (2, 2, 32, 3, 45, 1),
    (2, 2, 20, 1, 30, 2), (1, 1, 7, 3, 19, 2),
]


@pytest.mark.parametrize("case", FUSED_CASES)
def test_fused_family_under_emulation(emu_library, case):
    D, O, H, L, N, order = case
    desc = _lib.make_desc(D, O, H, L)
    assert emu_library.kernel_family(desc, order, True) == 1 and emu_library.kernel_family(desc, order, False) == 1
    run_case(emu_library, *case)


TILED_CASES = [
    (2, 2, 68, 3, 150, 1), (2, 1, 64, 2, 300, 2), (3, 3, 66, 2, 100, 1), (2, 1, 40, 1, 77, 0),
    (2, 1, 136, 1, 60, 2), (3, 1, 72, 2, 50, 3), (1, 1, 33, 1, 40, 2),
]


@pytest.mark.parametrize("case", TILED_CASES)
def test_tiled_family_under_emulation(emu_library, case):
    D, O, H, L, N, order = case
    desc = _lib.make_desc(D, O, H, L)
    assert emu_library.kernel_family(desc, order, True) == 2 and emu_library.kernel_family(desc, order, False) == 2
    run_case(emu_library, *case)


@pytest.mark.parametrize("case", [(2, 2, 68, 3, 70, 1), (3, 1, 40, 2, 70, 2), (2, 1, 130, 1, 33, 0)])
def test_tiled_forward_keeps_tape_for_backward(emu_library, case):
    """INSR_FLAG_KEEP_TAPE: the forward leaves the activations of every layer in the backward-sized workspace and the
    backward on the same workspace skips the recomputation -- same outputs and the same gradients as the plain pair;
    refused (tape_supported == 0) for the H <= 32 family"""
    lib = emu_library
    D, O, H, L, N, order = case
    rng = np.random.default_rng(11)
    theta = make_theta(rng, D, O, H, L)
    x = rng.uniform(-1, 1, (N, D)).astype(np.float32)
    gy = rng.standard_normal((N, O)).astype(np.float32)
    gj = rng.standard_normal((N, O, D)).astype(np.float32)
    gh = rng.standard_normal((N, O)).astype(np.float32)
    res = {}
    for keep in (False, True):
        desc = _lib.make_desc(D, O, H, L, flags=_lib.FLAG_KEEP_TAPE if keep else 0)
        assert lib.tape_supported(desc, N, order)
        y = np.full((N, O), np.nan, np.float32); jac = np.full((N, O, D), np.nan, np.float32); h2 = np.full((N, O), np.nan, np.float32)
        nb_b = lib.workspace_bytes(desc, N, order, True)
        ws = np.zeros(nb_b // 4 + 8, np.float32)
        nb_f = nb_b if keep else lib.workspace_bytes(desc, N, order, False)
        lib.forward(desc, ptr(theta), ptr(x), N, order, ptr(y), ptr(jac) if order >= 1 else None, ptr(h2) if order >= 2 else None,
                    ptr(ws), nb_f, None)
        if not keep:
            ws[:] = 0                                   # the plain backward recomputes: nothing of the forward is needed
        gth = np.zeros(theta.size, np.float32); gx = np.full((N, D), np.nan, np.float32)
        lib.backward(desc, ptr(theta), ptr(x), N, order, ptr(gy), ptr(gj) if order >= 1 else None, ptr(gh) if order >= 2 else None,
                     ptr(gth), ptr(gx), ptr(ws), nb_b, None)
        res[keep] = (y, jac if order >= 1 else None, h2 if order >= 2 else None, gth, gx)
    for i, (a, b) in enumerate(zip(res[False], res[True])):
        if a is not None:
            if i == 3:          # parameter gradient: same arithmetic, but the global reductions are unordered atomics
                assert rel(b, a) < 1e-5
            else:
                assert np.array_equal(a, b)
    # a forward with the flag but only the forward-sized workspace is refused, loudly
    desc = _lib.make_desc(D, O, H, L, flags=_lib.FLAG_KEEP_TAPE)
    nb_f = lib.workspace_bytes(_lib.make_desc(D, O, H, L), N, order, False)
    with pytest.raises(RuntimeError, match="KEEP_TAPE"):
        lib.forward(desc, ptr(theta), ptr(x), N, order, ptr(res[True][0]), ptr(gj) if order >= 1 else None, ptr(gh) if order >= 2 else None,
                    ptr(ws), nb_f, None)
    assert not lib.tape_supported(_lib.make_desc(2, 1, 32, 3), N, 2)


def test_family_dispatch_rules(emu_library):
    fam = emu_library.kernel_family
    assert fam(_lib.make_desc(2, 1, 32, 3), 2, True) == 1
    assert fam(_lib.make_desc(2, 1, 32, 3, flags=_lib.FLAG_FORCE_GENERIC), 2, True) == 0
    assert fam(_lib.make_desc(2, 1, 33, 3), 2, True) == 2        # wider than the resident-weights kernel: tiled GEMM family
    assert fam(_lib.make_desc(3, 3, 512, 5), 3, True) == 2
    assert fam(_lib.make_desc(2, 1, 64, 3, flags=_lib.FLAG_FORCE_GENERIC), 2, True) == 0
    assert fam(_lib.make_desc(2, 1, 32, 4), 2, True) == 0        # deeper than the register-resident gW partials
    assert fam(_lib.make_desc(2, 1, 32, 4), 2, False) == 1
    assert fam(_lib.make_desc(2, 1, 32, 3), 3, True) == 0        # full Hessian streams: generic
    assert fam(_lib.make_desc(3, 3, 32, 3), 1, True) == 0


@pytest.mark.parametrize("case", [(2, 1, 32, 3, 50, 2, 1), (2, 2, 32, 3, 37, 1, 2), (1, 1, 20, 2, 64, 1, 1), (2, 2, 32, 2, 20, 0, 2)])
def test_fused_lsq_step_under_emulation(emu_library, case):
    lib = emu_library
    D, O, H, L, N, order, R = case
    rng = np.random.default_rng(5)
    desc = _lib.make_desc(D, O, H, L, flags=_lib.FLAG_NO_TENSOR)     # the emulator covers the FFMA kernels (no tape workspace)
    theta = make_theta(rng, D, O, H, L)
    x = rng.uniform(-1, 1, (N, D)).astype(np.float32)
    cy = rng.standard_normal((R, O)).astype(np.float32)
    cj = rng.standard_normal((R, O, D)).astype(np.float32) if order >= 1 else np.zeros((R, O, D), np.float32)
    cl = rng.standard_normal((R, O)).astype(np.float32) if order == 2 else np.zeros((R, O), np.float32)
    target = rng.standard_normal((N, R)).astype(np.float32)
    scale = 1.0 / (N * R)
    loss = np.zeros(1, np.float32)
    gth = np.zeros(theta.size, np.float32)
    coef = cy.ravel().tolist() + cj.ravel().tolist() + cl.ravel().tolist()
    lib.lsq_step(desc, ptr(theta), ptr(x), N, order, R, coef, ptr(target), scale, ptr(loss), ptr(gth), None, 0, None)
    out = fm.forward(theta.astype(np.float64), x.astype(np.float64), D, O, H, L, order)
    r = out["y"] @ cy.T.astype(np.float64) - target
    kw = {}
    if order >= 1:
        r = r + np.einsum("nod,rod->nr", out["jac"], cj.astype(np.float64))
    if order == 2:
        r = r + out["lap"] @ cl.T.astype(np.float64)
    kw["gy"] = 2 * scale * r @ cy.astype(np.float64)
    if order >= 1:
        kw["gjac"] = 2 * scale * np.einsum("nr,rod->nod", r, cj.astype(np.float64))
    if order == 2:
        kw["glap"] = 2 * scale * r @ cl.astype(np.float64)
    gref, _ = fm.backward(theta.astype(np.float64), x.astype(np.float64), D, O, H, L, order, **kw)
    ref_loss = scale * (r ** 2).sum()
    assert abs(float(loss[0]) - ref_loss) < 1e-5 * abs(ref_loss)
    assert rel(gth, gref) < 1e-4


def _svd_cases(rng, d, n):
    F = rng.standard_normal((n, d, d)).astype(np.float32)
    F[: n // 4] = np.eye(d, dtype=np.float32) + 0.05 * F[: n // 4]                 # near identity (the elasticity regime)
    F[n // 4] = np.eye(d)                                                            # exactly identity: coincident singular values
    F[n // 4 + 1] = 0.0                                                              # zero matrix
    F[n // 4 + 2] = np.outer(rng.standard_normal(d), rng.standard_normal(d))         # rank one
    F[n // 4 + 3, :, 0] *= -1.0                                                      # negative determinant
    return F


@pytest.mark.parametrize("d", [2, 3])
def test_svd_small_and_elastic_energy_under_emulation(emu_library, d):
    """insr_svd_small / insr_elastic_energy (elasticity/model.py:143-147) against numpy's SVD in fp64"""
    lib = emu_library
    rng = np.random.default_rng(11 + d)
    n = 200
    F = _svd_cases(rng, d, n)
    U = np.full((n, d, d), np.nan, np.float32); S = np.full((n, d), np.nan, np.float32); V = np.full((n, d, d), np.nan, np.float32)
    lib.svd_small(ptr(F), n, d, ptr(U), ptr(S), ptr(V), None)
    Sref = np.linalg.svd(F.astype(np.float64), compute_uv=False)
    assert np.abs(S - Sref).max() < 2e-6 * max(1.0, Sref.max())
    assert (S[:, :-1] >= S[:, 1:]).all() and (S >= 0).all()
    rec = np.einsum("nik,nk,njk->nij", U.astype(np.float64), S.astype(np.float64), V.astype(np.float64))
    assert np.abs(rec - F).max() < 5e-6 * max(1.0, np.abs(F).max())
    eye = np.eye(d)
    assert np.abs(np.einsum("nki,nkj->nij", U, U) - eye).max() < 5e-6
    assert np.abs(np.einsum("nki,nkj->nij", V, V) - eye).max() < 5e-6
    # fused energy + adjoint:  E = ra sum (s-1)^2 + rv sum (prod s - 1)^2 ;  dE/dF = U diag(dE/ds) V^T
    ra, rv = 0.7, 1.3
    E = np.zeros(1, np.float32); gF = np.full((n, d, d), np.nan, np.float32)
    lib.elastic_energy(ptr(F), n, d, ra, rv, ptr(E), ptr(gF), None)
    Eref = ra * ((Sref - 1) ** 2).sum() + rv * ((Sref.prod(1) - 1) ** 2).sum()
    assert abs(E[0] - Eref) < 1e-5 * abs(Eref)
    # directional finite difference in fp64 on well-conditioned samples
    Fd = F.astype(np.float64)
    Hd = rng.standard_normal(F.shape)
    def energy(M):
        s = np.linalg.svd(M, compute_uv=False)
        return ra * ((s - 1) ** 2).sum(1) + rv * ((s.prod(1) - 1) ** 2)
    h = 1e-6
    fd = (energy(Fd + h * Hd) - energy(Fd - h * Hd)) / (2 * h)
    an = (gF.astype(np.float64) * Hd).sum((1, 2))
    good = (Sref[:, -1] > 1e-2) & ((Sref[:, :-1] - Sref[:, 1:]).min(1) > 1e-3)       # smooth points of the singular values
    assert good.sum() > n // 2
    assert np.abs(fd - an)[good].max() < 2e-4 * np.abs(fd[good]).max()
    # energy-only call accumulates into the scalar
    lib.elastic_energy(ptr(F), n, d, ra, rv, ptr(E), None, None)
    assert abs(E[0] - 2 * Eref) < 2e-5 * abs(Eref)


def test_sample_boxes_under_emulation(emu_library):
    """insr_sample_boxes (base/sampling.py:14-18, 45-64 distributions): bounds, moments, fresh draws per iteration
    through the device counter, and shard consistency through point_offset"""
    lib = emu_library
    eps = 1e-4
    counts = [4096, 300, 300]
    lo = [[-1.0, -1.0], [-1 - eps, -1.0], [1 - eps, -1.0]]
    hi = [[1.0, 1.0], [-1 + eps, 1.0], [1 + eps, 1.0]]
    n = sum(counts)
    counter = np.zeros(1, np.int64); ticket = np.zeros(1, np.uint32)
    a = np.full((n, 2), np.nan, np.float32); b = np.full((n, 2), np.nan, np.float32)
    lib.sample_boxes(counts, lo, hi, 2, 1234, ptr(counter), ptr(ticket), 0, ptr(a), None)
    assert counter[0] == 1 and ticket[0] == 0
    lib.sample_boxes(counts, lo, hi, 2, 1234, ptr(counter), ptr(ticket), 0, ptr(b), None)
    assert counter[0] == 2
    off = 0
    for c, l, h in zip(counts, lo, hi):
        blk = a[off:off + c]
        assert (blk >= np.asarray(l, np.float32) - 1e-7).all() and (blk <= np.asarray(h, np.float32) + 1e-7).all()
        off += c
    x = a[:4096]
    assert abs(x.mean()) < 0.05 and abs(x.var() - 1.0 / 3.0) < 0.03           # U[-1,1]: mean 0, variance 1/3
    assert abs(np.corrcoef(x[:, 0], x[:, 1])[0, 1]) < 0.06
    assert not np.array_equal(a, b) and len(np.unique(a[:4096, 0])) > 4000      # fresh draws, no repeats
    # same seed + same iteration + point offset = the same global stream, whichever rank draws it
    counter[0] = 0
    c1 = np.zeros((2048, 2), np.float32); c2 = np.zeros((2048, 2), np.float32)
    lib.sample_boxes([2048], [lo[0]], [hi[0]], 2, 1234, None, None, 0, ptr(c1), None)
    lib.sample_boxes([2048], [lo[0]], [hi[0]], 2, 1234, None, None, 2048, ptr(c2), None)
    assert np.array_equal(np.concatenate([c1, c2]), a[:4096])


def _cube_tets():
    """unit cube [0,1]^3 split into 6 tetrahedra around the main diagonal + a stretched copy (different volumes)"""
    corners = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float32)
    idx = lambda x, y, z: 4 * x + 2 * y + z
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = []
    for p in perms:
        cur = [0, 0, 0]; path = [idx(*cur)]
        for ax in p:
            cur[ax] = 1; path.append(idx(*cur))
        tets.append(path)
    far = corners * np.array([3.0, 1.0, 1.0], np.float32) + np.array([2.0, 0.0, 0.0], np.float32)   # [2,5]x[0,1]^2
    V = np.concatenate([corners, far]).astype(np.float32)
    T = np.array(tets + [[v + 8 for v in t] for t in tets], np.int32)
    return V, T


def test_sample_mesh_under_emulation(emu_library):
    """insr_sample_mesh: volume-weighted tetrahedra with Dirichlet(1,1,1,1) weights (torchgp/sample_volume.py:9-43) and
    area-weighted triangles with the sqrt(u) weights (torchgp/sample_surface.py:28-52) -- points lie inside the mesh,
    are uniform over it (element frequencies ~ measure, first moments), are fresh per iteration and shard by offset"""
    lib = emu_library
    V, T = _cube_tets()
    P = V[T]
    vol = np.abs(np.einsum("ij,ij->i", P[:, 3] - P[:, 0], np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]))) / 6
    cdf = (np.cumsum(vol.astype(np.float64)) / vol.sum()).astype(np.float32)
    n = 20000
    counter = np.zeros(1, np.int64); ticket = np.zeros(1, np.uint32)
    a = np.full((n, 3), np.nan, np.float32); b = np.full((n, 3), np.nan, np.float32)
    lib.sample_mesh(ptr(V), ptr(T), ptr(cdf), len(T), 4, n, 3, 77, ptr(counter), ptr(ticket), 0, ptr(a), None)
    lib.sample_mesh(ptr(V), ptr(T), ptr(cdf), len(T), 4, n, 3, 77, ptr(counter), ptr(ticket), 0, ptr(b), None)
    assert counter[0] == 2 and ticket[0] == 0 and not np.array_equal(a, b)
    assert np.isfinite(a).all()
    near = a[:, 0] <= 1.0 + 1e-6
    far = a[:, 0] >= 2.0 - 1e-6
    assert (near | far).all()                                              # nothing in the gap between the two boxes
    assert (a[:, 1:] >= -1e-6).all() and (a[:, 1:] <= 1 + 1e-6).all() and a[:, 0].min() >= -1e-6 and a[:, 0].max() <= 5 + 1e-6
    assert abs(far.mean() - 0.75) < 0.015                                  # volume 3 of 4
    # uniform inside each box: mean at the centre, variance (edge^2)/12
    assert np.abs(a[near].mean(0) - 0.5).max() < 0.02 and np.abs(a[near].var(0) - 1 / 12).max() < 0.01
    assert np.abs(a[far].mean(0) - np.array([3.5, 0.5, 0.5])).max() < 0.03 and abs(a[far][:, 0].var() - 9 / 12) < 0.03
    # dim_out slice and rank offset: the second half of a global draw is the draw at offset n/2
    h = n // 2
    c2 = np.zeros((h, 2), np.float32)
    counter[0] = 0
    lib.sample_mesh(ptr(V), ptr(T), ptr(cdf), len(T), 4, h, 2, 77, None, None, h, ptr(c2), None)
    assert np.array_equal(c2, a[h:, :2])
    # triangles: the two faces of the unit square in the plane z = 0 plus a bigger triangle
    Vt = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [2, 0, 0], [4, 0, 0], [2, 2, 0]], np.float32)
    Ft = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6]], np.int32)
    area = np.array([0.5, 0.5, 2.0]); cdf_t = (np.cumsum(area) / area.sum()).astype(np.float32)
    s = np.full((n, 2), np.nan, np.float32)
    lib.sample_mesh(ptr(Vt), ptr(Ft), ptr(cdf_t), 3, 3, n, 2, 5, None, None, 0, ptr(s), None)
    big = s[:, 0] >= 2 - 1e-6
    assert abs(big.mean() - 2 / 3) < 0.015
    sq = s[~big]
    assert (sq >= -1e-6).all() and (sq <= 1 + 1e-6).all() and np.abs(sq.mean(0) - 0.5).max() < 0.02 and np.abs(sq.var(0) - 1 / 12).max() < 0.01
    tri = s[big]
    assert ((tri[:, 0] - 2) + tri[:, 1] <= 2 + 1e-5).all() and (tri[:, 1] >= -1e-6).all()
    assert np.abs(tri.mean(0) - np.array([2 + 2 / 3, 2 / 3])).max() < 0.03     # centroid


@pytest.mark.parametrize("d", [2, 3])
def test_elastic_terms_under_emulation(emu_library, d):
    """insr_elastic_terms: every term of elasticity/model.py:127-189 + elasticity/losses.py:6-39 and its cotangents
    against the same expressions under torch autograd in fp64 (svdvals energies, kinematics, external force, both
    constraints, plane collision; the 2-D sphere collision for d = 2)"""
    import torch
    lib = emu_library
    rng = np.random.default_rng(40 + d)
    n, nl, nr = 300, 17, 23
    na = n + nl + nr
    y = (0.3 * rng.standard_normal((na, d))).astype(np.float32)
    J = (0.25 * rng.standard_normal((na, d, d))).astype(np.float32)
    x = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    yp = (y[:n] + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    ypp = (yp + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    t = _lib.ElasticTermsDesc()
    t.n, t.n_left, t.n_right, t.dt = n, nl, nr, 0.05
    t.r_arap, t.r_volume, t.r_kinematics, t.r_left, t.r_right = 2.0, 30.0, 1.5, 100.0, 80.0
    t.r_plane, t.plane_height = 7.0, -0.2
    t.r_sphere, t.radius = (5.0, 0.8) if d == 2 else (0.0, 0.0)
    ext, off, cen = [0.3, -2.0, 0.7], [0.4, -0.1, 0.2], [0.1, -0.3, 0.0]
    for i in range(3):
        t.external_force[i], t.offset_right[i], t.center[i] = ext[i], off[i], cen[i]
    loss = np.zeros(1, np.float32); gy = np.full((na, d), np.nan, np.float32); gJ = np.full((na, d, d), np.nan, np.float32)
    lib.elastic_terms(t, d, ptr(y), ptr(J), ptr(x), ptr(yp), ptr(ypp), ptr(loss), ptr(gy), ptr(gJ), None)

    Y = torch.tensor(y, dtype=torch.float64, requires_grad=True); Jt = torch.tensor(J, dtype=torch.float64, requires_grad=True)
    X, YP, YPP = (torch.tensor(a, dtype=torch.float64) for a in (x, yp, ypp))
    dt = 0.05
    q, q_prev, q_pp = Y[:n] + X, YP + X, YPP + X
    qdot, qdot_prev = (q - q_prev) / dt, (q_prev - q_pp) / dt
    S = torch.linalg.svdvals(Jt[:n] + torch.eye(d, dtype=torch.float64))
    E = 2.0 * ((S - 1) ** 2).sum() + 30.0 * ((S.prod(1) - 1) ** 2).sum()
    E = E + 1.5 * ((qdot - qdot_prev) ** 2).sum() - dt * (qdot * torch.tensor(ext[:d], dtype=torch.float64)).sum()
    E = E + 100.0 * (Y[n:n + nl] ** 2).sum() + 80.0 * ((Y[n + nl:] - torch.tensor(off[:d], dtype=torch.float64)) ** 2).sum()
    hit = (q[:, -1] < -0.2).double()
    assert 0 < hit.sum() < n
    E = E - dt * 7.0 * (hit * qdot[:, -1] * (-0.2 - q[:, -1])).sum()
    if d == 2:
        vec = q - torch.tensor(cen[:2], dtype=torch.float64)
        inside = (vec.norm(dim=1) < 0.8).double()
        assert 0 < inside.sum() < n
        E = E - dt * 5.0 * (inside[:, None] * qdot * vec).sum()
    E.backward()
    E = E.detach()
    assert abs(loss[0] - float(E)) < 3e-5 * abs(float(E))
    assert rel(gy, Y.grad.numpy()) < 3e-5
    assert rel(gJ[:n], Jt.grad.numpy()[:n]) < 1e-4 and not gJ[n:].any()
    # accumulation into the loss word, and the energy-free form (no J): gJ untouched
    t.r_arap = t.r_volume = 0.0
    before = float(loss[0])
    lib.elastic_terms(t, d, ptr(y), None, ptr(x), ptr(yp), ptr(ypp), ptr(loss), ptr(gy), None, None)
    E_s = 2.0 * ((S - 1) ** 2).sum() + 30.0 * ((S.prod(1) - 1) ** 2).sum()
    assert abs((loss[0] - before) - float(E - E_s.detach())) < 1e-4 * abs(float(E))
    with pytest.raises(RuntimeError, match="arap / volume need J"):
        t.r_arap = 1.0
        lib.elastic_terms(t, d, ptr(y), None, ptr(x), ptr(yp), ptr(ypp), ptr(loss), ptr(gy), None, None)
