import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from insr_pde_b200 import linalg
for d, n in ((2, 20000), (3, 26592), (2, 500000)):
    F = (torch.eye(d, device="cuda") + 0.1 * torch.randn(n, d, d, device="cuda")).requires_grad_(True)
    def ref():
        F.grad = None
        _, S, _ = torch.svd(F)
        E = 3.0 * ((S - 1) ** 2).sum() + 40.0 * ((S.prod(1) - 1) ** 2).sum()
        E.backward()
    def ours_svd():
        F.grad = None
        _, S, _ = linalg.svd(F)
        E = 3.0 * ((S - 1) ** 2).sum() + 40.0 * ((S.prod(1) - 1) ** 2).sum()
        E.backward()
    def ours_fused():
        F.grad = None
        linalg.elastic_energy(F, 3.0, 40.0).backward()
    for name, fn, reps in (("torch.svd + autograd", ref, 5), ("insr_svd_small + autograd", ours_svd, 50), ("insr_elastic_energy", ours_fused, 50)):
        fn(); fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        print(f"d={d} n={n} {name}: {(time.perf_counter() - t0) / reps * 1e3:.3f} ms", flush=True)
