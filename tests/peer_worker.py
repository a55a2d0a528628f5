"""Worker of tests/test_peer_gpu.py, run under torchrun (one process per GPU): the peer-memory exchange against NCCL.
Prints one JSON line on rank 0; every check is evaluated on EVERY rank and combined (a failure on any rank fails the test)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import insr_pde_b200 as ib  # noqa: E402
from insr_pde_b200 import fused, peer  # noqa: E402


def all_ok(flag, dev):
    t = torch.tensor([1.0 if flag else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item() > 0.5)


def same_on_all_ranks(t):
    """bit-identical on every rank"""
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t.contiguous())
    return all(torch.equal(parts[0], p) for p in parts[1:])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}

    # ---- 1. the stand-alone one-shot all-reduce against NCCL, many rounds (the epoch words keep cycling), several sizes
    ok_sizes = {}
    for n in (5, 3297, 50435, 1 << 20):
        pb = peer.PeerBuffer.create(n, dev)
        if not all_ok(pb is not None, dev):
            out["peer_memory"] = "unavailable"
            if rank == 0:
                print(json.dumps(out))
            dist.destroy_process_group()
            return
        res = torch.empty(n, device=dev)
        good, ident = True, True
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        for it in range(40):
            pb.data.copy_(torch.randn(n, generator=gen, device=dev))
            want = pb.data.clone()
            dist.all_reduce(want)
            pb.allreduce_into(res, scale=0.5)
            good = good and bool(torch.allclose(res, 0.5 * want, rtol=1e-5, atol=1e-6))
            if it % 13 == 0:
                ident = ident and same_on_all_ranks(res)
        # ---- 2. the same inside a CUDA graph (what the iteration graph does): replays keep the barrier epochs consistent
        src = torch.randn(n, generator=gen, device=dev)
        g = torch.cuda.CUDAGraph()
        pb.data.copy_(src)
        pb.allreduce_into(res)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            pb.data.copy_(src)
            pb.allreduce_into(res)
        for _ in range(25):
            g.replay()
        want = src.clone()
        dist.all_reduce(want)
        good = good and bool(torch.allclose(res, want, rtol=1e-5, atol=1e-6))
        ok_sizes[str(n)] = all_ok(good and ident and pb.healthy(), dev)
        del g
        pb.close()
    out["allreduce_matches_nccl"] = ok_sizes

    # ---- 3. the exchange fused into the iteration update: peer path against NCCL all-reduce + insr_iteration_update
    def nets():
        torch.manual_seed(7)
        return [ib.MLP(2, 2, 3, 32, nonlinearity="sine").to(dev), ib.MLP(2, 1, 3, 32, nonlinearity="sine").to(dev)]

    def run(use_peer, iters=12):
        os.environ["INSR_PEER_ALLREDUCE"] = "1" if use_peer else "0"
        ns = nets()
        shared = fused.SharedGradBuffer(ns)
        opt = fused.DeviceOptimizer(ns, 1e-3)
        hist = torch.zeros(64, 2, device=dev)
        idx = torch.zeros(1, dtype=torch.long, device=dev)
        gen = torch.Generator(device=dev).manual_seed(200 + rank)
        used_peer = shared.peer is not None
        for _ in range(iters):
            for n in ns:                                   # this rank's gradient and loss terms
                fused.flat_grad(n).copy_(torch.randn(fused.flat_grad(n).numel(), generator=gen, device=dev) * 1e-2)
            shared.scalars[:2].copy_(torch.rand(2, generator=gen, device=dev))
            if used_peer:
                opt.update_peer(shared.peer, shared.scalars[:2], 0, hist, idx, clear_losses=True)
            else:
                shared.allreduce(shared.scalars[:2])
                opt.update(shared.scalars[:2], 0, hist, idx, clear_losses=True)
        torch.cuda.synchronize()
        thetas = torch.cat([n.flat_theta().detach().clone() for n in ns])
        zeroed = all(float(fused.flat_grad(n).abs().max()) == 0.0 for n in ns) and float(shared.scalars[:2].abs().max()) == 0.0
        res = (thetas, hist[:iters].clone(), opt.sched.clone(), zeroed, used_peer, shared.peer.healthy() if used_peer else True)
        shared.close()
        return res

    th_p, hist_p, sched_p, zero_p, used, healthy = run(True)
    th_n, hist_n, sched_n, zero_n, _, _ = run(False)
    out["fused_update"] = {
        "used_peer": all_ok(used, dev), "healthy": all_ok(healthy, dev),
        "theta_matches_nccl": all_ok(torch.allclose(th_p, th_n, rtol=2e-5, atol=1e-7), dev),
        "log_matches_nccl": all_ok(torch.allclose(hist_p, hist_n, rtol=1e-5, atol=1e-7), dev),
        "schedule_matches_nccl": all_ok(torch.allclose(sched_p, sched_n), dev),
        "grads_and_losses_left_zeroed": all_ok(zero_p and zero_n, dev),
        "replicas_identical": all_ok(same_on_all_ranks(th_p), dev),
        "max_theta_diff": float((th_p - th_n).abs().max()),
    }

    # ---- 4. the graphed data-parallel fluid time step: peer exchange against the NCCL exchange, replicas identical
    def fluid(use_peer):
        os.environ["INSR_PEER_ALLREDUCE"] = "1" if use_peer else "0"
        torch.manual_seed(3)
        vel, prev, pres = (ib.MLP(2, 2, 3, 32, nonlinearity="sine").to(dev), ib.MLP(2, 2, 3, 32, nonlinearity="sine").to(dev),
                           ib.MLP(2, 1, 3, 32, nonlinearity="sine").to(dev))
        st = fused.FluidStepper(vel, prev, pres, sample_resolution=64, graphed=True, device_sampler=True, seed=11)
        st.data_parallel = True
        h0 = st.initialize(fused.taylorgreen_velocity, 30, world=world)
        h1, h2, h3 = st.step(30, world=world)
        used_peer = all(lp.shared is not None and lp.shared.peer is not None for lp in st._loops.values())
        theta = torch.cat([vel.flat_theta().detach().clone(), pres.flat_theta().detach().clone()])
        losses = torch.tensor([[d["main"] for d in h] for h in (h0, h1, h2, h3)], dtype=torch.float64)
        st.close()
        return theta, losses, used_peer

    tp, lp_, used_f = fluid(True)
    tn, ln_, _ = fluid(False)
    out["fluid_timestep"] = {
        "used_peer": all_ok(used_f, dev),
        "replicas_identical": all_ok(same_on_all_ranks(tp), dev),
        "loss_history_matches_nccl": all_ok(bool(torch.allclose(lp_, ln_, rtol=2e-3, atol=1e-7)), dev),
        "theta_matches_nccl": all_ok(bool(torch.allclose(tp, tn, rtol=0, atol=5e-4)), dev),
        "max_theta_diff": float((tp - tn).abs().max()), "final_losses_peer": [float(v) for v in lp_[:, -1]],
        "final_losses_nccl": [float(v) for v in ln_[:, -1]],
    }
    if rank == 0:
        print(json.dumps(out))
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
