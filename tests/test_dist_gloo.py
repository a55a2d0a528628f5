"""world_size-2 gloo test of the data-parallel host logic (CPU): sharded points + one flat
all-reduce reproduce the single-process gradient and keep replicas identical through Adam."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from insr_pde_b200 import dist as idist
    from oracle import closures, torch_port as tp
    torch.set_num_threads(1)
    r, w = idist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                               # identical replicas
    vel, pres = tp.RefMLP(2, 2, 3, 16), tp.RefMLP(2, 1, 3, 16)
    g = torch.Generator().manual_seed(7)               # identical GLOBAL sample set on every rank
    x_all = torch.rand(256, 2, generator=g) * 2 - 1
    bcx = tp.sample_boundary2D_separate(8, "horizontal")
    bcy = tp.sample_boundary2D_separate(8, "vertical")
    dist.broadcast(bcx, 0); dist.broadcast(bcy, 0)
    opt = torch.optim.Adam(list(vel.parameters()) + list(pres.parameters()), lr=1e-3)
    reducer = idist.GradAllReducer([vel, pres])
    reducer.install(opt)
    for it in range(3):
        x = idist.shard_points(x_all).clone().requires_grad_(True)
        assert x.shape[0] == 256 // world
        loss = closures.fluid_solve_pressure(vel, pres, tp, x, bcx.clone().requires_grad_(True),
                                             bcy.clone().requires_grad_(True))
        opt.zero_grad()
        sum(loss.values()).backward()
        if it == 0:
            g_local = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in pres.parameters()]).clone()
            extra = reducer.allreduce(extra_scalars=torch.stack([v.detach() for v in loss.values()]))
            g_red = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in pres.parameters()]).clone()
            np.savez(os.path.join(out_dir, f"rank{rank}.npz"), g_local=g_local.numpy(), g_red=g_red.numpy(),
                     extra=extra.numpy())
            opt.zero_grad()
            sum(closures.fluid_solve_pressure(vel, pres, tp, x, bcx.clone().requires_grad_(True),
                                              bcy.clone().requires_grad_(True)).values()).backward()
        opt.step()                                      # pre-hook all-reduces
    theta = torch.cat([p.detach().reshape(-1) for p in pres.parameters()])
    np.save(os.path.join(out_dir, f"theta{rank}.npy"), theta.numpy())
    dist.destroy_process_group()


def test_sharded_gradients_match_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["g_red"], r1["g_red"]) and np.array_equal(r0["extra"], r1["extra"])
    assert np.allclose(r0["g_red"], 0.5 * (r0["g_local"] + r1["g_local"]), rtol=1e-5, atol=1e-7)
    assert not np.allclose(r0["g_local"], r1["g_local"])
    # replicas stay bit-identical through the optimizer
    assert np.array_equal(np.load(tmp_path / "theta0.npy"), np.load(tmp_path / "theta1.npy"))
    # and equal the single-process run on the full batch (the bc terms are identical on every rank)
    import sys
    sys.path.insert(0, ROOT)
    from oracle import closures, torch_port as tp
    torch.manual_seed(0)
    vel, pres = tp.RefMLP(2, 2, 3, 16), tp.RefMLP(2, 1, 3, 16)
    g = torch.Generator().manual_seed(7)
    x_all = (torch.rand(256, 2, generator=g) * 2 - 1).requires_grad_(True)
    # same boundary points as the workers drew (rank 0's draw after manual_seed(0) + net construction)
    bcx = tp.sample_boundary2D_separate(8, "horizontal").requires_grad_(True)
    bcy = tp.sample_boundary2D_separate(8, "vertical").requires_grad_(True)
    loss = closures.fluid_solve_pressure(vel, pres, tp, x_all, bcx, bcy)
    sum(loss.values()).backward()
    g_full = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in pres.parameters()]).numpy()
    assert np.abs(g_full - r0["g_red"]).max() <= 2e-5 * np.abs(g_full).max()


def test_shard_partition_is_exact():
    from insr_pde_b200.sampling import shard
    x = torch.arange(103 * 2, dtype=torch.float32).reshape(103, 2)
    for world in (1, 2, 4, 8):
        parts = [shard(x, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts), x)
        assert max(p.shape[0] for p in parts) - min(p.shape[0] for p in parts) <= 1


def _worker_shared(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import insr_pde_b200 as ib
    from insr_pde_b200 import dist as idist, fused
    torch.set_num_threads(1)
    idist.init_from_env("gloo")
    torch.manual_seed(0)
    vel, pres = ib.MLP(2, 2, 3, 32, nonlinearity="sine"), ib.MLP(2, 1, 3, 32, nonlinearity="sine")   # 3330 / 3297 parameters
    shared = fused.SharedGradBuffer([vel, pres])
    gv, gp = fused.flat_grad(vel), fused.flat_grad(pres)
    # the nets' gradient buffers are 16-byte aligned slices of the one buffer, and the parameters' .grad are views of them
    assert gv.data_ptr() == shared.buf.data_ptr() and gp.data_ptr() == shared.buf.data_ptr() + 4 * 3332
    assert gp.data_ptr() % 16 == 0 and shared.scalars.data_ptr() == shared.buf.data_ptr() + 4 * (3332 + 3300)
    assert next(iter(pres.parameters())).grad.data_ptr() == gp.data_ptr()
    gv.fill_(float(rank + 1)); gp.copy_(torch.arange(3297, dtype=torch.float32) * (rank + 1))
    main, bc = shared.allreduce([torch.tensor(2.0 * rank), torch.tensor(10.0 + rank)])
    main, bc = float(main), float(bc)
    # the fused closures hand their loss terms over as ONE vector (fused.Losses.vector): same collective, no stack
    gv.fill_(float(rank + 1)); gp.copy_(torch.arange(3297, dtype=torch.float32) * (rank + 1))
    vec = shared.allreduce(torch.tensor([4.0 * rank, 1.0]))
    assert float(vec[0]) == 2.0 and float(vec[1]) == 1.0 and vec[0].data_ptr() == shared.scalars.data_ptr()
    np.savez(os.path.join(out_dir, f"shared{rank}.npz"), gv=gv.numpy(), gp=gp.numpy(), main=float(main), bc=float(bc),
             pgrad=list(pres.parameters())[-1].grad.numpy())


def test_shared_gradient_buffer_is_one_allreduce(tmp_path):
    """fused.SharedGradBuffer (one flat buffer [gradients of all nets | loss values] per iteration): slices
    aligned for the C ABI, parameters' .grad alias the buffer, one collective averages gradients and loss terms"""
    world = 2
    mp.spawn(_worker_shared, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"shared{k}.npz") for k in range(world)]
    for k in range(world):
        assert np.all(r[k]["gv"] == 1.5) and np.allclose(r[k]["gp"], np.arange(3297) * 1.5)
        assert r[k]["main"] == 1.0 and r[k]["bc"] == 10.5
        assert np.allclose(r[k]["pgrad"], r[k]["gp"][-1:])          # last parameter (output bias) = last element of the slice


def _worker_plateau(rank, world, port, out_dir):
    """the INTEGRATION.md recipe (install_global + shard_points under a loop shaped like base/baseModel.py:73-81,104-134)
    run PAST a ReduceLROnPlateau cut and the early-stop exit"""
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from insr_pde_b200 import dist as idist
    from oracle import torch_port as tp
    torch.set_num_threads(1)
    idist.init_from_env("gloo")
    torch.manual_seed(0)
    net = tp.RefMLP(1, 1, 1, 8)
    idist.install_global(lambda: [net])
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, min_lr=1e-5, patience=1)
    g = torch.Generator().manual_seed(3)
    lrs, n_iters = [], 0
    for it in range(60):
        x_all = torch.rand(64, 1, generator=g) * 2 - 1              # the same global set on every rank
        x = idist.shard_points(x_all)
        # shard-dependent noise: the LOCAL losses of the two ranks go up and down at different iterations
        loss = torch.mean((net(x) - torch.sin(3 * x)) ** 2) * (1.0 + 0.5 * ((it + rank) % 2))
        opt.zero_grad()
        loss.backward()
        opt.step()                                                  # pre-hook: gradients averaged
        sched.step(loss)                                            # patched: sees the rank-averaged loss
        lrs.append(opt.param_groups[0]["lr"])
        n_iters += 1
        if opt.param_groups[0]["lr"] <= 1.1e-5:                     # early stop as base/baseModel.py:132-134
            break
    theta = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    np.savez(os.path.join(out_dir, f"plateau{rank}.npz"), theta=theta.numpy(), lrs=np.array(lrs), n_iters=n_iters)
    dist.barrier()
    dist.destroy_process_group()


def test_replicas_stay_identical_through_plateau_cuts_and_early_stop(tmp_path):
    world = 2
    mp.spawn(_worker_plateau, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "plateau0.npz"), np.load(tmp_path / "plateau1.npz")
    assert int(a["n_iters"]) == int(b["n_iters"]) < 60              # both left the loop, at the same iteration
    assert np.array_equal(a["lrs"], b["lrs"]) and a["lrs"][-1] < a["lrs"][0]      # same LR decisions, and cuts happened
    assert np.array_equal(a["theta"], b["theta"])
