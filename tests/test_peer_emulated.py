"""The peer-memory exchange kernels (csrc/peer_kernels.cuh) on the host SIMT emulator: the RANKS are host threads of this
process that call the C ABI concurrently -- as the ranks of a box do -- and "peer memory" is ordinary memory every thread can
address (tests/emu: the emulation build's insr_peer_alloc hands out host buffers, its st.release / ld.acquire are C++ atomics).
What this exercises without a GPU: the flag barrier with its monotonically increasing epochs over many rounds, the rank-order
sum (bit-identical on every rank), tails and unaligned outputs of the stand-alone all-reduce, and the exchange fused into the
iteration update against torch.optim.Adam + ReduceLROnPlateau on the averaged gradient."""
import ctypes
import threading
import time

import numpy as np
import pytest
import torch

HEADER = 4096


def view(addr, n):
    return torch.frombuffer((ctypes.c_char * (4 * n)).from_address(addr), dtype=torch.float32)


def run_ranks(world, fn):
    errors = []

    def guard(r):
        try:
            fn(r)
        except BaseException as e:          # noqa: BLE001 -- reported by the test below
            errors.append((r, repr(e)))
    threads = [threading.Thread(target=guard, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a barrier"
    assert not errors, errors


@pytest.mark.parametrize("world", [2, 3])
def test_one_shot_allreduce_between_emulated_ranks(emu_library, world):
    lib = emu_library
    for n in (5, 1030, 2600):
        allocs = [lib.peer_alloc(4 * n) for _ in range(world)]
        bases = [lib.peer_open(h) for _, h in allocs]
        assert bases == [b for b, _ in allocs]
        data = [view(b + HEADER, n) for b in bases]
        outs = [torch.full((n + 1,), -7.0) for _ in range(world)]
        rounds, scale = 5, 0.5

        def value(r, it):
            return torch.randn(n, generator=torch.Generator().manual_seed(1000 * it + r))

        def rank(r):
            for it in range(rounds):
                if (it + r) % 3 == 0:
                    time.sleep(0.03 * (r + 1))                  # skewed arrivals: a rank may be a whole call behind the others
                data[r].copy_(value(r, it))                     # this rank's "gradient" of the round
                out = outs[r][1:] if it % 2 else outs[r][:n]    # odd rounds: an output that is not 16-byte aligned
                lib.peer_allreduce(world, r, bases, HEADER // 4, n, scale, out.data_ptr(), None)
                want = value(0, it)
                for q in range(1, world):
                    want = want + value(q, it)                  # rank order, like the kernel
                assert torch.equal(out, want * scale), (r, it, float((out - want * scale).abs().max()))
        run_ranks(world, rank)
        assert all(lib.peer_status(b) == 0 for b in bases)
        for b in bases:
            epoch = ctypes.c_uint32.from_address(b + 4 * 512).value
            assert epoch == 1 + 2 * rounds                      # two barriers per call, published by the last CTA
            lib.peer_free(b)


def test_exchange_fused_into_the_iteration_update_between_emulated_ranks(emu_library):
    lib = emu_library
    world, sizes, n_losses, iters = 2, [301, 77], 2, 6
    padded = [(s + 3) // 4 * 4 for s in sizes]
    n_buf = sum(padded) + 4
    peer_bytes = HEADER + (4 * n_buf + 255) // 256 * 256
    allocs = [lib.peer_alloc(4 * n_buf) for _ in range(world)]
    bases = [b for b, _ in allocs]
    bufs = [view(b + HEADER, n_buf) for b in bases]
    torch.manual_seed(0)
    theta0 = [torch.randn(s) for s in sizes]
    state = []
    for r in range(world):
        thetas = [t.clone() for t in theta0]
        ms, vs = [torch.zeros(s) for s in sizes], [torch.zeros(s) for s in sizes]
        grads, off = [], 0
        for s, pd in zip(sizes, padded):
            grads.append(bufs[r][off:off + s])
            off += pd
        state.append(dict(thetas=thetas, ms=ms, vs=vs, grads=grads, losses=bufs[r][off:off + n_losses],
                          sched=torch.tensor([1e-2, float("inf"), 0.0, 0.0]), red=torch.zeros(32),
                          hist=torch.full((16, n_losses), -1.0), idx=torch.zeros(1, dtype=torch.long)))

    def grad_of(r, it, k):
        return torch.randn(sizes[k], generator=torch.Generator().manual_seed(77 * it + 5 * r + k)) * (1 + it)

    def loss_of(r, it):
        return torch.tensor([3.0 + r + it, [1.0, 0.9, 0.95, 0.96, 0.97, 0.98][it] + 0.01 * r])

    def rank(r):
        st = state[r]
        for it in range(iters):
            for k in range(len(sizes)):
                st["grads"][k].copy_(grad_of(r, it, k))
            st["losses"].copy_(loss_of(r, it))
            lib.iteration_update_peer(world, r, bases, peer_bytes, 1.0 / world,
                                      [t.data_ptr() for t in st["thetas"]], [g.data_ptr() for g in st["grads"]],
                                      [m.data_ptr() for m in st["ms"]], [v.data_ptr() for v in st["vs"]], sizes,
                                      st["sched"].data_ptr(), st["losses"].data_ptr(), n_losses, 1, st["red"].data_ptr(),
                                      st["hist"].data_ptr(), st["hist"].shape[0], st["idx"].data_ptr(), 0.9, 0.999, 1e-8,
                                      0.1, 2, 1e-4, 1e-8, 1e-8, True, True, None)
            assert all(not g.any() for g in st["grads"]) and not st["losses"].any()      # left zeroed for the next iteration
    run_ranks(world, rank)

    # reference: Adam + ReduceLROnPlateau on the rank-averaged gradient / main loss (base/baseModel.py:55-81)
    refs = [torch.nn.Parameter(t.clone()) for t in theta0]
    opt = torch.optim.Adam(refs, lr=1e-2)
    sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, min_lr=1e-8, patience=2)
    for it in range(iters):
        for k, p in enumerate(refs):
            p.grad = sum(grad_of(r, it, k) for r in range(world)) / world
        opt.step()
        mean_losses = sum(loss_of(r, it) for r in range(world)) / world
        sch.step(float(mean_losses[1]))
        for r in range(world):
            assert torch.allclose(state[r]["hist"][it], mean_losses, rtol=1e-6, atol=0)
    assert opt.param_groups[0]["lr"] < 1e-2                                               # the schedule did cut
    for r in range(world):
        st = state[r]
        assert abs(float(st["sched"][0]) - opt.param_groups[0]["lr"]) < 1e-9 and int(st["idx"]) == iters
        for k in range(len(sizes)):
            assert float((st["thetas"][k] - refs[k].detach()).abs().max()) < 2e-6
            assert torch.equal(st["thetas"][k], state[0]["thetas"][k])                   # replicas bit-identical
    assert all(lib.peer_status(b) == 0 for b in bases)
    for b in bases:
        lib.peer_free(b)
