"""Host-side decisions of the peer-memory exchange (insr_pde_b200/peer.py) that need no GPU: when peer memory is not
attempted at all, and that the shared gradient buffer then is ordinary memory reduced by the process group's all-reduce."""
import os

import torch

import insr_pde_b200 as ib
from insr_pde_b200 import fused, peer


def test_peer_memory_is_not_attempted_without_a_multi_rank_process_group(monkeypatch):
    assert peer.PeerBuffer.create(128, torch.device("cpu")) is None            # no process group
    monkeypatch.setenv("INSR_PEER_ALLREDUCE", "0")
    assert not peer.enabled()
    monkeypatch.setenv("INSR_PEER_ALLREDUCE", "1")
    assert peer.enabled()
    assert peer.HEADER_BYTES == 4096 and peer.HEADER_FLOATS * 4 == peer.HEADER_BYTES


def test_shared_gradient_buffer_on_cpu_tensors_is_plain_memory():
    torch.manual_seed(0)
    nets = [ib.MLP(2, 2, 1, 8, nonlinearity="sine"), ib.MLP(2, 1, 1, 8, nonlinearity="sine")]
    shared = fused.SharedGradBuffer(nets)
    assert shared.peer is None and shared.buf.device.type == "cpu"
    sizes = [n.flat_theta().numel() for n in nets]
    assert shared.buf.numel() == sum((s + 3) // 4 * 4 for s in sizes) + 4
    for n in nets:                                                              # the nets' gradients ARE slices of the buffer
        g = fused.flat_grad(n)
        assert shared.buf.data_ptr() <= g.data_ptr() < shared.buf.data_ptr() + 4 * shared.buf.numel()
        assert (g.data_ptr() - shared.buf.data_ptr()) % 16 == 0
    vals = shared.allreduce(torch.tensor([1.5, 2.5]))                           # single process: values pass through unchanged
    assert [float(v) for v in vals] == [1.5, 2.5] and shared.scalars[:2].tolist() == [1.5, 2.5]
    own = shared.allreduce(shared.scalars[:2])                                  # already in place: no copy, same storage
    assert own[0].data_ptr() == shared.scalars.data_ptr()
    shared.close()                                                              # nothing to release without peer memory
    assert shared.buf is not None
