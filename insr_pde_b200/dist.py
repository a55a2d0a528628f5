"""Data parallelism over collocation points: one process per GPU, weights
replicated, points sharded, ONE all-reduce per optimisation iteration of a flat FP32 buffer
[all parameter gradients of all trainable nets || loss scalars] over NCCL (NVLink 5 / NVSwitch).

Every loss of the reference is a mean / sum over points (fluid/model.py:89,113,140;
elasticity/model.py:146-149) and points never interact, so the G-rank gradient equals the
1-rank gradient up to summation order when rank r takes the contiguous slice r of the same
global sample set.  After the all-reduce every rank applies the identical Adam step, so no
broadcast is needed.  The payload is 3.6-56 KB for the script configs (latency-bound).

The reference's training runtime (base/baseModel.py:55-81) is left untouched: the reducer is
attached as an *optimizer step pre-hook*, globally, because ``_reset_optimizer`` builds a fresh
Adam for every training loop.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .sampling import shard


def init_from_env(backend: str | None = None):
    """torchrun-style init (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world


def shard_points(points, group=None):
    """this rank's contiguous slice of a global (N, D) sample set"""
    if not dist.is_initialized():
        return points
    return shard(points, dist.get_rank(group), dist.get_world_size(group))


class GradAllReducer:
    """all-reduce (mean) the gradients of a set of modules as ONE flat buffer."""

    def __init__(self, nets, group=None, average=True):
        self.params = [p for n in nets for p in n.parameters() if p.requires_grad]
        self.group, self.average = group, average
        self._handle = None

    def _flat_view(self):
        """if every .grad is a view into one contiguous buffer in parameter order (which is what
        the fused backward produces), return that buffer without copying"""
        grads = [p.grad for p in self.params]
        if any(g is None for g in grads) or not grads:
            return None
        base = grads[0]
        if not all(g.is_contiguous() for g in grads):
            return None
        ptr, total = base.data_ptr(), 0
        for g in grads:
            if g.data_ptr() != ptr + 4 * total or g.dtype != torch.float32:
                return None
            total += g.numel()
        try:
            flat = base.as_strided((total,), (1,))
        except RuntimeError:
            return None
        return flat if flat.untyped_storage().nbytes() >= 4 * (total + base.storage_offset()) else None

    def allreduce(self, extra_scalars=None):
        """returns the all-reduced extra scalars (e.g. loss terms) or None"""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return extra_scalars
        world = dist.get_world_size(self.group)
        flat = self._flat_view() if extra_scalars is None else None
        if flat is not None:
            dist.all_reduce(flat, group=self.group)
            if self.average:
                flat.div_(world)
            return None
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        parts = [g.reshape(-1) for g in grads]
        if extra_scalars is not None:
            parts.append(extra_scalars.reshape(-1).to(parts[0].dtype))
        buf = torch.cat(parts)
        dist.all_reduce(buf, group=self.group)
        if self.average:
            buf.div_(world)
        off = 0
        for p, g in zip(self.params, grads):
            n = g.numel()
            if p.grad is None:
                p.grad = buf[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(buf[off:off + n].view_as(p))
            off += n
        return buf[off:] if extra_scalars is not None else None

    def install(self, optimizer):
        """reduce right before ``optimizer.step()`` (base/baseModel.py:79)"""
        return optimizer.register_step_pre_hook(lambda opt, args, kwargs: self.allreduce())


_GLOBAL_HANDLE = None
_PLATEAU_ORIG = None


def mean_over_ranks(value, group=None):
    """the rank-averaged value of a loss scalar (tensor or float) as a 0-dim tensor on the value's device"""
    t = value.detach().clone().float().reshape(1) if torch.is_tensor(value) else torch.tensor([float(value)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl" and not t.is_cuda:
            t = t.cuda()
        dist.all_reduce(t, group=group)
        t /= dist.get_world_size(group)
    return t[0]


def sync_plateau_schedulers(group=None):
    """``ReduceLROnPlateau.step(metric)`` sees the metric AVERAGED over the ranks.  The reference feeds the scheduler the
    loss of its own batch (base/baseModel.py:81) and leaves the loop when the learning rate has decayed to min_lr
    (:132-134); with sharded points every rank would otherwise cut its LR -- and stop iterating, i.e. stop entering the
    all-reduce -- at a different iteration.  One extra 4-byte all-reduce per iteration keeps the replicas identical."""
    global _PLATEAU_ORIG
    cls = torch.optim.lr_scheduler.ReduceLROnPlateau
    if _PLATEAU_ORIG is None:
        _PLATEAU_ORIG = cls.step

        def step(self, metrics, *args, **kwargs):
            return _PLATEAU_ORIG(self, mean_over_ranks(metrics, step._insr_group), *args, **kwargs)

        cls.step = step
    cls.step._insr_group = group


def install_global(nets_getter, group=None, average=True):
    """hook EVERY optimizer's step (the reference re-creates Adam per training loop,
    base/baseModel.py:55-62) and every plateau scheduler's step (``sync_plateau_schedulers``).
    ``nets_getter()`` returns the currently trainable modules."""
    global _GLOBAL_HANDLE
    from torch.optim.optimizer import register_optimizer_step_pre_hook

    def hook(opt, args, kwargs):
        GradAllReducer(nets_getter(), group=group, average=average).allreduce()

    if _GLOBAL_HANDLE is not None:
        _GLOBAL_HANDLE.remove()
    _GLOBAL_HANDLE = register_optimizer_step_pre_hook(hook)
    sync_plateau_schedulers(group)
    return _GLOBAL_HANDLE
