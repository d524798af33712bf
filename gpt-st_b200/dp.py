"""Data-parallel pre-training: shard the batch over ranks, one all-reduce of a flat gradient buffer per step.

The reference has no distributed code at all; this is the
B200-native addition: one process per GPU, NCCL over NVLink 5 / NVSwitch, a single fp32 sum all-reduce of
every gradient that exists after backward (4.1 MB for PEMS08), issued from an end-of-backward callback so
that unchanged trainer code (``loss.backward(); clip_grad_norm_; optimizer.step()`` -- reference
BasicTrainer.py:92-97) sees averaged gradients.  Parameters whose ``.grad`` is ``None`` in the current phase
(39 tensors in the random-mask phase, 20 afterwards; identical on all ranks because ``epoch`` is) stay ``None``
so Adam skips them exactly as it does single-GPU.

No data-path collective exists: no op of the model mixes samples (every contraction is per b).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Initialise torch.distributed from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world).  Single-process when WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank r takes rows r::world of the global batch."""
    return t if world == 1 else t[rank::world]


class FlatGradAllReduce:
    """Averages gradients across ranks with ONE collective per backward pass."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._queued = False
        self._handles = []
        self.calls = 0
        self.last_numel = 0

    # -- explicit use -----------------------------------------------------------------------------------
    def reduce(self) -> None:
        """Flatten every existing .grad, all-reduce(sum), scale by 1/world, scatter back in place."""
        if self.world == 1:
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.mul_(1.0 / self.world)
        off = 0
        views = []
        for g in grads:
            n = g.numel()
            views.append(flat[off:off + n].view_as(g))
            off += n
        torch._foreach_copy_(grads, views)
        self.calls += 1
        self.last_numel = flat.numel()

    # -- transparent use (unchanged trainer code) ---------------------------------------------------------
    def attach(self) -> "FlatGradAllReduce":
        """Hook the end of every backward pass: the first gradient produced queues a callback on the autograd
        engine that runs after the whole graph has been executed."""
        if self.world == 1:
            return self

        def hook(_p):
            if not self._queued:
                self._queued = True
                torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

        for p in self.params:
            self._handles.append(p.register_post_accumulate_grad_hook(hook))
        return self

    def _finalize(self):
        self._queued = False
        self.reduce()

    def detach(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers (one flat broadcast)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers() if b.is_floating_point()]
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
