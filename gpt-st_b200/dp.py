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


class BucketedGradAllReduce:
    """Gradient exchange of the captured training step (``PretrainStep``): rank-SUMMED gradients stay in flat buffers.

    * ``buckets``: lists of parameters in the order their gradients complete during backward (decoder first, encoder last).
      After ``backward()`` every bucket is flattened (one ``cat`` of the existing ``.grad`` tensors) and all-reduced (sum) with
      ONE collective; nothing is scaled or copied back -- ``views()`` hands ``{id(param): view into the flat buffer}`` to the
      fused optimiser, which reads the gradients from there and folds 1/world into its clip coefficient
      (``FusedAdamClip.step(grads=...)``, ``set_grad_scale(1/world)``).  That removes the ``mul_`` and the ~135-tensor
      ``_foreach_copy_`` ``FlatGradAllReduce.reduce`` pays between backward and the optimiser.
    * Every bucket but the last runs on a communication stream that waits only for the events recorded when ITS gradients were
      accumulated (post-accumulate-grad hooks run under the producing node's stream guard).  Inside the captured graph these
      are true dependencies, so the decoder bucket's all-reduce overlaps the encoder's backward kernels over NVLink; only the
      last bucket is exposed.  With CPU tensors (gloo tests) the same code runs without streams.

    Parameters whose ``.grad`` is None (mask-phase dependent, identical on all ranks) are skipped, as in FlatGradAllReduce.
    """

    def __init__(self, buckets, group=None):
        self.buckets: List[List[torch.nn.Parameter]] = [[p for p in b if p.requires_grad] for b in buckets]
        self.buckets = [b for b in self.buckets if b]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.calls = 0
        self.last_numel = 0
        self._events = {}
        self._handles = []
        self._comm = None
        self._views = {}

    @property
    def scale(self) -> float:
        return 1.0 / self.world

    def arm(self) -> "BucketedGradAllReduce":
        """Register the per-parameter ready events of the early buckets (CUDA parameters only; idempotent)."""
        if self._handles or self.world == 1 or len(self.buckets) < 2:
            return self

        def hook(p):
            if p.is_cuda:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(p.device))
                self._events[id(p)] = ev

        for b in self.buckets[:-1]:
            for p in b:
                self._handles.append(p.register_post_accumulate_grad_hook(hook))
        return self

    def disarm(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
        self._events = {}

    def _flatten_reduce(self, live):
        flat = torch.cat([p.grad.reshape(-1) for p in live])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for p in live:
            n = p.numel()
            self._views[id(p)] = flat[off:off + n].view_as(p)
            off += n
        self.last_numel += flat.numel()
        return flat

    def reduce(self):
        """Call after backward().  Returns {id(param): rank-summed gradient view}; multiply by ``scale`` for the average."""
        self._views = {}
        self.last_numel = 0
        if self.world == 1:
            return self._views
        nb = len(self.buckets)
        for i, b in enumerate(self.buckets):
            live = [p for p in b if p.grad is not None]
            if not live:
                continue
            early = i < nb - 1 and live[0].is_cuda and all(id(p) in self._events for p in live)
            if early:
                dev = live[0].device
                cur = torch.cuda.current_stream(dev)
                if self._comm is None or self._comm.device != dev:
                    self._comm = torch.cuda.Stream(device=dev, priority=-1)
                comm = self._comm
                for p in live:
                    comm.wait_event(self._events[id(p)])
                with torch.cuda.stream(comm):
                    flat = self._flatten_reduce(live)
                for p in live:
                    p.grad.record_stream(comm)
                flat.record_stream(cur)
                done = torch.cuda.Event()
                done.record(comm)
                cur.wait_event(done)        # (a join, not a stall: the captured graph only gains an edge to the optimiser)
            else:
                self._flatten_reduce(live)
        self._events = {}
        self.calls += 1
        return self._views

    def views(self):
        return self._views


def pretrain_buckets(model) -> List[List[torch.nn.Parameter]]:
    """Buckets of a GPTST_Model for BucketedGradAllReduce, in the order backward finishes them: the decoder's parameters
    (their gradients are complete when the encoder's backward starts), then everything else (encoder, mask scorer)."""
    dec = [p for p in model.decoder.parameters() if p.requires_grad]
    seen = {id(p) for p in dec}
    rest = [p for p in model.parameters() if p.requires_grad and id(p) not in seen]
    return [dec, rest]


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
