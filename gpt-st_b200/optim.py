"""Fused global-norm clipping + Adam (csrc/optim.cu): two kernel launches for the whole parameter set.

Drop-in for the reference's ``clip_grad_norm_(model.parameters(), max_grad_norm); optimizer.step()``
(BasicTrainer.py:94-97, Adam created at Run.py:134).  Parameters whose ``.grad`` is None are skipped exactly as
torch.optim.Adam skips them; every parameter has its own step counter ON THE DEVICE (bumped by the kernel), so a parameter that
gets its first gradient after any number of CUDA-graph replays starts at t = 1 like torch.optim.Adam's per-parameter state.  CUDA-graph safe: the pointer table
travels through a pinned host buffer (a memcpy node that re-reads the same, still valid, addresses at replay).
"""
from __future__ import annotations

import os
from typing import Iterable, Optional

import torch

from . import _lib, ops


class FusedAdamClip:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 3e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: Optional[float] = 5.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamClip needs CUDA parameters (no CPU fallback)")
        self.device = dev
        self.state = {}                                    # param -> (exp_avg, exp_avg_sq, int32 device step counter)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        # last entry: gradient scale (1/world when the step hands over rank-summed gradients, see step(grads=...))
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, max_grad_norm if max_grad_norm else 0.0, 1.0],
                                  dtype=torch.float32, device=dev)
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._tables = {}                                  # key -> [pinned table, dev table, dev block map, partial, nblocks, pinned map]
        self._captures = 0
        self.chunk = _lib.lib().gptst_opt_chunk()
        self._pre = None                                   # buffers of prefetch_tables() waiting for step()
        self._grads = None
        self._reserved = None                              # pinned buffers allocated by reserve() for the next capture

    def set_lr(self, lr: float) -> None:
        self.hyper[0:1].fill_(lr)

    def set_grad_scale(self, scale: float) -> None:
        """Every gradient is multiplied by `scale` before the norm and the update (1/world for rank-summed gradients)."""
        self.hyper[5:6].fill_(float(scale))

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def reserve(self) -> None:
        """Allocate the pinned staging buffers of the NEXT captured step (call right before the capture begins).  Pinning
        memory inside a capture makes torch's host allocator poll the events of earlier pinned blocks (cudaEventQuery), which
        is illegal while a global-mode capture is under way and invalidates it -- depending on what earlier code left in the
        host cache (seen as a capture failure in the middle of the GPU test suite)."""
        nblocks = sum((p.numel() + self.chunk - 1) // self.chunk for p in self.params)
        self._reserved = (torch.zeros((len(self.params), 6), dtype=torch.int64).pin_memory(),
                          torch.zeros((nblocks, 2), dtype=torch.int32).pin_memory())

    def _table(self, live):
        """(Re)build the pointer table for the current gradient tensors.  Inside a CUDA-graph capture this runs once:
        the addresses are static across replays."""
        rows, bmap = [], []
        for idx, p in enumerate(live):
            if p not in self.state:
                self.state[p] = (torch.zeros_like(p, memory_format=torch.contiguous_format),
                                 torch.zeros_like(p, memory_format=torch.contiguous_format),
                                 torch.zeros(1, dtype=torch.int32, device=p.device))
            m, v, tcount = self.state[p]
            g = self._grad_of(p)
            if not (p.is_contiguous() and g.is_contiguous()) or g.dtype != torch.float32:
                raise RuntimeError("FusedAdamClip: fp32 contiguous parameters / gradients only")
            n = p.numel()
            rows.append([p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, tcount.data_ptr()])
            bmap += [[idx, c] for c in range((n + self.chunk - 1) // self.chunk)]
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            self._captures += 1
        # one entry per LIVE SET (the block map depends on which tensors are live, not only on how many); a captured graph owns
        # its table (never rewritten later)
        key = (tuple(id(p) for p in live), self._captures if capturing else 0)
        ent = self._tables.get(key)
        if ent is None:
            if capturing and self._reserved is not None:
                pinned, bpin = self._reserved[0][:len(rows)], self._reserved[1][:len(bmap)]
                bpin.copy_(torch.tensor(bmap, dtype=torch.int32))
                self._reserved = None                               # the captured graph owns them from here on
            else:
                pinned = torch.empty((len(rows), 6), dtype=torch.int64).pin_memory()
                bpin = torch.tensor(bmap, dtype=torch.int32).pin_memory()
            bdev = torch.empty((len(bmap), 2), dtype=torch.int32, device=self.device)
            bdev.copy_(bpin, non_blocking=True)                     # pinned -> device: legal inside a capture
            ent = [pinned, torch.empty((len(rows), 6), dtype=torch.int64, device=self.device), bdev,
                   torch.empty(len(bmap), dtype=torch.float32, device=self.device), len(bmap), bpin, None]
            self._tables[key] = ent
        elif ent[6] is not None:
            ent[6].synchronize()                                    # eager mode: the previous step's H2D copy has read the pinned rows
        ent[0].copy_(torch.tensor(rows, dtype=torch.int64))
        ent[1].copy_(ent[0], non_blocking=True)
        if not capturing:
            ent[6] = torch.cuda.Event()
            ent[6].record()
        return ent

    def prefetch_tables(self) -> None:
        """Call at the START of a step that is being captured into a CUDA graph (opt-in, GPTST_B200_OPT_PREFETCH=1): issues the
        two pinned -> device table copies on a forked stream right away, so that their memcpy nodes hang off the root of the
        graph instead of sitting between the last gradient kernel and the optimiser.  The pinned buffers are filled by
        `step()` later in the same capture -- a memcpy node reads its source when the graph is replayed, not when it is
        captured.  Outside a capture (or when disabled) this is a no-op and `step()` builds the tables as before."""
        self._pre = None
        if os.environ.get("GPTST_B200_OPT_PREFETCH", "0") != "1" or not torch.cuda.is_current_stream_capturing():
            return
        nblocks = sum((p.numel() + self.chunk - 1) // self.chunk for p in self.params)
        if self._reserved is not None:
            pinned, bpin = self._reserved
            self._reserved = None
        else:
            pinned = torch.zeros((len(self.params), 6), dtype=torch.int64).pin_memory()
            bpin = torch.zeros((nblocks, 2), dtype=torch.int32).pin_memory()
        tdev = torch.empty((len(self.params), 6), dtype=torch.int64, device=self.device)
        bdev = torch.empty((nblocks, 2), dtype=torch.int32, device=self.device)
        part = torch.empty(nblocks, dtype=torch.float32, device=self.device)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            tdev.copy_(pinned, non_blocking=True)
            bdev.copy_(bpin, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
        self._pre = (pinned, bpin, tdev, bdev, part, ev, side)

    def _table_prefetched(self, live):
        """Fill the pinned buffers of `prefetch_tables` for the gradients that exist now and join its stream."""
        pinned, bpin, tdev, bdev, part, ev, _side = self._pre
        rows, bmap = [], []
        for idx, p in enumerate(live):
            if p not in self.state:
                self.state[p] = (torch.zeros_like(p, memory_format=torch.contiguous_format),
                                 torch.zeros_like(p, memory_format=torch.contiguous_format),
                                 torch.zeros(1, dtype=torch.int32, device=p.device))
            m, v, tcount = self.state[p]
            g = self._grad_of(p)
            if not (p.is_contiguous() and g.is_contiguous()) or g.dtype != torch.float32:
                raise RuntimeError("FusedAdamClip: fp32 contiguous parameters / gradients only")
            n = p.numel()
            rows.append([p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, tcount.data_ptr()])
            bmap += [[idx, c] for c in range((n + self.chunk - 1) // self.chunk)]
        pinned[:len(rows)].copy_(torch.tensor(rows, dtype=torch.int64))          # host writes: the captured copies read them at replay
        bpin[:len(bmap)].copy_(torch.tensor(bmap, dtype=torch.int32))
        torch.cuda.current_stream().wait_event(ev)                                # join the forked branch
        self._captures += 1
        self._tables[("prefetched", self._captures)] = self._pre                 # the graph owns these buffers
        self._pre = None
        return [pinned, tdev, bdev, part, len(bmap), bpin, None]

    def _grad_of(self, p):
        g = self._grads.get(id(p)) if self._grads else None
        return p.grad if g is None else g

    def step(self, grads=None) -> None:
        """grads: optional {id(param): tensor} -- gradients to read INSTEAD of ``param.grad`` (views into the flat all-reduce
        buffers of data-parallel training, so nothing is copied back into ``.grad``); which parameters are live is still
        decided by ``param.grad is not None``."""
        self._grads = grads
        live = [p for p in self.params if p.grad is not None]
        pre = getattr(self, "_pre", None)
        if not live:
            if pre is not None:
                torch.cuda.current_stream().wait_event(pre[5])
                self._pre = None
            return
        ent = self._table_prefetched(live) if pre is not None else self._table(live)
        L = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        ops._count(2)
        _lib.check(L.gptst_adam_clip(ent[1].data_ptr(), ent[2].data_ptr(), ent[4], ent[3].data_ptr(), self.step_count.data_ptr(),
                                     self.hyper.data_ptr(), self.norm.data_ptr(), st), "gptst_adam_clip")
