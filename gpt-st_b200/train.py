"""B200-native pre-training step: forward, loss, backward, gradient all-reduce, clipping and Adam as ONE CUDA graph.

The reference trainer (BasicTrainer.py:72-103) issues >3 600 eager ATen ops and 2-3 `.item()` syncs per step and is
launch-bound on any modern GPU (27 ms/step on a B200 for PEMS08 / batch 64, measured).  Here the whole step is
stream-ordered and sync-free (device-side mask planning, sync-free loss, capturable Adam), so it is captured once per
mask phase and replayed: per step the host only refreshes a 12-element mask plan, copies the batch from pinned memory
and launches one graph.

    step = PretrainStep(model, lr=3e-3, max_grad_norm=5, loss="probe")
    loss = step(source_host_or_device, epoch)          # device scalar; .item() when the value is needed

`GPTST_Model` stays usable eagerly (reference Run.py / BasicTrainer unchanged) -- this module is the fast path for
callers that own their training loop.
"""
from __future__ import annotations

import os
import random
from typing import Callable, Optional, Union

import torch

from . import dp as _dp
from .losses import pretrain_loss_syncfree, probe_loss


class PretrainStep:
    def __init__(self, model, lr: float = 3e-3, max_grad_norm: Optional[float] = 5.0, loss: Union[str, Callable] = "probe",
                 scaler_mean: float = 0.0, scaler_std: float = 1.0, mask_value: float = 0.0, use_graph: bool = True, reducer: Optional[_dp.FlatGradAllReduce] = None,
                 optimizer: Optional[torch.optim.Optimizer] = None, fused_optimizer: bool = True):
        self.model = model
        self.enc = model.encoder
        self.max_grad_norm = max_grad_norm
        self.use_graph = use_graph
        self.reducer = reducer
        self.params = [p for p in model.parameters() if p.requires_grad]
        # graph mode: fused clip+Adam (two launches); eager mode / user-supplied optimiser: torch's foreach path
        self.fused_opt = None
        if optimizer is None and use_graph and fused_optimizer:
            from .optim import FusedAdamClip
            self.fused_opt = FusedAdamClip(self.params, lr=lr, eps=1e-8, max_grad_norm=max_grad_norm)
        if isinstance(reducer, _dp.BucketedGradAllReduce):
            if self.fused_opt is None:
                raise ValueError("BucketedGradAllReduce leaves the rank-summed gradients in flat buffers: it needs the fused "
                                 "optimiser (use FlatGradAllReduce with torch optimisers)")
            self.fused_opt.set_grad_scale(reducer.scale)
            reducer.arm()
        self.opt = optimizer or (None if self.fused_opt else torch.optim.Adam(self.params, lr=lr, eps=1e-8, capturable=use_graph,
                                                                              foreach=True))
        if isinstance(loss, str):
            from . import ops as _ops
            if loss == "probe":
                self.loss_fn = lambda outs, src, ep: _ops.fused_probe_loss(outs, src, ep > self.enc.change_epoch)
            elif loss == "mask_mae":
                self.loss_fn = lambda outs, src, ep: _ops.fused_mask_mae_loss(outs, src, ep > self.enc.change_epoch,
                                                                              scaler_mean, scaler_std, mask_value)
            elif loss == "probe_torch":
                self.loss_fn = lambda outs, src, ep: probe_loss(outs, src, ep, self.enc.change_epoch)
            elif loss == "mask_mae_torch":
                self.loss_fn = lambda outs, src, ep: pretrain_loss_syncfree(outs, src, ep, self.enc.change_epoch, scaler_mean,
                                                                            scaler_std, model.output_dim, mask_value)
            else:
                raise ValueError(loss)
        else:
            self.loss_fn = loss
        self._graphs = {}        # phase -> (graph, static_src, static_loss, plan_dev)
        self._warm = {}          # phase -> eager warm-up steps done
        self.replays = 0
        self.launches_per_step = 0
        self._pf = None              # (host tensor, staged device copy, ready event) of the batch announced by `prefetch=`
        self._copy_stream = None
        self._stage_bufs = {}

    # -- one eager step (also the body that gets captured) ------------------------------------------------
    def _zero_grad(self):
        for p in self.params:
            p.grad = None

    def _body(self, src, epoch):
        self._zero_grad()
        if self.fused_opt is not None:
            self.fused_opt.prefetch_tables()           # no-op unless capturing with GPTST_B200_OPT_PREFETCH=1
        outs = self.model(src, src, None, epoch)
        loss = self.loss_fn(outs, src, epoch)
        loss.backward()
        gviews = None
        if self.reducer is not None:
            gviews = self.reducer.reduce()             # Bucketed...: {id(param): rank-summed view}; Flat...: None (in place)
        if self.fused_opt is not None:
            self.fused_opt.step(gviews or None)        # clip + Adam, two launches (1/world folded into the clip coefficient)
        else:
            if self.max_grad_norm is not None:
                torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm, foreach=True)
            self.opt.step()
        return loss.detach()

    def _phase(self, epoch):
        return 2 if epoch > self.enc.change_epoch else 1

    def _capture(self, src, epoch, phase):
        dev = src.device
        static_src = torch.empty_like(src)
        static_src.copy_(src)
        plan_dev = None
        if phase == 2:
            n = src.shape[0] * src.shape[1] * src.shape[2]
            plan_dev = self.enc.mask_plan(n, epoch).to(dev)
            self.enc.plan_override = plan_dev
        dot = os.environ.get("GPTST_B200_GRAPH_DOT")      # debugging aid: dump the captured graph (nodes + dependency edges)
        g = torch.cuda.CUDAGraph(keep_graph=True) if dot else torch.cuda.CUDAGraph()
        self._zero_grad()
        if self.fused_opt is not None:
            self.fused_opt.reserve()               # pinned staging buffers: never allocate pinned memory inside the capture
        torch.cuda.synchronize()
        from . import ops as _ops
        l0 = _ops.launch_count()
        # capture on a HIGH-priority stream: the main chain's kernel nodes inherit it, while the table prologues and their
        # gradients sit on default (lowest) priority side streams and only fill the SMs the main chain leaves idle
        cap_stream = torch.cuda.Stream(device=dev, priority=-1)
        # No cyclic garbage collection while the capture is under way: collecting an older PretrainStep (its CUDA graph and
        # private memory pool: cudaGraphExecDestroy / cudaFree) from inside a global-mode capture invalidates it -- a
        # timing-dependent failure seen in the middle of the GPU test suite.  (torch.cuda.graph collects once on entry.)
        import gc
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(g, stream=cap_stream):
                static_loss = self._body(static_src, epoch)
        finally:
            if gc_was_on:
                gc.enable()
            self.enc.plan_override = None
        self.launches_per_step = _ops.launch_count() - l0   # libgptst_b200 kernels inside one replay
        if dot:
            g.debug_dump(f"{dot}.phase{phase}.dot")
            g.instantiate()
        self._graphs[(phase, tuple(src.shape))] = (g, static_src, static_loss, plan_dev)

    def _stage(self, nxt, consumed):
        """Start the host -> device copy of the NEXT step's batch on a copy stream, under the graph that was just launched."""
        if nxt is None or nxt.is_cuda:
            self._pf = None
            return
        dev = next(self.model.parameters()).device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        stage = self._stage_bufs.get(tuple(nxt.shape))
        if stage is None:
            stage = self._stage_bufs[tuple(nxt.shape)] = torch.empty(nxt.shape, dtype=nxt.dtype, device=dev)
        cs = self._copy_stream
        cs.wait_event(consumed)             # the staged batch of THIS step has been copied out (NOT a wait for the graph)
        with torch.cuda.stream(cs):
            stage.copy_(nxt, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._pf = (nxt, stage, ev)

    def __call__(self, source: torch.Tensor, epoch: int, prefetch: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training step on `source` (host or device tensor).  prefetch = the NEXT step's host batch (pinned): its H2D copy
        is issued behind this step's graph launch on a copy stream and overlaps the step; the next call (with that same tensor
        object as `source`) then starts from the staged device copy instead of waiting for PCIe."""
        dev = next(self.model.parameters()).device
        if not self.use_graph:
            if not source.is_cuda:
                source = source.to(dev, non_blocking=True)
            return self._body(source, epoch)
        phase = self._phase(epoch)
        key = (phase, tuple(source.shape))
        if key not in self._graphs:
            # a few eager steps first: allocator warm-up, optimizer state creation (params that only get gradients in
            # this phase), cuBLAS workspaces -- required before capture
            done = self._warm.get(key, 0)
            src = source.to(dev, non_blocking=True) if not source.is_cuda else source
            if done < 3:
                self._warm[key] = done + 1
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    loss = self._body(src, epoch)
                torch.cuda.current_stream().wait_stream(side)
                return loss
            self._capture(src, epoch, phase)
            fresh = True                     # the capture drew this step's class order already (one shuffle per step, ref :357-358)
        else:
            fresh = False
        g, static_src, static_loss, plan_dev = self._graphs[key]
        if plan_dev is not None and not fresh:
            n = source.shape[0] * source.shape[1] * source.shape[2]
            plan_dev.copy_(self.enc.mask_plan(n, epoch), non_blocking=True)
        pf, self._pf = self._pf, None
        cur = torch.cuda.current_stream()
        if pf is not None and pf[0] is source:
            cur.wait_event(pf[2])
            source = pf[1]                                      # staged by the previous call: device -> device
        if source.is_cuda:
            # a copy KERNEL, not cudaMemcpyAsync: the copy engine takes 17 us for these 1.5 MB, in front of the whole graph
            torch._foreach_copy_([static_src], [source])
        else:
            static_src.copy_(source, non_blocking=True)
        consumed = None
        if prefetch is not None:
            consumed = torch.cuda.Event()
            consumed.record(cur)
        g.replay()
        self.replays += 1
        if prefetch is not None:
            self._stage(prefetch, consumed)
        return static_loss
