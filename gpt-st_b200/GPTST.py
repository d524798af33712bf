"""Drop-in replacement for the reference ``model/Pretrain_model/GPTST.py`` (class ``GPTST_Model``).

Same constructor ``args`` reads, same parameter names / shapes / registration order (so the shipped
``GPTST_ada.pth`` checkpoints load with ``strict=True`` and ``Run.py``'s init loop and Adam see the
parameters in the same order), same ``forward(source, label, batch_seen=None, epoch=None)`` 5-tuple
(reference GPTST.py:459-493).  The three heavy blocks run hand-written sm_100a kernels through
``gptst_b200.ops``:

    hyperTem  (ref :144-163)   -> ops.hypertem_core   (temporal hypergraph two-hop + time-adaptive projection)
    cap       (ref :79-141)    -> ops.cap_core        (routing, inter-cluster hop, node-adaptive GCN)
    MLP_RL    (ref :6-34)      -> ops.node_adaptive_proj / ops.time_adaptive_proj

Everything (B,T,N,*)-sized runs in libgptst_b200 kernels: the fused hyperTem blocks (csrc/htem_fused.cu), the cap chain, the
scorer, the time-embedding MLPs and low-rank table generators (csrc/small_ops.cu), the mask construction (csrc/mask.cu: same
``rand_like`` / ``random.shuffle`` draws as the reference, so masks are bit-identical for equal seeds on the same device), the
input / output projections (the masked fill of the encoder input is part of its projection kernel) and the loss.  PyTorch keeps
the module tree, autograd bookkeeping, streams and a few elementwise glue ops ((1 - mask), gradient accumulation of shared leaves).

The model never touches ``'cuda:0'`` literally: it follows the device of its inputs, so one process per
GPU (LOCAL_RANK) works for data parallel training.
"""
from __future__ import annotations

import os
import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def _affine(lin, x):
    """nn.Linear with one input feature as a single broadcast multiply-add (a K = 1 GEMM plus a bias kernel otherwise)."""
    if lin.in_features == 1 and x.is_cuda:
        return ops.affine1(x, lin.weight, lin.bias)
    return lin(x)


def _uninit(*shape):
    # the reference registers uninitialised torch.FloatTensor pools (ref :13-17,92-93,149-150); Run.py:79-85 fills them
    return nn.Parameter(torch.empty(*shape))


class MLP_RL(nn.Module):
    """Adaptive-mask scorer: node-adaptive then time-adaptive MLP (ref :6-34)."""

    def __init__(self, dim_in, dim_out, hidden_dim, embed_dim, device):
        super().__init__()
        self.ln1 = nn.Linear(dim_in, hidden_dim)
        self.ln3 = nn.Linear(hidden_dim, dim_out)
        self.weights_pool_spa = _uninit(embed_dim, hidden_dim, hidden_dim)
        self.bias_pool_spa = _uninit(embed_dim, hidden_dim)
        self.weights_pool_tem = _uninit(embed_dim, hidden_dim, hidden_dim)
        self.bias_pool_tem = _uninit(embed_dim, hidden_dim)
        self.device = device

    def tables_spa(self, node_eb):
        return ops.lowrank_table(node_eb, self.weights_pool_spa), ops.lowrank_table(node_eb, self.bias_pool_spa)

    def tables_tem(self, time_eb):
        return ops.lowrank_table(time_eb, self.weights_pool_tem), ops.lowrank_table(time_eb, self.bias_pool_tem)

    def forward(self, eb, time_eb, node_eb, tables=None, probs=False):
        """tables = ((Wn, bn, event), (Wt, bt, event)) when the caller produced them on side streams.  probs=True returns
        softmax(logits) through the fused score head instead of the logits."""
        h0 = _affine(self.ln1, eb)                                                # (B,T,N,D)
        if tables is None:
            Wn, bn = self.tables_spa(node_eb)
            h1 = ops.node_adaptive_proj(h0, Wn, bn)
            Wt, bt = self.tables_tem(time_eb)
            h2 = ops.time_adaptive_proj(h1, Wt, bt)
        else:
            (Wn, bn, ev_n), (Wt, bt, ev_t) = tables
            torch.cuda.current_stream().wait_event(ev_n)
            h1 = ops.node_adaptive_proj(h0, Wn, bn)
            torch.cuda.current_stream().wait_event(ev_t)
            h2 = ops.time_adaptive_proj(h1, Wt, bt)
        if probs:
            return ops.score_head(h2, self.ln3.weight, self.ln3.bias)
        return self.ln3(h2)


class cap(nn.Module):
    """Hierarchical intra/inter-cluster hypergraph propagation + node-adaptive GCN (ref :79-141)."""

    def __init__(self, dim, num_nodes, timesteps, embed_dim, embed_dim_spa, HS, HT, num_route):
        super().__init__()
        self.num_nodes, self.timesteps, self.dim = num_nodes, timesteps, dim
        self.num_route, self.HS, self.TT = num_route, HS, HS * timesteps
        self.ln_p = nn.Linear(dim, dim)
        self.t_adj = nn.Parameter(torch.randn(embed_dim_spa, HT, self.TT))
        self.adj = nn.Parameter(torch.randn(embed_dim_spa, HS, num_nodes))
        self.weights_spa = _uninit(embed_dim, dim, dim)
        self.bias_spa = _uninit(embed_dim, dim)
        # kept for state_dict compatibility; the kernels use tau_t = (t+1)/12 directly (ref :97,125)
        self.register_buffer("mask_template", torch.linspace(1, timesteps, steps=timesteps) / 12.0)

    def tables(self, node_embeddings, time_eb, teb):
        """Parameter-side contractions (independent of x): incidence logits, inter-cluster adjacency, node-adaptive weights."""
        if self.timesteps != 12:
            raise RuntimeError("cap: the reference (and the kernels) hard-wire 12 time steps")
        dadj = ops.lowrank_table(teb, self.adj)                         # einsum("btd,dhn->bthn")
        dyn = ops.lowrank_table(time_eb, self.t_adj)                    # einsum("bd,dhk->bhk")
        Wn = ops.lowrank_table(node_embeddings, self.weights_spa)       # einsum("nd,dio->nio")
        bn = ops.lowrank_table(node_embeddings, self.bias_spa)
        if not dadj.is_cuda:
            return dadj, dyn, Wn, bn, None, None
        # stride-0 (P, ...) views of the five parameter-side inputs: the block's backward returns its per-CTA gradient partials
        # for them and the sums run HERE (this stream) as the expands' backward, off the main chain (ops.expand_partials)
        B, T, H, N = dadj.shape
        exp = ops.cap_expand(self.ln_p.weight, self.ln_p.bias, dyn, Wn, bn, B, T, N, self.dim, H)
        # fragment-ordered fp16 copy of W_n for the fused reconstruction + projection kernel, packed here (off the main chain)
        wnf = ops.cap_pack_wn(Wn) if ops.cap_fused_enabled(N, self.dim, H, T, ops.default_precision()) else None
        return dadj, dyn, Wn, bn, exp, wnf

    def forward(self, x, node_embeddings, time_eb, teb, tables=None):
        dadj, dyn, Wn, bn, exp, wnf = tables if tables is not None else self.tables(node_embeddings, time_eb, teb)
        if exp is None:
            out, c = ops.cap_core(x, self.ln_p.weight, self.ln_p.bias, dadj, dyn, Wn, bn, self.num_route)
        else:
            out, c = ops.cap_core(x, exp[0], exp[1], dadj, exp[2], exp[3], exp[4], self.num_route, expanded=True, wnf=wnf)
        return out, c.unsqueeze(-1), dyn.detach()


class hyperTem(nn.Module):
    """Temporal hypergraph block with time-adaptive weights (ref :144-163)."""

    def __init__(self, timesteps, num_node, dim_in, dim_out, embed_dim, HT_Tem):
        super().__init__()
        self.c_out = dim_out
        self.adj = nn.Parameter(torch.randn(embed_dim, HT_Tem, timesteps))
        self.weights_pool = _uninit(embed_dim, dim_in, dim_out)
        self.bias_pool = _uninit(embed_dim, dim_out)

    def tables(self, node_embeddings, time_eb, aux=None):
        """Parameter-side contractions (independent of eb): per-node T x T mix and time-adaptive weights.  aux = a second side
        stream for the mix-matrix branch (A_n, M_n): in backward its gradient chain (dM_n -> dA_n -> d adj, dE) then runs beside
        the weight branch (dW_bt -> d pool, d time_eb) instead of behind it."""
        fused = (node_embeddings.is_cuda and ops.hypertem_fused_enabled(self.weights_pool.shape[-1], self.adj.shape[-1],
                                                                        ops.default_precision()))
        mb = {} if fused else None
        cur = torch.cuda.current_stream() if node_embeddings.is_cuda else None

        def mix_branch():
            A = ops.lowrank_table(node_embeddings, self.adj)            # einsum("nk,kht->nht")
            Mn = ops.mix_matrix(A) if A.is_cuda else torch.einsum("nht,nhs->nts", A, A)   # two hops, no nonlinearity in between
            return ops.hypertem_params_m(mb, Mn) if fused else Mn

        if fused and aux is not None:
            fork = torch.cuda.Event()
            fork.record(cur)
            aux.wait_event(fork)
            with torch.cuda.stream(aux):
                Mn = mix_branch()
                ev_m = torch.cuda.Event()
                ev_m.record(aux)
            Mn.record_stream(cur)
        else:
            Mn, ev_m = mix_branch(), None
        W = ops.lowrank_table(time_eb, self.weights_pool)               # einsum("btd,dio->btio")
        bias = ops.lowrank_table(time_eb, self.bias_pool)
        if fused:
            # fused block: the parameter-side nodes live on THESE streams, so dM_n / dW_bt / db_bt are computed here in backward
            W, bias = ops.hypertem_params_w(mb, W, bias)
            if ev_m is not None:
                cur.wait_event(ev_m)       # forward join: the caller's single event covers both branches
            return Mn, W, bias, mb["wf"], mb["wb"], mb
        if Mn.is_cuda:
            # (P, N, T, T) stride-0 view: the dM_n partials of the backward are summed on this stream (ops.expand_partials)
            Mn = ops.expand_partials(Mn, ops.hypertem_partial_count(time_eb.shape[0], Mn.shape[0], W.shape[-1]))
        return Mn, W, bias

    def forward(self, eb, node_embeddings, time_eb, tables=None):
        tb = tables if tables is not None else self.tables(node_embeddings, time_eb)
        if len(tb) == 6:
            return ops.hypertem_fused(eb, tb[0], tb[1], tb[2], tb[5])
        Mn, W, bias = tb
        return ops.hypertem_core(eb, Mn, W, bias)


class time_feature(nn.Module):
    def __init__(self, embed_dim, first=1):
        super().__init__()
        self.ln_day = nn.Linear(first, embed_dim)
        self.ln_week = nn.Linear(first, embed_dim)
        self.ln1 = nn.Linear(embed_dim, embed_dim)
        self.ln2 = nn.Linear(embed_dim, embed_dim)
        self.ln = nn.Linear(embed_dim, embed_dim)

    def forward(self, eb):
        if eb.is_cuda and self.ln1.in_features <= 16:
            B, T = eb.shape[0], eb.shape[1]
            ab = eb.permute(2, 0, 1).reshape(2, B * T, 1)                 # one fused kernel each way instead of ~12 / ~25
            return ops.time_mlp(ab, self).view(B, T, -1)
        h = self.ln_day(eb[:, :, 0:1]) + self.ln_week(eb[:, :, 1:2])
        return self.ln(F.relu(self.ln2(F.relu(self.ln1(h)))))


class time_feature_spg(time_feature):
    """Same MLP but the first layer maps the T (=12) axis (ref :204-219)."""

    def __init__(self, embed_dim):
        super().__init__(embed_dim, first=12)

    def forward(self, eb):
        if eb.is_cuda and self.ln1.in_features <= 16 and eb.shape[1] <= 12:
            return ops.time_mlp(eb.permute(2, 0, 1), self)                 # (2, B, 12) -> (B, ds)
        h = self.ln_day(eb[:, :, 0]) + self.ln_week(eb[:, :, 1])
        return self.ln(F.relu(self.ln2(F.relu(self.ln1(h)))))


class STHCN(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.num_node, self.input_base_dim = args.num_nodes, args.input_base_dim
        self.hidden_dim, self.embed_dim, self.embed_dim_spa = args.hidden_dim, args.embed_dim, args.embed_dim_spa
        D, d, ds = self.hidden_dim, self.embed_dim, self.embed_dim_spa
        self.node_embeddings = nn.Parameter(torch.randn(self.num_node, d))
        self.node_embeddings_spg = nn.Parameter(torch.randn(self.num_node, d))
        for i in range(1, 5):
            setattr(self, f"hyperTem{i}", hyperTem(args.horizon, args.num_nodes, D, D, d, args.HT_Tem))
        self.time_feature1 = time_feature(d)
        self.time_feature1_ = time_feature(ds)
        self.time_feature2 = time_feature_spg(ds)
        for i in (1, 2):
            setattr(self, f"cap{i}", cap(D, args.num_nodes, args.horizon, d, ds, args.HS, args.HT, args.num_route))

    def prologue(self, source):
        """Everything of this STHCN that does not depend on the activations: the three time embeddings (ref :256-261)
        and the adaptive tables of the six blocks.  GPTST_Model runs it on a side stream."""
        i0 = self.input_base_dim
        tf_in = source[:, :, 0, i0:i0 + 2]                             # day / week index of node 0 (ref :256-257)
        time_eb = self.time_feature1(tf_in)
        teb = self.time_feature1_(tf_in)
        time_eb_spg = self.time_feature2(tf_in)
        E, Es = self.node_embeddings, self.node_embeddings_spg
        return {"ht": [getattr(self, f"hyperTem{i}").tables(E, time_eb) for i in range(1, 5)],
                "cap": [getattr(self, f"cap{i}").tables(Es, time_eb_spg, teb) for i in (1, 2)]}

    def prologue_streams(self, source, pool, fork, main):
        """`prologue` spread over side streams (three time embeddings, then one stream per block and a second one per hyperTem
        for its mix-matrix branch, two for the gradient accumulation of the shared node embeddings): inside a captured
        CUDA graph each stream is a linear dependency chain, and autograd replays the same streams in backward, so with
        one stream per block the table gradients of a block only wait for that block's own backward kernel instead of
        queueing behind every other block's (they used to form a 570 us serial tail after the last main kernel)."""
        i0 = self.input_base_dim
        tf_in = source[:, :, 0, i0:i0 + 2]
        embs, evs = [], []
        for st, mod in zip(pool[0:3], (self.time_feature1, self.time_feature1_, self.time_feature2)):
            st.wait_event(fork)
            with torch.cuda.stream(st):
                e = mod(tf_in)
                ev = torch.cuda.Event()
                ev.record(st)
            embs.append(e)
            evs.append(ev)
        time_eb, teb, time_eb_spg = embs
        # The node embeddings are shared by the four hyperTem (two cap) blocks.  Autograd accumulates a leaf's gradient on the
        # stream of its FIRST use -- hyperTem1's (cap1's) table stream, the block whose backward runs LAST: its parameter
        # gradients then queue behind the additions of the other blocks' contributions, i.e. behind their whole table chains
        # (measured: the last block's side chain started 100 us after its inputs were ready).  A view made on a dedicated
        # stream moves the accumulation (and the leaf's AccumulateGrad node) there.
        E, Es = self.node_embeddings, self.node_embeddings_spg
        if len(pool) >= 15:
            for k, name in ((13, "E"), (14, "Es")):
                pool[k].wait_event(fork)
                with torch.cuda.stream(pool[k]):
                    if name == "E":
                        E = E.view_as(E)
                    else:
                        Es = Es.view_as(Es)
                    evj = torch.cuda.Event()
                    evj.record(pool[k])
                main.wait_event(evj)                      # (no kernel ran there; a forked capture stream must be joined)
        pro = {"ht": [], "cap": [], "ht_ev": [], "cap_ev": []}
        for i in range(4):
            st = pool[3 + i]
            st.wait_event(evs[0])
            time_eb.record_stream(st)
            with torch.cuda.stream(st):
                tb = getattr(self, f"hyperTem{i + 1}").tables(E, time_eb, pool[9 + i] if len(pool) >= 15 else None)
                ev = torch.cuda.Event()
                ev.record(st)
            for t in tb:
                if isinstance(t, torch.Tensor):          # (the expanded views of a cap share these tensors' storage)
                    t.record_stream(main)
            pro["ht"].append(tb)
            pro["ht_ev"].append(ev)
        for i in range(2):
            st = pool[7 + i]
            st.wait_event(evs[1])
            st.wait_event(evs[2])
            teb.record_stream(st)
            time_eb_spg.record_stream(st)
            with torch.cuda.stream(st):
                tb = getattr(self, f"cap{i + 1}").tables(Es, time_eb_spg, teb)
                ev = torch.cuda.Event()
                ev.record(st)
            for t in tb:
                if isinstance(t, torch.Tensor):          # (the expanded views of a cap share these tensors' storage)
                    t.record_stream(main)
            pro["cap"].append(tb)
            pro["cap_ev"].append(ev)
        return pro

    def forward(self, source, x_in, pro=None):
        if pro is None:
            pro = self.prologue(source)
        ht, cp = pro["ht"], pro["cap"]
        hev, cev = pro.get("ht_ev"), pro.get("cap_ev")

        def join(evl, i):
            if evl is not None:
                torch.cuda.current_stream().wait_event(evl[i])

        join(hev, 0)
        x = self.hyperTem1(x_in, None, None, ht[0])
        join(cev, 0)
        x, HS1, _ = self.cap1(x, None, None, None, cp[0])
        join(hev, 1)
        x = self.hyperTem2(x, None, None, ht[1])
        join(hev, 2)
        x = self.hyperTem3(x, None, None, ht[2])
        join(cev, 1)
        x, HS3, _ = self.cap2(x, None, None, None, cp[1])
        join(hev, 3)
        x = self.hyperTem4(x, None, None, ht[3])
        return x, HS1, HS3


def _exact_count_mask(u, k):
    """1 everywhere, 0 at the k largest entries of u -- same sort/scatter ops as the reference (:317-321)."""
    _, order = torch.sort(u, dim=0, descending=True)
    return torch.ones_like(order).scatter_(0, order[:k], 0)


class Hypergraph_encoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.device = args.device
        self.num_node, self.input_base_dim = args.num_nodes, args.input_base_dim
        self.hidden_dim, self.horizon = args.hidden_dim, args.lag
        self.embed_dim, self.HS = args.embed_dim, args.HS
        self.mode, self.scaler_zeros = args.mode, args.scaler_zeros
        self.mask_ratio, self.ada_mask_ratio, self.ada_type = args.mask_ratio, args.ada_mask_ratio, args.ada_type
        self.change_epoch, self.epochs = args.change_epoch, args.epochs
        self.dim_in_flow = nn.Linear(self.input_base_dim, self.hidden_dim, bias=True)
        self.STHCN_encode = STHCN(args)
        # the reference draws an unused (D,T,HS,N) tensor here (ref :305); keep the CPU RNG stream aligned
        torch.randn(self.hidden_dim * self.horizon * self.HS * self.num_node)
        self.MLP_RL = MLP_RL(args.input_base_dim, self.HS, self.hidden_dim, self.embed_dim, self.device)
        self.teb4mask = time_feature(self.embed_dim)
        self.neb4mask = nn.Parameter(torch.randn(self.num_node, self.embed_dim))
        # hook for parity tests: replaces the arg-max class labels (near-ties can flip between implementations)
        self.label_c_override = None
        # graph replay hook: int64 device vector [order, adaptive_num, random_num] (see mask_plan)
        self.plan_override = None
        # hook for exact parity tests: {"u1": ..., "u2": ...} device vectors used INSTEAD of the torch.rand_like draws
        # (ref :316 / :389,:400); static buffers, so a captured step can be fed the oracle's draws replay by replay
        self.draws_override = None
        self._k_cache = {}

    # -- mask scorer (both phases), ref :326-332 / :338-343
    def _scores(self, source, aux=None):
        """Mask scorer.  aux = two side streams: the time embedding + time-adaptive tables and the node-adaptive tables are
        independent of the flow, so they run beside the first layer instead of in front of it (this chain sits on the
        critical path of the adaptive phase: the encoder cannot start before the mask exists)."""
        i0 = self.input_base_dim
        if aux is None:
            time_eb = self.teb4mask(source[:, :, 0, i0:i0 + 2])
            if source.is_cuda:
                return self.MLP_RL(source[..., 0:i0], time_eb, self.neb4mask, None, True)
            logits = self.MLP_RL(source[..., 0:i0], time_eb, self.neb4mask)
            return F.softmax(logits, dim=-1)
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        sa, sb = aux
        sa.wait_event(fork)
        with torch.cuda.stream(sa):
            time_eb = self.teb4mask(source[:, :, 0, i0:i0 + 2])
            Wt, bt = self.MLP_RL.tables_tem(time_eb)
            ev_t = torch.cuda.Event()
            ev_t.record(sa)
        sb.wait_event(fork)
        with torch.cuda.stream(sb):
            Wn, bn = self.MLP_RL.tables_spa(self.neb4mask)
            ev_n = torch.cuda.Event()
            ev_n.record(sb)
        for t_ in (Wt, bt, Wn, bn):
            t_.record_stream(cur)
        return self.MLP_RL(source[..., 0:i0], None, None, ((Wn, bn, ev_n), (Wt, bt, ev_t)), True)

    def _budgets(self, n_cells, epoch):
        tp = ((epoch - self.change_epoch) / (self.epochs - self.change_epoch)) * self.ada_mask_ratio
        tp = 1 if tp > 1 else tp
        total = int(n_cells * self.mask_ratio)
        ada = int(total * tp)
        return ada, total - ada

    def mask_plan(self, n_cells, epoch):
        """Host half of the phase-2 mask: shuffled class order (ref :357-358) and the two budgets (ref :348-353) as one
        int64 vector [order[0..H), adaptive_num, random_num].  Consumes python `random` exactly like the reference."""
        ada_num, rnd_num = self._budgets(n_cells, epoch)
        order = list(range(self.HS))
        random.shuffle(order)
        return torch.tensor(order + [ada_num, rnd_num], dtype=torch.int64)

    def _adaptive_mask(self, source, prob, epoch, draws=None):
        """Phase-2 mask (ref :344-413): whole classes in a shuffled order until the adaptive budget is reached, the
        last class sub-sampled, then an exact-count random fill.  Restated without host synchronisation: the class
        selection loop of the reference (one `torch.sum(...)` D2H per class) becomes a 10-element prefix sum on the
        device, and `order[:k]` with a device-side k becomes a scatter of (rank >= k).  Same draws, same sorts."""
        i0, H = self.input_base_dim, self.HS
        dev = prob.device
        if not prob.is_cuda:                         # host-side unit tests of the mask logic (the model itself is CUDA-only)
            return self._adaptive_mask_torch(source, prob, epoch)
        n = prob.numel() // H
        if self.plan_override is not None:           # graph replay: a static device buffer refreshed by the caller
            plan = self.plan_override
        else:
            plan = self.mask_plan(n, epoch).to(dev)
        # the same two draws, in the same order, as the reference (:389, :400)
        if self.draws_override is not None:
            u1, u2 = self.draws_override["u1"], self.draws_override["u2"]
        elif draws is not None:
            u1, u2 = draws
        else:
            u1 = torch.rand_like(source[..., 0:1].reshape(-1))
            u2 = torch.rand_like(source[..., 0:1].reshape(-1))
        # label = arg-max class (== sort(descending)[..., 0] of ref :344-345 up to exact ties) unless a test injects labels
        final = ops.mask_adaptive(prob, self.label_c_override, plan, u1, u2, i0, self.ada_type == "all")
        return final.reshape(prob.shape[:-1] + (i0,))

    def _adaptive_mask_torch(self, source, prob, epoch):
        """The same mask with torch ops only (the sort-based restatement the kernels replace); kept as the readable
        specification and used by the parity tests."""
        i0, H = self.input_base_dim, self.HS
        if self.label_c_override is not None:
            label_c = self.label_c_override
        else:
            label_c = torch.argmax(prob, dim=-1)      # == sort(descending)[..., 0] of the reference (:344-345) up to exact ties
        flat = label_c.reshape(-1)
        n = flat.numel()
        dev = flat.device
        if self.plan_override is not None:           # graph replay: a static device buffer refreshed by the caller
            plan = self.plan_override
        else:
            plan = self.mask_plan(n, epoch).to(dev)
        order, ada, rnd = plan[:H], plan[H], plan[H + 1]
        # class histogram without atomics (scatter_add_ of 130k ones into 10 bins serialises: 68 us on a B200)
        counts = (flat.unsqueeze(1) == torch.arange(H, device=dev)).sum(0)
        co = counts[order]
        picked = (torch.cumsum(co, 0) - co) < ada           # classes the reference's while-loop would add
        npick = picked.sum()
        idx = torch.arange(H, device=dev)
        if self.ada_type == "all":
            last = picked & (idx == npick - 1)
            multi = npick >= 2
            full_o = picked & ~last & multi                  # masked outright
            part_o = torch.where(multi, last, picked)        # sub-sampled
        else:
            full_o = torch.zeros_like(picked)
            part_o = picked
        lut = torch.zeros(H, dtype=torch.int64, device=dev).scatter_(0, order, part_o.long() + 2 * full_o.long())
        n_full = (co * full_o).sum()
        role = lut[flat]
        full = (role == 2).long()
        part = (role == 1).long()
        pos = torch.arange(n, device=dev)
        u1 = torch.rand_like(source[..., 0:1].reshape(-1))
        o1 = torch.sort(part * u1, dim=0, descending=True)[1]
        m_ada = torch.empty_like(o1).scatter_(0, o1, (pos >= (ada - n_full)).long()) * (1 - full)
        u2 = torch.rand_like(source[..., 0:1].reshape(-1))
        o2 = torch.sort(m_ada * u2, dim=0, descending=True)[1]
        m_rnd = torch.empty_like(o2).scatter_(0, o2, (pos >= rnd).long())
        final = (m_ada * m_rnd).reshape(label_c.shape).unsqueeze(-1)
        if i0 != 1:
            final = final.repeat(1, 1, 1, i0)
        return final

    def forward(self, source, label, epoch=None, pro=None):
        i0 = self.input_base_dim
        flow = source[..., 0:i0]
        if self.mode != "pretrain":
            enc, _, _ = self.STHCN_encode(source, _affine(self.dim_in_flow, flow), pro)
            return enc
        score = pro.get("score") if pro is not None else None     # (prob, event) computed on a side stream
        if epoch <= self.change_epoch:
            u = self.draws_override["u1"] if self.draws_override is not None else torch.rand_like(flow.reshape(-1))
            k = int(u.shape[0] * self.mask_ratio)
            kd = self._k_cache.get((k, u.device))
            if kd is None:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("first use of a mask budget inside a CUDA-graph capture: run one eager step first")
                kd = self._k_cache[(k, u.device)] = torch.tensor([k], dtype=torch.int64, device=u.device)
            final_mask = ops.mask_random(u, kd)
            final_mask = final_mask.reshape(-1, self.horizon, self.num_node, i0)
            if score is None:
                prob = self._scores(source)
            else:
                prob = score[0]                                   # joined by the caller before the loss
        else:
            draws = None
            if score is not None and self.draws_override is None and source.is_cuda:
                # the two uniform draws of the mask (ref :389, :400) do not depend on the scores: issue them BEFORE joining the
                # scorer stream, so they run beside it instead of between the score head and the mask kernels (same order, same
                # generator state as the reference)
                draws = (torch.rand_like(source[..., 0:1].reshape(-1)), torch.rand_like(source[..., 0:1].reshape(-1)))
            if score is None:
                prob = self._scores(source)
            else:
                torch.cuda.current_stream().wait_event(score[1])
                prob = score[0]
            final_mask = self._adaptive_mask(source, prob, epoch, draws)
        final_mask = final_mask.detach()
        if (i0 == 1 and source.is_cuda and source.dtype == torch.float32 and source.is_contiguous()
                and final_mask.dtype == torch.int64 and final_mask.is_contiguous()):
            # where(mask == 0, scaler_zeros, mask * flow) + dim_in_flow in one launch (ref :419-421)
            x_in = ops.masked_affine1(source, final_mask, self.dim_in_flow.weight, self.dim_in_flow.bias, float(self.scaler_zeros))
        else:
            masked = torch.where(final_mask == 0, torch.full_like(flow, float(self.scaler_zeros)), final_mask * flow)
            x_in = _affine(self.dim_in_flow, masked)
        enc, HS1, _ = self.STHCN_encode(source, x_in, pro)
        return enc, final_mask[..., :i0], prob, HS1.squeeze(-1).transpose(-1, -2)


class Hypergraph_decoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.input_base_dim, self.hidden_dim, self.mode = args.input_base_dim, args.hidden_dim, args.mode
        self.time_feature1_ = time_feature(args.embed_dim_spa)   # registered but never called (ref :446-447)
        self.time_feature2_ = time_feature(args.embed_dim_spa)
        self.STHCN_decode = STHCN(args)
        self.dim_flow_out = nn.Linear(self.hidden_dim, self.input_base_dim, bias=True)

    def forward(self, source, flow_encode_eb, pro=None):
        flow_decode, _, _ = self.STHCN_decode(source, flow_encode_eb, pro)
        lin = self.dim_flow_out
        if flow_decode.is_cuda and lin.out_features <= 4:
            return ops.proj_out(flow_decode, lin.weight, lin.bias), flow_decode      # one streaming kernel each way
        return lin(flow_decode), flow_decode


class GPTST_Model(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.num_node, self.input_base_dim, self.input_extra_dim = args.num_nodes, args.input_base_dim, args.input_extra_dim
        self.hidden_dim, self.output_dim, self.horizon = args.hidden_dim, args.output_dim, args.horizon
        self.embed_dim, self.embed_dim_spa = args.embed_dim, args.embed_dim_spa
        self.HS, self.HT, self.HT_Tem, self.num_route = args.HS, args.HT, args.HT_Tem, args.num_route
        self.mode, self.model = args.mode, args.model
        if args.hidden_dim not in (64, 128):
            raise ValueError("gptst_b200 kernels are built for hidden_dim 64 or 128")
        self.encoder = Hypergraph_encoder(args)
        self.decoder = Hypergraph_decoder(args)
        # overlap the parameter-side prologues on side streams (pretrain mode); plain attributes, deepcopy-safe
        self.side_streams = os.environ.get("GPTST_B200_SIDE_STREAMS", "1") != "0"
        self._streams = None

    def __deepcopy__(self, memo):
        streams, self._streams = self._streams, None          # CUDA stream handles are not copyable state
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy as _copy
            for k, v in self.__dict__.items():
                setattr(new, k, _copy.deepcopy(v, memo))
            return new
        finally:
            self._streams = streams

    def _side_prologues(self, source):
        """Run the activation-independent halves of both STHCNs on two side streams, concurrently with the mask
        scoring on the caller's stream.  Under CUDA-graph capture the fork/join becomes parallel graph branches; the
        autograd engine replays the same streams in backward, so the table gradients overlap the big kernels too."""
        main = torch.cuda.current_stream()
        dev = source.device
        key = (dev.index, main.cuda_stream)
        if self._streams is None or self._streams[0] != key:
            # the scorer (and its two table streams) is on the critical front of the adaptive phase -- the encoder cannot start
            # before the mask exists -- so it gets the priority of the main chain; the STHCN prologues keep the default (lowest).
            # Autograd replays the scorer's backward (KL branch) on the same streams, where it is NOT critical but competes with
            # the head of the main backward chain: GPTST_B200_SCORER_PRIO=low is the A/B knob for that trade.
            sp = 0 if os.environ.get("GPTST_B200_SCORER_PRIO", "high") == "low" else -1
            # per STHCN: 3 time embeddings, 4 hyperTem weight branches, 2 caps, 4 hyperTem mix-matrix branches, 2 accumulation streams
            self._streams = (key, [torch.cuda.Stream(device=dev) for _ in range(15)], [torch.cuda.Stream(device=dev) for _ in range(15)],
                             torch.cuda.Stream(device=dev, priority=sp), [torch.cuda.Stream(device=dev, priority=sp) for _ in range(2)])
        fork = torch.cuda.Event()
        fork.record(main)
        out = []
        # mask scorer (teb4mask + MLP_RL + softmax) on its own stream: off the critical path in the random-mask phase, and
        # its backward (KL branch) overlaps the main backward chain in the adaptive phase
        s3 = self._streams[3]
        s3.wait_event(fork)
        with torch.cuda.stream(s3):
            prob = self.encoder._scores(source, self._streams[4])
            ev3 = torch.cuda.Event()
            ev3.record(s3)
        prob.record_stream(main)
        for pool, sthcn in ((self._streams[1], self.encoder.STHCN_encode), (self._streams[2], self.decoder.STHCN_decode)):
            out.append(sthcn.prologue_streams(source, pool, fork, main))
        out[0]["score"] = (prob, ev3)
        return out

    def forward_pretrain(self, source, label, batch_seen=None, epoch=None):
        enc_pro, dec_pro = self._side_prologues(source) if self.side_streams else (None, None)
        flow_encode_eb, mask, probability, HS1 = self.encoder(source, label, epoch, enc_pro)
        flow_out, flow_decode = self.decoder(source, flow_encode_eb, dec_pro)
        if enc_pro is not None:
            torch.cuda.current_stream().wait_event(enc_pro["score"][1])   # join the scorer stream before the outputs are used
        return flow_out, flow_decode, 1 - mask, probability, HS1

    def forward_fune(self, source, label):
        e = self.encoder(source, label)
        return e, e, e, e, e

    def forward(self, source, label, batch_seen=None, epoch=None):
        if not source.is_cuda:
            raise RuntimeError("gptst_b200.GPTST_Model runs on CUDA only (no CPU fallback)")
        if self.mode == "pretrain":
            return self.forward_pretrain(source, label, batch_seen, epoch)
        return self.forward_fune(source, label)
