"""CPU oracle for the GPT-ST pre-training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement, in plain torch (CPU, fp32 or fp64), of the
arithmetic the reference performs in ``model/Pretrain_model/GPTST.py``.  It exists so
that the CUDA path in ``gpt-st_b200/`` can be checked against something that is (a)
readable, (b) importable on the GPU box where ``/root/reference`` does not exist and
(c) differentiable (``torch.autograd`` supplies the reference gradients).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``gpt-st_b200/`` does.

Pinning: the reference ships no tests / golden vectors for this path,
so the oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py``
imports ``/root/reference/model/Pretrain_model/GPTST.py`` (with the six ``'cuda:0'``
literals rewritten to ``'cpu'``), runs it on seeded inputs and stores inputs + outputs +
gradients under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through
this file.  It is also checked against the survey-time golden numbers for the shipped
PEMS08 checkpoint when ``baseline/_ref`` is present.

Everything is written functionally over a flat ``dict`` of tensors whose keys are the
reference ``state_dict`` names, so a reference checkpoint can be fed in
unchanged.

The restatement deliberately does NOT follow the reference's op sequence where an
algebraically identical but cheaper form exists (verified to <=2e-6 against the
reference by the golden tests):
  * the (B,T,H,N,D) outer product ``Dcaps_in`` (GPTST.py:106-107) is never built:
    ``sum_n c[h,n] * u[h,:] * P[n,:] == u[h,:] * (c @ P)[h,:]``;
  * all 'softmax over H' are taken on a (..., H, N) layout exactly like the reference.
"""
from __future__ import annotations

import math
import random as _pyrandom
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LRELU_SLOPE = 0.01  # nn.LeakyReLU() default, GPTST.py:18,96,152


# ----------------------------------------------------------------------------------------
# elementary pieces
# ----------------------------------------------------------------------------------------
def lrelu(x: Tensor) -> Tensor:
    return torch.where(x >= 0, x, x * LRELU_SLOPE)


def squash(x: Tensor) -> Tensor:
    """Capsule squash over the last dim.  GPTST.py:36-39."""
    q = (x * x).sum(dim=-1, keepdim=True)
    return (q / (1.0 + q)) * x / (q.sqrt() + 1e-8)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def time_feature(P: Dict[str, Tensor], pre: str, tf_in: Tensor) -> Tensor:
    """time_feature.forward, GPTST.py:198-202.  tf_in (B,T,2) -> (B,T,e)."""
    day = linear(tf_in[..., 0:1], P[pre + "ln_day.weight"], P[pre + "ln_day.bias"])
    week = linear(tf_in[..., 1:2], P[pre + "ln_week.weight"], P[pre + "ln_week.bias"])
    h = torch.relu(linear(day + week, P[pre + "ln1.weight"], P[pre + "ln1.bias"]))
    h = torch.relu(linear(h, P[pre + "ln2.weight"], P[pre + "ln2.bias"]))
    return linear(h, P[pre + "ln.weight"], P[pre + "ln.bias"])


def time_feature_spg(P: Dict[str, Tensor], pre: str, tf_in: Tensor) -> Tensor:
    """time_feature_spg.forward, GPTST.py:215-219.  Maps the T(=12) axis: (B,T,2) -> (B,e)."""
    day = linear(tf_in[..., 0], P[pre + "ln_day.weight"], P[pre + "ln_day.bias"])
    week = linear(tf_in[..., 1], P[pre + "ln_week.weight"], P[pre + "ln_week.bias"])
    h = torch.relu(linear(day + week, P[pre + "ln1.weight"], P[pre + "ln1.bias"]))
    h = torch.relu(linear(h, P[pre + "ln2.weight"], P[pre + "ln2.bias"]))
    return linear(h, P[pre + "ln.weight"], P[pre + "ln.bias"])


# ----------------------------------------------------------------------------------------
# a2: hyperTem  (GPTST.py:154-163)
# ----------------------------------------------------------------------------------------
def hypertem(eb: Tensor, node_emb: Tensor, time_eb: Tensor, adj: Tensor,
             weights_pool: Tensor, bias_pool: Tensor) -> Tensor:
    """eb (B,T,N,D); node_emb (N,d); time_eb (B,T,d); adj (d,Ht,T); pool (d,D,D); bias (d,D)."""
    # per-node temporal incidence A[n] (Ht,T)                                   :156
    A = torch.einsum("nk,kht->nht", node_emb, adj)
    # two hops over T with no nonlinearity in between                           :157-158
    hyper = torch.einsum("nht,btnd->bhnd", A, eb)
    ret = torch.einsum("nht,bhnd->btnd", A, hyper)
    # time-adaptive projection                                                   :160-162
    W = torch.einsum("btk,kio->btio", time_eb, weights_pool)
    bias = (time_eb @ bias_pool).unsqueeze(2)
    y = torch.einsum("btni,btio->btno", ret, W) + bias + eb
    return lrelu(y)                                                             # :163


# ----------------------------------------------------------------------------------------
# a1: cap  (GPTST.py:100-141)
# ----------------------------------------------------------------------------------------
def cap_routing(Pd: Tensor, ud: Tensor, dadj: Tensor, num_route: int) -> Tensor:
    """Dynamic routing on detached tensors, GPTST.py:108-120.

    Pd (B,T,N,D) primary capsules, ud (B,T,H,D) squashed cluster prototypes, dadj (B,T,H,N).
    Returns the routing logits b (B,T,H,N) (constant w.r.t. autograd).
    """
    b = torch.zeros_like(dadj)
    for _ in range(num_route):
        c = b.softmax(dim=2)                                   # over H        :114
        s = ud * torch.einsum("bthn,btnd->bthd", c, Pd)        # == (c*Dcaps_in).sum(-2)  :115
        v = squash(s)                                          #               :116
        b = b + torch.einsum("bthd,btnd->bthn", v, Pd)         #               :117-118
    return b


def cap(x: Tensor, node_emb: Tensor, time_eb_spg: Tensor, teb: Tensor,
        ln_p_w: Tensor, ln_p_b: Tensor, adj: Tensor, t_adj: Tensor,
        weights_spa: Tensor, bias_spa: Tensor, num_route: int = 2
        ) -> Tuple[Tensor, Tensor, Tensor]:
    """x (B,T,N,D); node_emb (N,d); time_eb_spg (B,ds); teb (B,T,ds); adj (ds,H,N);
    t_adj (ds,HT,T*H).  Returns (out (B,T,N,D), c (B,T,H,N) detached, dyn (B,HT,T*H) detached)."""
    B, T, N, D = x.shape
    H = adj.shape[1]
    P = squash(linear(x, ln_p_w, ln_p_b))                                       # :102-103
    dadj = torch.einsum("btk,khn->bthn", teb, adj)                              # :104
    u = squash(torch.einsum("bthn,btnd->bthd", dadj.softmax(dim=2), P))         # :105 (+squash of :106)
    b = cap_routing(P.detach(), u.detach(), dadj.detach(), num_route)           # :108-118
    c = (b + dadj).softmax(dim=2)                                               # :120
    s = torch.einsum("bthn,btnd->bthd", c, P)                                   # :123
    tau = (torch.arange(1, T + 1, dtype=x.dtype, device=x.device) / 12.0).view(1, T, 1, 1)  # :97,125
    S = (s + tau).reshape(B, T * H, D)                                          # :126-127
    dyn = torch.einsum("bk,khj->bhj", time_eb_spg, t_adj)                       # :129
    e1 = lrelu(dyn @ S)                                                         # :130
    r = lrelu(dyn.transpose(1, 2) @ e1).reshape(B, T, H, D) + s                 # :131-132
    v = squash(r)                                                               # :134
    recon = torch.einsum("bthn,bthd->btnd", c, v)                               # :135
    W = torch.einsum("nk,kio->nio", node_emb, weights_spa)                      # :137
    bias = node_emb @ bias_spa                                                  # :138
    y = torch.einsum("btni,nio->btno", recon, W) + bias + x                     # :139,141
    return lrelu(y), c.detach(), dyn.detach()


# ----------------------------------------------------------------------------------------
# a3: MLP_RL mask scorer (GPTST.py:21-34)
# ----------------------------------------------------------------------------------------
def mlp_rl(flow: Tensor, time_eb: Tensor, node_eb: Tensor, P: Dict[str, Tensor], pre: str) -> Tensor:
    """flow (B,T,N,ibd) -> logits (B,T,N,H)."""
    h0 = linear(flow, P[pre + "ln1.weight"], P[pre + "ln1.bias"])              # :22
    Wn = torch.einsum("nk,kio->nio", node_eb, P[pre + "weights_pool_spa"])     # :24
    bn = node_eb @ P[pre + "bias_pool_spa"]                                    # :25
    h1 = lrelu(torch.einsum("btni,nio->btno", h0, Wn) + bn)                    # :26-27
    Wt = torch.einsum("btk,kio->btio", time_eb, P[pre + "weights_pool_tem"])   # :29
    bt = (time_eb @ P[pre + "bias_pool_tem"]).unsqueeze(-2)                    # :30
    h2 = lrelu(torch.einsum("btni,btio->btno", h1, Wt) + bt)                   # :31-32
    return linear(h2, P[pre + "ln3.weight"], P[pre + "ln3.bias"])              # :33


# ----------------------------------------------------------------------------------------
# a6: STHCN (GPTST.py:253-273)
# ----------------------------------------------------------------------------------------
def sthcn(P: Dict[str, Tensor], pre: str, source: Tensor, x_in: Tensor, ibd: int, num_route: int
          ) -> Tuple[Tensor, Tensor, Tensor]:
    tf_in = source[:, :, 0, ibd:ibd + 2]                                        # :256-257 (node 0)
    time_eb = time_feature(P, pre + "time_feature1.", tf_in)                    # :259
    teb = time_feature(P, pre + "time_feature1_.", tf_in)                       # :260
    time_eb_spg = time_feature_spg(P, pre + "time_feature2.", tf_in)            # :261
    E, Es = P[pre + "node_embeddings"], P[pre + "node_embeddings_spg"]

    def ht(i: int, x: Tensor) -> Tensor:
        q = f"{pre}hyperTem{i}."
        return hypertem(x, E, time_eb, P[q + "adj"], P[q + "weights_pool"], P[q + "bias_pool"])

    def cp(i: int, x: Tensor):
        q = f"{pre}cap{i}."
        return cap(x, Es, time_eb_spg, teb, P[q + "ln_p.weight"], P[q + "ln_p.bias"], P[q + "adj"],
                   P[q + "t_adj"], P[q + "weights_spa"], P[q + "bias_spa"], num_route)

    x1 = ht(1, x_in)                                                            # :265
    x2, hs1, _ = cp(1, x1)                                                      # :266
    x3 = ht(2, x2)                                                              # :267
    x4 = ht(3, x3)                                                              # :269
    x5, hs3, _ = cp(2, x4)                                                      # :270
    x6 = ht(4, x5)                                                              # :271
    return x6, hs1, hs3


# ----------------------------------------------------------------------------------------
# a5: mask construction (GPTST.py:312-418).  Random draws are INPUTS so that the same
# draws can be fed to the reference, the oracle and the CUDA path.
# ----------------------------------------------------------------------------------------
def exact_count_mask(u: Tensor, k: int) -> Tensor:
    """ones everywhere except the positions of the k largest entries of u (ties broken the way
    torch.sort(descending=True) breaks them).  GPTST.py:317-321 / 392-397 / 402-406."""
    _, order = torch.sort(u, dim=0, descending=True)
    m = torch.ones_like(order)
    return m.scatter_(0, order[:k], 0)


def mask_budgets(n_cells: int, mask_ratio: float, ada_mask_ratio: float, epoch: int,
                 change_epoch: int, epochs: int) -> Tuple[int, int, int]:
    """(mask_num_sum, adaptive_mask_num, random_mask_num), GPTST.py:348-353."""
    tp = ((epoch - change_epoch) / (epochs - change_epoch)) * ada_mask_ratio
    if tp > 1:
        tp = 1
    total = int(n_cells * mask_ratio)
    ada = int(total * tp)
    return total, ada, total - ada


def adaptive_mask(label_c: Tensor, class_order: Sequence[int], u_adaptive: Tensor, u_random: Tensor,
                  adaptive_num: int, random_num: int, ada_type: str) -> Tensor:
    """Phase-2 mask, GPTST.py:356-410.  label_c (B,T,N) int64; u_* flat (B*T*N,) uniform draws.
    Returns final_mask (B,T,N) int64 with 0 = masked."""
    flat = label_c.reshape(-1)
    counts = torch.bincount(flat, minlength=max(class_order) + 1)
    # add whole classes in the shuffled order until the budget is reached      :365-369 / 380-382
    picked, total = 0, 0
    while total < adaptive_num:
        total += int(counts[class_order[picked]])
        picked += 1
    chosen = class_order[:picked]
    if ada_type == "all" and picked >= 2:
        # all but the last picked class are masked outright, the last is sub-sampled  :370-374
        full = torch.isin(flat, torch.tensor(chosen[:-1], dtype=flat.dtype)).to(torch.int64)
        part = (flat == chosen[-1]).to(torch.int64)
        n_full = int(full.sum())
    else:
        full = torch.zeros_like(flat)
        part = (torch.isin(flat, torch.tensor(chosen, dtype=flat.dtype)).to(torch.int64)
                if picked > 0 else torch.zeros_like(flat))
        n_full = 0
    m_ada = exact_count_mask(part.to(u_adaptive.dtype) * u_adaptive, adaptive_num - n_full)  # :389-397
    m_ada = m_ada * (1 - full)                                                               # :398
    m_rnd = exact_count_mask(m_ada.to(u_random.dtype) * u_random, random_num)                # :400-406
    return (m_ada * m_rnd).reshape(label_c.shape)                                            # :410-411


# ----------------------------------------------------------------------------------------
# a5/a8: encoder / decoder / model
# ----------------------------------------------------------------------------------------
class Draws:
    """The random numbers one pre-training forward consumes, in the order the reference draws them."""

    def __init__(self, u1: Tensor, class_order: Optional[List[int]] = None, u2: Optional[Tensor] = None):
        self.u1, self.class_order, self.u2 = u1, class_order, u2

    @staticmethod
    def sample(n_cells_phase1: int, n_cells: int, H: int, phase2: bool, gen: torch.Generator,
               pyrand: _pyrandom.Random, dtype=torch.float32) -> "Draws":
        if not phase2:
            return Draws(torch.rand(n_cells_phase1, generator=gen, dtype=dtype))
        order = list(range(H))
        pyrand.shuffle(order)                                                   # :357-358
        u1 = torch.rand(n_cells, generator=gen, dtype=dtype)                    # :389
        u2 = torch.rand(n_cells, generator=gen, dtype=dtype)                    # :400
        return Draws(u1, order, u2)


def encoder(P: Dict[str, Tensor], cfg, source: Tensor, epoch: Optional[int], draws: Optional[Draws],
            label_c_override: Optional[Tensor] = None):
    """Hypergraph_encoder.forward, GPTST.py:312-427.  cfg: namespace with the args fields."""
    ibd = cfg.input_base_dim
    pre = "encoder."
    flow = source[..., 0:ibd]
    if cfg.mode != "pretrain":
        x = linear(flow, P[pre + "dim_in_flow.weight"], P[pre + "dim_in_flow.bias"])    # :420
        enc, _, _ = sthcn(P, pre + "STHCN_encode.", source, x, ibd, cfg.num_route)
        return enc

    B, T, N, _ = source.shape
    tf_in = source[:, :, 0, ibd:ibd + 2]
    logits = mlp_rl(flow, time_feature(P, pre + "teb4mask.", tf_in), P[pre + "neb4mask"],
                    P, pre + "MLP_RL.")                                         # :326-329 / 338-340
    prob = logits.softmax(dim=-1)                                               # :332 / 343
    if epoch <= cfg.change_epoch:
        k = int(B * T * N * ibd * cfg.mask_ratio)                               # :318
        final_mask = exact_count_mask(draws.u1, k).reshape(B, T, N, ibd)        # :316-323
    else:
        if label_c_override is not None:
            label_c = label_c_override
        else:
            label_c = torch.sort(prob, dim=-1, descending=True)[1][..., 0]      # :344-345
        _, ada, rnd = mask_budgets(B * T * N, cfg.mask_ratio, cfg.ada_mask_ratio, epoch,
                                   cfg.change_epoch, cfg.epochs)
        final_mask = adaptive_mask(label_c, draws.class_order, draws.u1, draws.u2, ada, rnd,
                                   cfg.ada_type).unsqueeze(-1)
        if ibd != 1:
            final_mask = final_mask.repeat(1, 1, 1, ibd)                        # :412-413
    keep = final_mask.to(flow.dtype)
    masked = torch.where(final_mask == 0, torch.full_like(flow, float(cfg.scaler_zeros)), keep * flow)  # :416-417
    x = linear(masked, P[pre + "dim_in_flow.weight"], P[pre + "dim_in_flow.bias"])       # :418
    enc, hs1, _ = sthcn(P, pre + "STHCN_encode.", source, x, ibd, cfg.num_route)
    return enc, final_mask, prob, hs1.transpose(-1, -2)                         # :423-425


def model_forward(P: Dict[str, Tensor], cfg, source: Tensor, epoch: Optional[int] = None,
                  draws: Optional[Draws] = None, label_c_override: Optional[Tensor] = None):
    """GPTST_Model.forward, GPTST.py:480-493.  Same 5-tuple as the reference."""
    if cfg.mode != "pretrain":
        e = encoder(P, cfg, source, None, None)
        return e, e, e, e, e
    enc, mask, prob, hs1 = encoder(P, cfg, source, epoch, draws, label_c_override)
    dec, _, _ = sthcn(P, "decoder.STHCN_decode.", source, enc, cfg.input_base_dim, cfg.num_route)
    flow_out = linear(dec, P["decoder.dim_flow_out.weight"], P["decoder.dim_flow_out.bias"])  # :455
    return flow_out, dec, 1 - mask, prob, hs1


# ----------------------------------------------------------------------------------------
# a9: losses as the trainer applies them (Run.py:91-101, lib/metrics.py:11-18, BasicTrainer.py:83-88)
# ----------------------------------------------------------------------------------------
def masked_mae(pred: Tensor, true: Tensor, mask: Tensor, mean: float, std: float,
               mask_value: Optional[float] = 0.0) -> Tensor:
    p = (pred * std + mean) * mask
    t = (true * std + mean) * mask
    if mask_value is not None:
        sel = t > mask_value
        p, t = p[sel], t[sel]
    return (t - p).abs().mean()


def kl_sum(prob: Tensor, target: Tensor) -> Tensor:
    """nn.KLDivLoss(reduction='sum')(prob.log(), target), Run.py:132 + BasicTrainer.py:85."""
    return torch.xlogy(target, target).sum() - (target * prob.log()).sum()


def pretrain_loss(outs, source: Tensor, cfg, epoch: int, mean: float, std: float) -> Tensor:
    flow_out, _, inv_mask, prob, hs1 = outs
    loss = masked_mae(flow_out, source[..., :cfg.output_dim], inv_mask, mean, std, 0.0)
    if epoch > cfg.change_epoch:
        loss = loss + 0.1 * kl_sum(prob, hs1)
    return loss


def synthetic_loss(outs, source: Tensor, epoch: int, change_epoch: int = 10) -> Tensor:
    """The loss of the survey/driver GPU probe: |(o - x) * mask|.mean() + 0.1 KL."""
    o, _, inv_mask, prob, hs = outs
    loss = ((o - source[..., :1]) * inv_mask).abs().mean()
    if epoch > change_epoch:
        loss = loss + 0.1 * kl_sum(prob, hs)
    return loss


# ----------------------------------------------------------------------------------------
# parameter construction with the reference's names / shapes / registration order (§8b)
# ----------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------------
# eval path: the gate of Enhance_model and STGCN's gated temporal convolution
# ---------------------------------------------------------------------------------------------------
def fusion_gate(flow_eb: Tensor, time_eb: Tensor, P: Dict[str, Tensor], pre: str = "") -> Tensor:
    """reference model/Model.py:12-18 (Fusion.forward): z = sigmoid(HS_fc(x) + HT_fc(y)); output_fc(z x + (1 - z) y)."""
    z = torch.sigmoid(linear(flow_eb, P[pre + "HS_fc.weight"], P[pre + "HS_fc.bias"]) +
                      linear(time_eb, P[pre + "HT_fc.weight"], P[pre + "HT_fc.bias"]))
    h = z * flow_eb + (1 - z) * time_eb
    return linear(h, P[pre + "output_fc.weight"], P[pre + "output_fc.bias"])


def eval_glue(source: Tensor, x_pre: Tensor, P: Dict[str, Tensor], ibd: int) -> Tensor:
    """reference model/Model.py:106-109: lin_test on the raw flow channels, then the gate with the encoder output."""
    x_t1 = linear(source[..., :ibd], P["lin_test.weight"], P["lin_test.bias"])
    return fusion_gate(x_pre, x_t1, P, "fusion.")


def temporal_conv_glu(x: Tensor, conv_w: Tensor, conv_b: Tensor, align_w: Optional[Tensor] = None,
                      align_b: Optional[Tensor] = None) -> Tensor:
    """reference model/STGCN/stgcn.py:10-53, TemporalConvLayer with act = "GLU" on x (B, c_in, T, N):
    conv = Conv2d(c_in, 2 c_out, (kt, 1), padding (kt-1)//2) over the time axis; x_in = Align(x) (1x1 conv when c_in > c_out,
    zero-padded channels when c_in < c_out); out = (conv[:, :c_out] + x_in) * sigmoid(conv[:, c_out:]).
    Restated without nn.Conv2d: an explicit sum over the kt taps of shifted, zero-padded slices."""
    B, c_in, T, N = x.shape
    c2, _, kt, _ = conv_w.shape
    c_out = c2 // 2
    pad = (kt - 1) // 2
    T_out = T + 2 * pad - kt + 1
    xp = torch.zeros(B, c_in, T + 2 * pad, N, dtype=x.dtype)
    xp[:, :, pad:pad + T] = x
    conv = conv_b.view(1, c2, 1, 1).expand(B, c2, T_out, N).clone()
    for k in range(kt):
        conv = conv + torch.einsum("oi,bitn->botn", conv_w[:, :, k, 0], xp[:, :, k:k + T_out])
    if c_in > c_out:
        x_in = torch.einsum("oi,bitn->botn", align_w[:, :, 0, 0], x) + align_b.view(1, c_out, 1, 1)
    elif c_in < c_out:
        x_in = torch.cat([x, torch.zeros(B, c_out - c_in, T, N, dtype=x.dtype)], dim=1)
    else:
        x_in = x
    return (conv[:, :c_out] + x_in) * torch.sigmoid(conv[:, c_out:])


def param_shapes(cfg) -> "list[tuple[str, tuple]]":
    N, D, d, ds = cfg.num_nodes, cfg.hidden_dim, cfg.embed_dim, cfg.embed_dim_spa
    H, HT, Ht, T, ibd = cfg.HS, cfg.HT, cfg.HT_Tem, cfg.horizon, cfg.input_base_dim

    def tfeat(pre, e, first):
        return [(pre + "ln_day.weight", (e, first)), (pre + "ln_day.bias", (e,)),
                (pre + "ln_week.weight", (e, first)), (pre + "ln_week.bias", (e,)),
                (pre + "ln1.weight", (e, e)), (pre + "ln1.bias", (e,)),
                (pre + "ln2.weight", (e, e)), (pre + "ln2.bias", (e,)),
                (pre + "ln.weight", (e, e)), (pre + "ln.bias", (e,))]

    def st(pre):
        out = [(pre + "node_embeddings", (N, d)), (pre + "node_embeddings_spg", (N, d))]
        for i in range(1, 5):
            q = f"{pre}hyperTem{i}."
            out += [(q + "adj", (d, Ht, T)), (q + "weights_pool", (d, D, D)), (q + "bias_pool", (d, D))]
        out += tfeat(pre + "time_feature1.", d, 1) + tfeat(pre + "time_feature1_.", ds, 1)
        out += tfeat(pre + "time_feature2.", ds, 12)
        for i in (1, 2):
            q = f"{pre}cap{i}."
            out += [(q + "t_adj", (ds, HT, T * H)), (q + "adj", (ds, H, N)), (q + "weights_spa", (d, D, D)),
                    (q + "bias_spa", (d, D)), (q + "ln_p.weight", (D, D)), (q + "ln_p.bias", (D,))]
        return out

    shapes = [("encoder.neb4mask", (N, d)), ("encoder.dim_in_flow.weight", (D, ibd)),
              ("encoder.dim_in_flow.bias", (D,))]
    shapes += st("encoder.STHCN_encode.")
    q = "encoder.MLP_RL."
    shapes += [(q + "weights_pool_spa", (d, D, D)), (q + "bias_pool_spa", (d, D)),
               (q + "weights_pool_tem", (d, D, D)), (q + "bias_pool_tem", (d, D)),
               (q + "ln1.weight", (D, ibd)), (q + "ln1.bias", (D,)),
               (q + "ln3.weight", (H, D)), (q + "ln3.bias", (H,))]
    shapes += tfeat("encoder.teb4mask.", d, 1)
    shapes += tfeat("decoder.time_feature1_.", ds, 1) + tfeat("decoder.time_feature2_.", ds, 1)
    shapes += st("decoder.STHCN_decode.")
    shapes += [("decoder.dim_flow_out.weight", (ibd, D)), ("decoder.dim_flow_out.bias", (ibd,))]
    return shapes


def init_params(cfg, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """xavier_uniform_ for dim>1, uniform(0,1) for 1-D params, as Run.py:79-85 does, from a
    private CPU generator (values are NOT bit-equal to a reference init; tests copy tensors)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shp in param_shapes(cfg):
        if len(shp) > 1:
            rf = 1
            for s in shp[2:]:
                rf *= s
            fan_in, fan_out = shp[1] * rf, shp[0] * rf
            a = math.sqrt(6.0 / (fan_in + fan_out))
            out[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype)
        else:
            out[name] = torch.rand(shp, generator=g, dtype=torch.float64).to(dtype)
    return out
