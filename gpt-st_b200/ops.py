"""Autograd functions over the C ABI (include/gptst_b200.h) for the three heavy blocks of
GPT-ST's pre-training model: ``hypertem`` (GPTST.py:154-163), ``cap`` (GPTST.py:100-141) and the
adaptive projections of ``MLP_RL`` (GPTST.py:24-32).

Only the (B,T,N,D)-sized work runs here; the parameter-sized contractions that feed it (node/time
adaptive weight tables, incidence logits) are produced by the caller (GPTST.py of this package).
CUDA tensors only -- there is no fallback path.
"""
from __future__ import annotations

import os

import torch

from . import _lib

PREC_TF32 = 1
PREC_3XTF32 = 3


def default_precision() -> int:
    """3xTF32 (fp32-faithful) unless GPTST_B200_PRECISION=tf32 asks for single-pass TF32."""
    v = os.environ.get("GPTST_B200_PRECISION", "3xtf32").lower()
    if v in ("tf32", "1"):
        return PREC_TF32
    if v in ("3xtf32", "3", "fp32"):
        return PREC_3XTF32
    raise ValueError(f"GPTST_B200_PRECISION={v!r}: expected 'tf32' or '3xtf32'")


_launches = 0  # kernels of libgptst_b200.so launched by this process (every C-ABI call launches exactly one)


def launch_count() -> int:
    return _launches


_CAPTURE_DEBUG = os.environ.get("GPTST_B200_CAPTURE_DEBUG", "0") == "1"


def _stream() -> int:
    global _launches
    _launches += 1
    st = torch.cuda.current_stream().cuda_stream
    if _CAPTURE_DEBUG:     # debugging aid: find the first launch after a stream capture got invalidated
        from cuda.bindings import runtime as _rt
        err, status = _rt.cudaStreamIsCapturing(st)
        if int(status) == 2:
            raise RuntimeError(f"stream capture already INVALIDATED before launch #{_launches} (err {err})")
    return st


def _count(n: int) -> None:
    global _launches
    _launches += n


def _chk(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("gptst_b200 ops need CUDA tensors (no CPU fallback)")
        if t.device.index != torch.cuda.current_device():
            raise RuntimeError(f"gptst_b200 ops launch on the CURRENT device's stream: tensor on {t.device}, current device "
                               f"cuda:{torch.cuda.current_device()} (wrap the call in torch.cuda.device(t.device))")
        if t.dtype != torch.float32:
            raise RuntimeError(f"gptst_b200 ops are fp32, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("gptst_b200 ops need contiguous tensors")


def _c(t):
    return None if t is None else t.contiguous()


def _p(t):
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------------
# thin kernel wrappers (no autograd)
# ---------------------------------------------------------------------------------------------------
def sum_partials(*parts):
    """Each argument is a (P, ...) tensor of per-CTA / per-split partials; returns the tuple of sums over dim 0, all computed by
    ONE launch (fixed order, deterministic).  Tensors with P == 1 are returned as views."""
    import ctypes as C
    todo = [(i, t.contiguous()) for i, t in enumerate(parts) if t.shape[0] > 1]
    outs = [t[0] if t.shape[0] == 1 else None for t in parts]
    for k0 in range(0, len(todo), 8):
        grp = todo[k0:k0 + 8]
        res = [torch.empty(t.shape[1:], device=t.device, dtype=torch.float32) for _, t in grp]
        n = len(grp)
        ins = (C.c_void_p * n)(*[t.data_ptr() for _, t in grp])
        ous = (C.c_void_p * n)(*[r.data_ptr() for r in res])
        numel = (C.c_long * n)(*[r.numel() for r in res])
        cnt = (C.c_int * n)(*[t.shape[0] for _, t in grp])
        _lib.check(_lib.lib().gptst_sum_partials(C.cast(ins, C.c_void_p), C.cast(ous, C.c_void_p), C.cast(numel, C.c_void_p),
                                                 C.cast(cnt, C.c_void_p), n, _stream()), "gptst_sum_partials")
        for (i, _), r in zip(grp, res):
            outs[i] = r
    return tuple(outs)
def gproj_fwd(X, W, bias, res, *, node_grouped: bool, act: bool, prec: int):
    """X (B,T,N,D).  time-grouped: W (B*T,D,D)/(B,T,D,D); node-grouped: W (N,D,D)."""
    B, T, N, D = X.shape
    X, W, bias, res = _c(X), _c(W), _c(bias), _c(res)
    _chk(X, W)
    Y = torch.empty_like(X)
    if node_grouped:
        G, R, gs, rs = N, B * T, D, N * D
    else:
        G, R, gs, rs = B * T, N, N * D, D
    rc = _lib.lib().gptst_gproj_fwd(_p(X), _p(W), _p(bias), _p(res), _p(Y), G, R, gs, rs, D, int(act), prec, _stream())
    _lib.check(rc, "gptst_gproj_fwd")
    return Y


def gproj_bwd(dY, Y, X, W, *, node_grouped: bool, act: bool, prec: int, want_dres: bool, sum_parts: bool = True):
    B, T, N, D = X.shape
    dY, Y, X, W = _c(dY), _c(Y), _c(X), _c(W)
    _chk(dY, X, W)
    if node_grouped:
        G, R, gs, rs = N, B * T, D, N * D
    else:
        G, R, gs, rs = B * T, N, N * D, D
    L = _lib.lib()
    splits = L.gptst_gproj_splits(G, R, D)
    dX = torch.empty_like(X)
    dWp = torch.empty((splits, G, D, D), device=X.device, dtype=torch.float32)
    dbp = torch.empty((splits, G, D), device=X.device, dtype=torch.float32)
    dres = torch.empty_like(X) if want_dres else None
    rc = L.gptst_gproj_bwd(_p(dY), _p(Y) if act else None, _p(X), _p(W), _p(dX), _p(dWp), _p(dbp), _p(dres), G, R, gs, rs,
                           D, int(act), prec, splits, _stream())
    _lib.check(rc, "gptst_gproj_bwd")
    if not sum_parts:
        return dX, dWp, dbp, dres          # raw (splits, ...) partials: the caller sums them together with others
    dW, db = sum_partials(dWp, dbp)
    return dX, dW, db, dres


def tmix(x, M, out=None, *, transpose=False, accumulate=False):
    B, T, N, D = x.shape
    x, M = _c(x), _c(M)
    _chk(x, M)
    if out is None:
        out = torch.empty_like(x)
    rc = _lib.lib().gptst_tmix(_p(x), _p(M), _p(out), B, T, N, D, int(transpose), int(accumulate), _stream())
    _lib.check(rc, "gptst_tmix")
    return out


def tmix_bwd(dy, x, M, dx_io, prec, raw=False):
    """Fused backward of the mix (D = 64): dx_io += M^T o dy (in place) and returns dM (raw: its (splits, N, T, T) partials)."""
    B, T, N, D = x.shape
    dy, x, M = _c(dy), _c(x), _c(M)
    _chk(dy, x, M, dx_io)
    L = _lib.lib()
    splits = L.gptst_tmix_bwd_splits(B, N)
    part = torch.empty((splits, N, T, T), device=x.device, dtype=torch.float32)
    rc = L.gptst_tmix_bwd(_p(dy), _p(x), _p(M), _p(dx_io), _p(part), B, T, N, D, prec, splits, _stream())
    _lib.check(rc, "gptst_tmix_bwd")
    if raw:
        return part
    return part[0] if splits == 1 else part.sum(0)


def tmix_dM(dy, x, raw=False):
    B, T, N, D = x.shape
    dy, x = _c(dy), _c(x)
    _chk(dy, x)
    L = _lib.lib()
    splits = L.gptst_tmix_dM_splits(B, N)
    part = torch.empty((splits, N, T, T), device=x.device, dtype=torch.float32)
    rc = L.gptst_tmix_dM(_p(dy), _p(x), _p(part), B, T, N, D, splits, _stream())
    _lib.check(rc, "gptst_tmix_dM")
    if raw:
        return part
    return part[0] if splits == 1 else part.sum(0)


# ---------------------------------------------------------------------------------------------------
# Deferred partial sums.  The backward kernels write per-CTA / per-split partials of the parameter-side gradients
# (dM_n, dW_n, db_n, ddyn, dWp, dbp).  Summing them at the tail of the block's backward put ~150 us of reductions on the
# main chain of a PEMS08 step although nothing on that chain reads the sums.  Instead the block takes those inputs
# EXPANDED to (P, *shape) (a stride-0 view, `expand_partials`) and returns the raw (P, *shape) partials as their
# gradient; the reduction then is the expand's own backward, which autograd runs on the stream the expand was made on --
# the block's table stream when `cap.tables` / `hyperTem.tables` (GPTST.py of this package) made it in the prologue.
# ---------------------------------------------------------------------------------------------------
class _ExpandPartials(torch.autograd.Function):
    """(shape) -> (P, *shape) stride-0 views of several tensors at once; the backward sums the P gradient partials of all of them
    with ONE gptst_sum_partials launch (fixed order, deterministic) on the stream this node's forward ran on."""

    @staticmethod
    def forward(ctx, counts, *ts):
        ctx.set_materialize_grads(False)
        return tuple(t.unsqueeze(0).expand((int(p),) + tuple(t.shape)) for t, p in zip(ts, counts))

    @staticmethod
    def backward(ctx, *gs):
        idx = [i for i, g in enumerate(gs) if g is not None]
        sums = sum_partials(*[gs[i] for i in idx]) if idx else ()
        out = [None] * len(gs)
        for i, s in zip(idx, sums):
            out[i] = s
        return (None, *out)


def expand_partials_many(ts, counts):
    """Stride-0 (P_i, *shape_i) views of the CUDA tensors `ts`; gradients arriving as (P_i, *shape_i) partials are summed by the
    views' backward (one launch for all of them; GPTST_B200_EXPAND=native uses torch's own expand / sum_to_size instead)."""
    if os.environ.get("GPTST_B200_EXPAND", "fused") == "native":
        return tuple(t.unsqueeze(0).expand((int(p),) + tuple(t.shape)) for t, p in zip(ts, counts))
    return _ExpandPartials.apply(tuple(int(p) for p in counts), *ts)


def expand_partials(t, P: int):
    """(shape) -> (P, *shape) stride-0 view whose backward sums the P gradient partials."""
    return expand_partials_many((t,), (P,))[0]


def hypertem_partial_count(B: int, N: int, D: int) -> int:
    """Number of dM_n partials the hyperTem backward writes for this geometry."""
    L = _lib.lib()
    return int(L.gptst_tmix_bwd_splits(B, N) if D == 64 else L.gptst_tmix_dM_splits(B, N))


def cap_partial_counts(B: int, T: int, N: int, D: int, H: int):
    """(P of dW_n / db_n, P of ddyn, P of dWp / dbp) the cap backward writes for this geometry."""
    L = _lib.lib()
    p_wn = int(L.gptst_gproj_splits(N, B * T, D))
    p_dyn = int(L.gptst_cap_hop_bwd_parts(D))
    if L.gptst_cap_route2_supported(N, D, H):
        p_wp = int(L.gptst_linear_bwd_acc_splits(B * T * N, D))
    else:
        p_wp = int(L.gptst_cap_route_bwd_parts(B, T, N, D, H))
    return p_wn, p_dyn, p_wp


# ---------------------------------------------------------------------------------------------------
# hyperTem core:  out = LReLU( (M_n o eb) . W_bt + bias_bt + eb )
# ---------------------------------------------------------------------------------------------------
class _HyperTemCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eb, Mn_e, W, bias, prec):
        eb, W, bias = eb.contiguous(), W.contiguous(), bias.contiguous()
        B, T, N, D = eb.shape
        if Mn_e.dim() != 4 or Mn_e.shape[0] != hypertem_partial_count(B, N, D):
            raise RuntimeError("hypertem_core: Mn must be expanded to (hypertem_partial_count(B, N, D), N, T, T)")
        Mn = Mn_e[0].contiguous()
        ret = tmix(eb, Mn)
        out = gproj_fwd(ret, W, bias, eb, node_grouped=False, act=True, prec=prec)
        ctx.save_for_backward(eb, Mn, W, ret, out)
        ctx.prec = prec
        return out

    @staticmethod
    def backward(ctx, dout):
        eb, Mn, W, ret, out = ctx.saved_tensors
        dout = dout.contiguous()
        dret, dW, db, deb = gproj_bwd(dout, out, ret, W, node_grouped=False, act=True, prec=ctx.prec, want_dres=True)
        B, T, N, D = eb.shape
        if D == 64:
            dM_part = tmix_bwd(dret, eb, Mn, deb, ctx.prec, raw=True)
        else:
            tmix(dret, Mn, deb, transpose=True, accumulate=True)
            dM_part = tmix_dM(dret, eb, raw=True)
        return deb, dM_part, dW.view(B, T, D, D), db.view(B, T, D), None


# ---------------------------------------------------------------------------------------------------
# fused hyperTem (csrc/htem_fused.cu): ONE TMA-fed kernel per direction on the main chain.  The parameter-side gradients
# (dW_bt = ret^T dy, db_bt, dM_n = dret . eb) are not needed by anything on the main chain, so they belong to a second
# autograd node, `_HyperTemParams`, that sits between the table generators and the block.  Autograd runs a node's backward
# on the stream its forward ran on: when `hyperTem.tables` builds the node on the block's table stream, those two
# kernels run there, overlapped with the next block's backward.  The block hands them their inputs through a mailbox and
# returns stride-0 zero placeholders as the "gradients" of the node's outputs (autograd only needs the dependency).
# ---------------------------------------------------------------------------------------------------
def hypertem_fused_enabled(D: int, T: int, prec: int) -> bool:
    """Fused kernels cover D = 64, T = 12, three-term split (GPTST_B200_HTEM=split selects the unfused pair)."""
    return D == 64 and T == 12 and prec == PREC_3XTF32 and os.environ.get("GPTST_B200_HTEM", "fused") == "fused"


_zero_scalars = {}


def _zero_like_shape(shape, device):
    """Stride-0 zero tensor of `shape` (placeholder gradient; nobody reads its values)."""
    z = _zero_scalars.get(device)
    if z is None:
        z = _zero_scalars[device] = torch.zeros((), device=device, dtype=torch.float32)
    return z.expand(tuple(shape))


class _HyperTemParamsW(torch.autograd.Function):
    """(W, bias) -> the same two tensors; the forward also packs W into the fragment-ordered fp16 tables of the fused kernels
    (mailbox['wf'], ['wb']).  The backward ignores its incoming placeholders: the block's backward has already launched the
    dW_bt / db_bt kernel on THIS node's stream (as soon as dOut existed, beside the main backward kernel) and left the
    per-split partials in the mailbox; here they are summed."""

    @staticmethod
    def forward(ctx, mb, W, bias):
        W, bias = W.contiguous(), bias.contiguous()
        _chk(W, bias)
        L = _lib.lib()
        G = W.shape[0] * W.shape[1]
        nbytes = L.gptst_hypertem_wfrag_bytes(G)
        wf = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
        wb = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
        _lib.check(L.gptst_hypertem_pack_w(_p(W), _p(wf), _p(wb), G, _stream()), "gptst_hypertem_pack_w")
        mb["wf"], mb["wb"], mb["stream_w"] = wf, wb, torch.cuda.current_stream()
        ctx.mb = mb
        return W.view_as(W), bias.view_as(bias)

    @staticmethod
    def backward(ctx, _gW, _gb):
        mb = ctx.mb
        dWp, dbp = mb.pop("dWp"), mb.pop("dbp")
        dW, db = sum_partials(dWp, dbp)
        return None, dW, db


class _HyperTemParamsM(torch.autograd.Function):
    """Mn -> Mn on the stream where the mix matrices are produced; the backward computes the real dM_n there from what the
    block's backward left in the mailbox (dret, eb), independent of the weight-side chain."""

    @staticmethod
    def forward(ctx, mb, Mn):
        Mn = Mn.contiguous()
        _chk(Mn)
        mb["stream_m"] = torch.cuda.current_stream()
        ctx.mb = mb
        return Mn.view_as(Mn)

    @staticmethod
    def backward(ctx, _gM):
        mb = ctx.mb
        dret, eb = mb.pop("dret"), mb.pop("eb")
        B, T, N, D = eb.shape
        L = _lib.lib()
        sp = L.gptst_tmix_bwd_splits(B, N)
        dMp = torch.empty((sp, N, T, T), device=eb.device, dtype=torch.float32)
        _lib.check(L.gptst_tmix_dM2(_p(dret), _p(eb), _p(dMp), B, T, N, D, sp, _stream()), "gptst_tmix_dM2")
        (dM,) = sum_partials(dMp)
        return None, dM


class _HyperTemFused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eb, Mn, W, bias, mb):
        eb, Mn, bias = eb.contiguous(), Mn.contiguous(), bias.contiguous()
        _chk(eb, Mn, bias)
        B, T, N, D = eb.shape
        L = _lib.lib()
        npad = (N + 15) // 16 * 16
        out = torch.empty_like(eb)
        # ret = M_n o eb is only kept for the side-stream dW kernel; a forward without parameter gradients skips the store
        ret = torch.empty_like(eb) if any(ctx.needs_input_grad[1:4]) else None
        mask = torch.empty((B * T, npad, 2), device=eb.device, dtype=torch.int32)
        _lib.check(L.gptst_hypertem_fwd(_p(eb), _p(Mn), _p(mb["wf"]), _p(bias), _p(out), _p(mask), _p(ret), B, T, N, D, _stream()),
                   "gptst_hypertem_fwd")
        ctx.save_for_backward(eb, Mn, ret, mask, mb["wb"])
        ctx.mb = mb
        ctx.want_params = ret is not None
        ctx.shapes = (Mn.shape, W.shape, bias.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        eb, Mn, ret, mask, wb = ctx.saved_tensors
        mb = ctx.mb
        dout = dout.contiguous()
        B, T, N, D = eb.shape
        L = _lib.lib()
        cur = torch.cuda.current_stream()
        if ctx.want_params:
            # dW_bt / db_bt only need dOut, the sign mask and ret: launch them NOW on the weight-side stream, beside the main
            # backward kernel (they used to wait for it and formed the tail of the step behind the first block's backward)
            side = mb["stream_w"]
            G = B * T
            splits = L.gptst_gproj_splits(G, N, D)
            if side != cur:
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
                for t in (dout, mask, ret):
                    t.record_stream(side)
            with torch.cuda.stream(side):
                dWp = torch.empty((splits, B, T, D, D), device=eb.device, dtype=torch.float32)
                dbp = torch.empty((splits, B, T, D), device=eb.device, dtype=torch.float32)
                _lib.check(L.gptst_hypertem_dw(_p(dout), _p(mask), _p(ret), _p(dWp), _p(dbp), B, T, N, D, mask.shape[1], splits,
                                               _stream()), "gptst_hypertem_dw")
            mb.update(dWp=dWp, dbp=dbp)
        deb = torch.empty_like(eb)
        dret = torch.empty_like(eb) if ctx.want_params else None
        _lib.check(L.gptst_hypertem_bwd(_p(dout), _p(mask), _p(Mn), _p(wb), _p(deb), _p(dret), B, T, N, D, _stream()),
                   "gptst_hypertem_bwd")
        if not ctx.want_params:
            return deb, None, None, None, None
        sm_ = mb["stream_m"]
        if sm_ != cur:
            for t in (dret, eb):
                t.record_stream(sm_)
        mb.update(dret=dret, eb=eb)
        sM, sW, sb = ctx.shapes
        dev = eb.device
        return deb, _zero_like_shape(sM, dev), _zero_like_shape(sW, dev), _zero_like_shape(sb, dev), None


def hypertem_params_w(mb, W, bias):
    """Weight-side parameter node of the fused hyperTem block (call it where W_bt / bias_bt are produced)."""
    return _HyperTemParamsW.apply(mb, W, bias)


def hypertem_params_m(mb, Mn):
    """Mix-matrix-side parameter node of the fused hyperTem block (call it where M_n is produced)."""
    return _HyperTemParamsM.apply(mb, Mn)


def hypertem_params(Mn, W, bias):
    """Both parameter-side nodes on the current stream.  Returns (Mn, W, bias, mailbox) for `hypertem_fused`."""
    mb = {}
    Mn2 = hypertem_params_m(mb, Mn)
    W2, b2 = hypertem_params_w(mb, W, bias)
    return Mn2, W2, b2, mb


def hypertem_fused(eb, Mn2, W2, b2, mb):
    return _HyperTemFused.apply(eb, Mn2, W2, b2, mb)


def hypertem_core(eb, Mn, W, bias, prec=None):
    """eb (B,T,N,D); Mn (N,T,T) = A_n^T A_n, or already expanded to (P,N,T,T) by `expand_partials`; W (B,T,D,D); bias (B,T,D)."""
    prec = default_precision() if prec is None else prec
    if Mn.dim() == 3:
        if not eb.is_cuda:
            raise RuntimeError("gptst_b200 ops need CUDA tensors (no CPU fallback)")
        if hypertem_fused_enabled(eb.shape[3], eb.shape[1], prec):
            return hypertem_fused(eb, *hypertem_params(Mn, W, bias))
        Mn = expand_partials(Mn, hypertem_partial_count(eb.shape[0], eb.shape[2], eb.shape[3]))
    return _HyperTemCore.apply(eb, Mn, W, bias, prec)


# ---------------------------------------------------------------------------------------------------
# cap core
# ---------------------------------------------------------------------------------------------------
def cap_fused_enabled(N: int, D: int, H: int, T: int, prec: int) -> bool:
    """Round-2 forward (hop in one launch, reconstruction fused into the node-adaptive projection): D = 64, N <= 256, T = 12,
    three-term split.  GPTST_B200_CAP=split selects the four-launch chain (A/B runs, other geometries use it anyway)."""
    return (D == 64 and T == 12 and prec == PREC_3XTF32 and os.environ.get("GPTST_B200_CAP", "fused") == "fused"
            and bool(_lib.lib().gptst_cap_route2_supported(N, D, H)))


def cap_pack_wn(Wn):
    """Node-adaptive weights (N, 64, 64) [in][out] -> the fragment-ordered fp16 hi/lo table `gptst_cap_recon_proj` reads (no
    gradient flows through it: the backward uses W_n itself).  Call it where W_n is produced (the block's table stream)."""
    Wn = Wn.detach().contiguous()
    _chk(Wn)
    L = _lib.lib()
    wf = torch.empty(L.gptst_hypertem_wfrag_bytes(Wn.shape[0]), dtype=torch.uint8, device=Wn.device)
    _lib.check(L.gptst_hypertem_pack_w(_p(Wn), _p(wf), None, Wn.shape[0], _stream()), "gptst_hypertem_pack_w")
    return wf


class _CapCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wp_e, bp_e, dadj, dyn_e, Wn_e, bn_e, num_route, prec, wnf):
        x, dadj = x.contiguous(), dadj.contiguous()
        B, T, N, D = x.shape
        H, HT = dadj.shape[2], dyn_e.shape[2]
        p_wn, p_dyn, p_wp = cap_partial_counts(B, T, N, D, H)
        if (Wp_e.shape[0], bp_e.shape[0], dyn_e.shape[0], Wn_e.shape[0], bn_e.shape[0]) != (p_wp, p_wp, p_dyn, p_wn, p_wn):
            raise RuntimeError("cap_core: parameter-side inputs must be expanded by cap_partial_counts(B, T, N, D, H)")
        Wp, bp, dyn, Wn, bn = (t[0].contiguous() for t in (Wp_e, bp_e, dyn_e, Wn_e, bn_e))
        _chk(x, Wp, bp, dadj, dyn, Wn, bn)
        L = _lib.lib()
        st = _stream()
        _count(2)  # route_fwd + hop_e1 + recon_hop share `st`
        c = torch.empty((B, T, H, N), device=x.device, dtype=torch.float32)
        s = torch.empty((B, T, H, D), device=x.device, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad[:7])
        # training on the second-generation kernels: the routing forward also stores Z = x Wp^T + bp, the backward reads it instead
        # of recomputing it (these kernels are issue / shared-memory bound, not HBM bound: one more activation is the cheaper side)
        zst = None
        if need_grad and L.gptst_cap_route2_supported(N, D, H) and os.environ.get("GPTST_B200_CAP_Z", "store") == "store":
            zst = torch.empty_like(x)
            _lib.check(L.gptst_cap_route_fwd_z(_p(x), _p(Wp), _p(bp), _p(dadj), _p(c), _p(s), _p(zst), B, T, N, D, H, int(num_route),
                                               prec, st), "gptst_cap_route_fwd_z")
        else:
            _lib.check(L.gptst_cap_route_fwd(_p(x), _p(Wp), _p(bp), _p(dadj), _p(c), _p(s), B, T, N, D, H, int(num_route),
                                             prec, st), "gptst_cap_route_fwd")
        v = torch.empty_like(s)
        e1 = torch.empty((B, HT, D), device=x.device, dtype=torch.float32)
        if cap_fused_enabled(N, D, H, T, prec):
            # round 2: the hop in one launch, then reconstruction + node-adaptive projection + residual in one launch; `recon`
            # is only written because the backward reads it (never in inference)
            if wnf is None:
                wnf = cap_pack_wn(Wn)
            _lib.check(L.gptst_cap_hop_ev(_p(s), _p(dyn), _p(e1), _p(v), B, T, D, H, HT, st), "gptst_cap_hop_ev")
            recon = torch.empty_like(x) if need_grad else None
            out = torch.empty_like(x)
            _lib.check(L.gptst_cap_recon_proj(_p(c), _p(v), _p(x), _p(wnf), _p(bn), _p(out), _p(recon), B, T, N, D, H, st),
                       "gptst_cap_recon_proj")
        else:
            recon = torch.empty_like(x)
            if os.environ.get("GPTST_B200_HOP", "split") == "fused":
                # opt-in A/B variant: hop_e1 folded into recon_hop (one launch).  Measured SLOWER on the B200 (80 us vs 6.5 + 22 us:
                # every slab CTA recomputes its sample's E1 out of L2), so the split pair stays the default.
                _count(-1)
                _lib.check(L.gptst_cap_recon_hop_fused(_p(c), _p(s), _p(dyn), _p(e1), _p(v), _p(recon), B, T, N, D, H, HT, st),
                           "gptst_cap_recon_hop_fused")
            else:
                _lib.check(L.gptst_cap_hop_e1(_p(s), _p(dyn), _p(e1), B, T, D, H, HT, st), "gptst_cap_hop_e1")
                _lib.check(L.gptst_cap_recon_hop(_p(c), _p(s), _p(dyn), _p(e1), _p(v), _p(recon), B, T, N, D, H, HT, st),
                           "gptst_cap_recon_hop")
            out = gproj_fwd(recon, Wn, bn, x, node_grouped=True, act=True, prec=prec)
        if need_grad:
            ctx.save_for_backward(x, Wp, bp, dyn, Wn, c, s, v, recon, out, e1, *([zst] if zst is not None else []))
        ctx.prec = prec
        ctx.mark_non_differentiable(c)
        return out, c

    @staticmethod
    def backward(ctx, dout, _dc):
        x, Wp, bp, dyn, Wn, c, s, v, recon, out, e1 = ctx.saved_tensors[:11]
        zst = ctx.saved_tensors[11] if len(ctx.saved_tensors) > 11 else None
        B, T, N, D = x.shape
        H, HT = c.shape[2], dyn.shape[1]
        L = _lib.lib()
        st = _stream()
        _count(3)  # dv_dcr + hop_bwd2 (two launches) + route_bwd share `st`
        dout = dout.contiguous()
        drecon, dWn_part, dbn_part, dx = gproj_bwd(dout, out, recon, Wn, node_grouped=True, act=True, prec=ctx.prec, want_dres=True,
                                                   sum_parts=False)
        dcr = torch.empty_like(c)
        ds = torch.empty_like(s)
        dr_tmp, dp2_tmp = torch.empty_like(s), torch.empty_like(s)
        ddyn_part = torch.empty((L.gptst_cap_hop_bwd_parts(D),) + tuple(dyn.shape), device=x.device, dtype=torch.float32)
        if L.gptst_cap_route2_supported(N, D, H):
            # dv/dcr with the row pass of the hop backward in its tail, then the column pass
            _count(-1)
            _lib.check(L.gptst_cap_dv_dcr_hoprows(_p(c), _p(v), _p(drecon), _p(s), _p(dyn), _p(e1), _p(dcr), _p(dr_tmp), _p(dp2_tmp), B, T,
                                                  N, D, H, HT, st), "gptst_cap_dv_dcr_hoprows")
            _lib.check(L.gptst_cap_hop_bwd_cols(_p(s), _p(dyn), _p(e1), _p(dr_tmp), _p(dp2_tmp), _p(ds), _p(ddyn_part), B, T, D, H, HT,
                                                st), "gptst_cap_hop_bwd_cols")
        else:
            dv = torch.empty_like(s)
            _lib.check(L.gptst_cap_dv_dcr(_p(c), _p(v), _p(drecon), _p(dv), _p(dcr), B, T, N, D, H, st), "gptst_cap_dv_dcr")
            _lib.check(L.gptst_cap_hop_bwd2(_p(s), _p(dyn), _p(e1), _p(dv), _p(dr_tmp), _p(dp2_tmp), _p(ds), _p(ddyn_part), B, T, D,
                                            H, HT, st), "gptst_cap_hop_bwd2")
        ddadj = torch.empty_like(c)
        if L.gptst_cap_route2_supported(N, D, H):
            # second generation: dZ per slab, then the shared-weight contractions as one linear-layer backward
            dZ = torch.empty_like(x)
            if zst is not None:
                _lib.check(L.gptst_cap_route_bwd_dz_z(_p(zst), _p(c), _p(ds), _p(dcr), _p(dZ), _p(ddadj), B, T, N, D, H, ctx.prec, st),
                           "gptst_cap_route_bwd_dz_z")
            else:
                _lib.check(L.gptst_cap_route_bwd_dz(_p(x), _p(Wp), _p(bp), _p(c), _p(ds), _p(dcr), _p(dZ), _p(ddadj), B, T, N, D, H,
                                                    ctx.prec, st), "gptst_cap_route_bwd_dz")
            rows = B * T * N
            parts = L.gptst_linear_bwd_acc_splits(rows, D)
            dWp_part = torch.empty((parts, D, D), device=x.device, dtype=torch.float32)
            dbp_part = torch.empty((parts, D), device=x.device, dtype=torch.float32)
            _count(1)
            _lib.check(L.gptst_linear_bwd_acc(_p(dZ), _p(x), _p(Wp), _p(dx), _p(dWp_part), _p(dbp_part), rows, D, ctx.prec, parts,
                                              st), "gptst_linear_bwd_acc")
        else:
            parts = L.gptst_cap_route_bwd_parts(B, T, N, D, H)
            dWp_part = torch.empty((parts, D, D), device=x.device, dtype=torch.float32)
            dbp_part = torch.empty((parts, D), device=x.device, dtype=torch.float32)
            _lib.check(L.gptst_cap_route_bwd(_p(x), _p(Wp), _p(bp), _p(c), _p(ds), _p(dcr), _p(dx), _p(ddadj), _p(dWp_part),
                                             _p(dbp_part), B, T, N, D, H, ctx.prec, st), "gptst_cap_route_bwd")
        # the five partial buffers leave as they are: they are the gradients of the EXPANDED inputs, summed by the expands' own
        # backward on the stream that made them (see `expand_partials`)
        return dx, dWp_part, dbp_part, ddadj, ddyn_part, dWn_part, dbn_part, None, None, None


def cap_expand(Wp, bp, dyn, Wn, bn, B, T, N, D, H):
    """The five parameter-side inputs of `cap_core` expanded for this geometry (call it where those tensors are produced)."""
    p_wn, p_dyn, p_wp = cap_partial_counts(B, T, N, D, H)
    return expand_partials_many((Wp, bp, dyn, Wn, bn), (p_wp, p_wp, p_dyn, p_wn, p_wn))


def cap_core(x, Wp, bp, dadj, dyn, Wn, bn, num_route, prec=None, expanded=False, wnf=None):
    """x (B,T,N,D); Wp (D,D) [out,in]; dadj (B,T,H,N); dyn (B,HT,T*H); Wn (N,D,D); bn (N,D) -- or, with expanded=True, Wp, bp,
    dyn, Wn, bn as returned by `cap_expand`; wnf: `cap_pack_wn(Wn)` when the caller packed it ahead of time (else packed here).
    Returns (out (B,T,N,D), c (B,T,H,N) non-differentiable)."""
    if not expanded:
        if not x.is_cuda:
            raise RuntimeError("gptst_b200 ops need CUDA tensors (no CPU fallback)")
        B, T, N, D = x.shape
        Wp, bp, dyn, Wn, bn = cap_expand(Wp, bp, dyn, Wn, bn, B, T, N, D, dadj.shape[2])
    return _CapCore.apply(x, Wp, bp, dadj, dyn, Wn, bn, num_route, default_precision() if prec is None else prec, wnf)


# ---------------------------------------------------------------------------------------------------
# adaptive projection (MLP_RL stages):  y = LReLU(x . W_g + b_g)
# ---------------------------------------------------------------------------------------------------
class _AdaptiveProj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, node_grouped, prec):
        x, W, b = x.contiguous(), W.contiguous(), b.contiguous()
        y = gproj_fwd(x, W, b, None, node_grouped=node_grouped, act=True, prec=prec)
        ctx.save_for_backward(x, W, y)
        ctx.node_grouped, ctx.prec = node_grouped, prec
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dX, dW, db, _ = gproj_bwd(dy.contiguous(), y, x, W, node_grouped=ctx.node_grouped, act=True, prec=ctx.prec,
                                  want_dres=False)
        return dX, dW.view_as(W), db.view(W.shape[:-2] + (W.shape[-1],)), None, None


class _SharedLinear(torch.autograd.Function):
    """y = x W^T + b with ONE shared (D, D) weight in nn.Linear layout, through the grouped-projection kernels with a single
    group (rows = every cell).  Used by the eval-path fusion gate (reference model/Model.py:5-18)."""

    @staticmethod
    def forward(ctx, x, weight, bias, prec):
        x = x.contiguous()
        Wt = weight.t().contiguous().unsqueeze(0)              # [in][out], one group
        D = x.shape[-1]
        y = gproj_fwd(x.view(1, 1, -1, D), Wt, bias.contiguous().view(1, D), None, node_grouped=False, act=False, prec=prec)
        ctx.save_for_backward(x, Wt)
        ctx.prec = prec
        return y.view(x.shape[:-1] + (D,))

    @staticmethod
    def backward(ctx, dy):
        x, Wt = ctx.saved_tensors
        D = x.shape[-1]
        dX, dW, db, _ = gproj_bwd(dy.contiguous().view(1, 1, -1, D), None, x.view(1, 1, -1, D), Wt, node_grouped=False, act=False,
                                  prec=ctx.prec, want_dres=False)
        return dX.view_as(x), dW.view(D, D).t(), db.view(D), None


def shared_linear(x, weight, bias, prec=None):
    """nn.Linear(D, D) on (..., D) CUDA tensors through the sm_100a projection kernels (D = 64 or 128)."""
    return _SharedLinear.apply(x, weight, bias, default_precision() if prec is None else prec)


def node_adaptive_proj(x, Wn, bn, prec=None):
    """y[b,t,n,:] = LReLU(x[b,t,n,:] . Wn[n] + bn[n])      GPTST.py:24-27"""
    return _AdaptiveProj.apply(x, Wn, bn, True, default_precision() if prec is None else prec)


def time_adaptive_proj(x, Wt, bt, prec=None):
    """y[b,t,n,:] = LReLU(x[b,t,n,:] . Wt[b,t] + bt[b,t])  GPTST.py:29-32"""
    return _AdaptiveProj.apply(x, Wt, bt, False, default_precision() if prec is None else prec)


# ---------------------------------------------------------------------------------------------------
# small parameter-side ops (time-embedding MLP, one-feature input embedding)
# ---------------------------------------------------------------------------------------------------
class _TimeMLP(torch.autograd.Function):
    """time_feature / time_feature_spg (GPTST.py:187-219) as one forward and one backward kernel.  ab: (2, R, F) contiguous
    (the day / week inputs), weights as the nn.Linear modules hold them."""

    @staticmethod
    def forward(ctx, ab, Wd, bd, Ww, bw, W1, b1, W2, b2, W3, b3):
        ab = ab.contiguous()
        ps = [t.contiguous() for t in (Wd, bd, Ww, bw, W1, b1, W2, b2, W3, b3)]
        _chk(ab, *ps)
        _, R, F = ab.shape
        e = W1.shape[0]
        dev = ab.device
        h0, z1, z2, out = (torch.empty((R, e), device=dev, dtype=torch.float32) for _ in range(4))
        a, b = ab[0], ab[1]
        rc = _lib.lib().gptst_time_mlp_fwd(_p(a), _p(b), *[_p(t) for t in ps], _p(h0), _p(z1), _p(z2), _p(out), R, F, e, F, _stream())
        _lib.check(rc, "gptst_time_mlp_fwd")
        ctx.save_for_backward(ab, ps[4], ps[6], ps[8], h0, z1, z2)
        return out

    @staticmethod
    def backward(ctx, g):
        ab, W1, W2, W3, h0, z1, z2 = ctx.saved_tensors
        _, R, F = ab.shape
        e = W1.shape[0]
        L = _lib.lib()
        chunks, nf = L.gptst_time_mlp_chunks(R), L.gptst_time_mlp_grad_floats(e, F)
        part = torch.empty((chunks, nf), device=ab.device, dtype=torch.float32)
        rc = L.gptst_time_mlp_bwd(_p(ab[0]), _p(ab[1]), _p(W1), _p(W2), _p(W3), _p(h0), _p(z1), _p(z2), _p(g.contiguous()), _p(part), R, F,
                                  e, F, _stream())
        _lib.check(rc, "gptst_time_mlp_bwd")
        tot = part[0] if chunks == 1 else part.sum(0)
        sizes = [e * e, e, e * e, e, e * e, e, e * F, e, e * F, e]
        dW3, db3, dW2, db2, dW1, db1, dWd, dbd, dWw, dbw = torch.split(tot, sizes)
        return (None, dWd.view(e, F), dbd, dWw.view(e, F), dbw, dW1.view(e, e), db1, dW2.view(e, e), db2, dW3.view(e, e), db3)


def time_mlp(ab, mod):
    """mod: a time_feature module (ln_day, ln_week, ln1, ln2, ln); ab (2, R, F)."""
    return _TimeMLP.apply(ab, mod.ln_day.weight, mod.ln_day.bias, mod.ln_week.weight, mod.ln_week.bias, mod.ln1.weight, mod.ln1.bias,
                          mod.ln2.weight, mod.ln2.bias, mod.ln.weight, mod.ln.bias)


class _LowRankTable(torch.autograd.Function):
    """Tab = te . pool with te (G, d), pool (d, C), d <= 16: the generators of every adaptive table of the model.  Forward is
    a streaming kernel (the K = 16 product is write-bound); the backward pair of skinny products runs fused on the tensor
    cores (big tables) or as two streaming kernels, instead of cuBLAS SIMT split-K GEMMs."""

    @staticmethod
    def forward(ctx, te, pool):
        te, pool = te.contiguous(), pool.contiguous()
        _chk(te, pool)
        ctx.save_for_backward(te, pool)
        G, d = te.shape
        C = pool.shape[1]
        tab = torch.empty((G, C), device=te.device, dtype=torch.float32)
        _lib.check(_lib.lib().gptst_table_fwd(_p(te), _p(pool), _p(tab), G, d, C, _stream()), "gptst_table_fwd")
        return tab

    @staticmethod
    def backward(ctx, dtab):
        te, pool = ctx.saved_tensors
        te, pool, dtab = te.contiguous(), pool.contiguous(), dtab.contiguous()
        _chk(te, pool, dtab)
        G, d = te.shape
        C = pool.shape[1]
        L = _lib.lib()
        st = _stream()
        if G * C >= 65536:
            # big tables: one fused tensor-core pass over dTab, partials per 64-row / 128-column chunk
            rc, cc = (G + 63) // 64, (C + 127) // 128
            dpool_part = torch.empty((rc, 16, C), device=te.device, dtype=torch.float32)
            dte_part = torch.empty((cc, G, 16), device=te.device, dtype=torch.float32)
            _lib.check(L.gptst_table_bwd2(_p(te), _p(pool), _p(dtab), _p(dpool_part), _p(dte_part), G, d, C, st), "gptst_table_bwd2")
            dpool, dte = sum_partials(dpool_part, dte_part)
            dpool, dte = dpool[:d], dte[:, :d]
            return dte, dpool
        dpool = torch.empty_like(pool) if ctx.needs_input_grad[1] else None
        dte = torch.empty_like(te) if ctx.needs_input_grad[0] else None
        if dpool is not None and dte is not None:
            _count(1)
        _lib.check(L.gptst_table_bwd(_p(te), _p(pool), _p(dtab), _p(dpool), _p(dte), G, d, C, st), "gptst_table_bwd")
        return dte, dpool


class _MixMatrix(torch.autograd.Function):
    """M_n = A_n^T A_n for A (N, Ht, T) (hyperTem's two hops without a nonlinearity in between, GPTST.py:157-158)."""

    @staticmethod
    def forward(ctx, A):
        A = A.contiguous()
        _chk(A)
        N, Ht, T = A.shape
        M = torch.empty((N, T, T), device=A.device, dtype=torch.float32)
        _lib.check(_lib.lib().gptst_mn_fwd(_p(A), _p(M), N, Ht, T, _stream()), "gptst_mn_fwd")
        ctx.save_for_backward(A)
        return M

    @staticmethod
    def backward(ctx, dM):
        (A,) = ctx.saved_tensors
        N, Ht, T = A.shape
        dA = torch.empty_like(A)
        _lib.check(_lib.lib().gptst_mn_bwd(_p(A), _p(dM.contiguous()), _p(dA), N, Ht, T, _stream()), "gptst_mn_bwd")
        return dA


def mix_matrix(A):
    return _MixMatrix.apply(A)


def lowrank_table(te, pool):
    """te (..., d) x pool (d, ...) -> (te.shape[:-1] + pool.shape[1:]); einsum('...d,d***->...***')."""
    d = te.shape[-1]
    if d > 16 or not te.is_cuda:
        return (te.reshape(-1, d) @ pool.reshape(d, -1)).view(te.shape[:-1] + pool.shape[1:])
    out = _LowRankTable.apply(te.reshape(-1, d), pool.reshape(d, -1))
    return out.view(te.shape[:-1] + pool.shape[1:])


class _ScoreHead(torch.autograd.Function):
    """prob = softmax(h W3^T + b3): forward in one kernel (critical front of the adaptive phase), backward in one kernel plus
    the partial sum (the library version needed two 187 us split-K GEMMs with K = B*T*N)."""

    @staticmethod
    def forward(ctx, h, W3, b3):
        h, W3, b3 = h.contiguous(), W3.contiguous(), b3.contiguous()
        _chk(h, W3, b3)
        D, H = h.shape[-1], W3.shape[0]
        rows = h.numel() // D
        prob = torch.empty(h.shape[:-1] + (H,), device=h.device, dtype=torch.float32)
        _lib.check(_lib.lib().gptst_score_head_fwd(_p(h), _p(W3), _p(b3), _p(prob), rows, D, H, _stream()), "gptst_score_head_fwd")
        ctx.save_for_backward(h, W3, prob)
        return prob

    @staticmethod
    def backward(ctx, dprob):
        h, W3, prob = ctx.saved_tensors
        D, H = h.shape[-1], W3.shape[0]
        rows = h.numel() // D
        L = _lib.lib()
        parts = L.gptst_score_head_bwd_parts(rows)
        part = torch.empty((parts, H * D + H), device=h.device, dtype=torch.float32)
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        _lib.check(L.gptst_score_head_bwd(_p(h), _p(W3), _p(prob), _p(dprob.contiguous()), _p(dh), _p(part), rows, D, H, _stream()),
                   "gptst_score_head_bwd")
        (tot,) = sum_partials(part)
        return dh, tot[:H * D].view(H, D), tot[H * D:]


def score_head(h, W3, b3):
    return _ScoreHead.apply(h, W3, b3)


class _Affine1(torch.autograd.Function):
    """y = x w + b for a linear layer with ONE input feature; the backward is a single pass over dy."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        w, b = weight.contiguous().view(-1), bias.contiguous()
        _chk(x, w, b)
        ctx.save_for_backward(x, weight)
        D = w.numel()
        y = torch.empty(x.shape[:-1] + (D,), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().gptst_affine1_fwd(_p(x), _p(w), _p(b), _p(y), x.numel(), D, _stream()), "gptst_affine1_fwd")
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        _chk(dy)
        D = dy.shape[-1]
        n = dy.numel() // D
        L = _lib.lib()
        parts = L.gptst_affine1_bwd_parts(n)
        part = torch.empty((parts, 2, D), device=dy.device, dtype=torch.float32)
        _lib.check(L.gptst_affine1_bwd(_p(dy), _p(x.contiguous()), _p(part), n, D, parts, _stream()), "gptst_affine1_bwd")
        (tot,) = sum_partials(part)
        dx = (dy * weight.view(-1)).sum(-1, keepdim=True) if ctx.needs_input_grad[0] else None
        return dx, tot[0].view_as(weight), tot[1]


def affine1(x, weight, bias):
    return _Affine1.apply(x, weight, bias)


class _MaskedAffine1(torch.autograd.Function):
    """Encoder input embedding (reference GPTST.py:419-421): y = dim_in_flow(where(mask == 0, fill, mask * flow)) for ONE flow
    channel, in one launch; `flow` is read in place from `source` (element stride = source.shape[-1]).  No gradient reaches the
    data; dw / db come from the stored masked input through the affine backward kernel."""

    @staticmethod
    def forward(ctx, source, mask, weight, bias, fill):
        w, b = weight.contiguous().view(-1), bias.contiguous()
        _chk(source, w, b)
        if mask.dtype != torch.int64 or not mask.is_contiguous() or mask.numel() * source.shape[-1] != source.numel():
            raise RuntimeError("masked_affine1: mask must be a contiguous int64 tensor with one entry per (b, t, n) cell")
        D = w.numel()
        n = mask.numel()
        xm = torch.empty(n, device=source.device, dtype=torch.float32)
        y = torch.empty(tuple(source.shape[:-1]) + (D,), device=source.device, dtype=torch.float32)
        _lib.check(_lib.lib().gptst_masked_affine1_fwd(_p(source), source.shape[-1], _p(mask), float(fill), _p(w), _p(b), _p(xm),
                                                       _p(y), n, D, _stream()), "gptst_masked_affine1_fwd")
        ctx.save_for_backward(xm, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        xm, weight = ctx.saved_tensors
        dy = dy.contiguous()
        _chk(dy)
        D = dy.shape[-1]
        n = dy.numel() // D
        L = _lib.lib()
        parts = L.gptst_affine1_bwd_parts(n)
        part = torch.empty((parts, 2, D), device=dy.device, dtype=torch.float32)
        _lib.check(L.gptst_affine1_bwd(_p(dy), _p(xm), _p(part), n, D, parts, _stream()), "gptst_affine1_bwd")
        (tot,) = sum_partials(part)
        return None, None, tot[0].view_as(weight), tot[1], None


def masked_affine1(source, mask, weight, bias, fill):
    """source (B,T,N,C) fp32 (channel 0 = flow), mask (B,T,N,1) int64 -> (B,T,N,D)."""
    return _MaskedAffine1.apply(source, mask, weight, bias, fill)


class _ProjOut(torch.autograd.Function):
    """y = x W^T + b for nn.Linear(D, O) with O <= 4 outputs (decoder.dim_flow_out, GPTST.py:454-458): one streaming pass over x
    forward, one backward (dX written, dW / db as per-CTA partials) instead of a GEMV and three library GEMM / reduce kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight, bias = x.contiguous(), weight.contiguous(), bias.contiguous()
        _chk(x, weight, bias)
        O, D = weight.shape
        rows = x.numel() // D
        y = torch.empty(x.shape[:-1] + (O,), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().gptst_proj_out_fwd(_p(x), _p(weight), _p(bias), _p(y), rows, D, O, _stream()), "gptst_proj_out_fwd")
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        _chk(dy)
        O, D = weight.shape
        rows = x.numel() // D
        L = _lib.lib()
        parts = L.gptst_proj_out_bwd_parts(rows)
        part = torch.empty((parts, O * D + O), device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        _lib.check(L.gptst_proj_out_bwd(_p(dy), _p(x), _p(weight), _p(dx), _p(part), rows, D, O, parts, _stream()), "gptst_proj_out_bwd")
        (tot,) = sum_partials(part)
        return dx, tot[:O * D].view(O, D), tot[O * D:]


def proj_out(x, weight, bias):
    """(..., D) -> (..., O): nn.Linear with O <= 4 output features through the sm_100a streaming kernels (D = 64 or 128)."""
    return _ProjOut.apply(x, weight, bias)


# ---------------------------------------------------------------------------------------------------
# on-device mask construction (row f1)
# ---------------------------------------------------------------------------------------------------
def mask_labels(prob):
    """prob (..., H) fp32 -> (label uint8 (n,), counts int32 (H,)): arg-max class of every cell and the class histogram."""
    prob = prob.contiguous()
    _chk(prob)
    H = prob.shape[-1]
    n = prob.numel() // H
    label = torch.empty(n, dtype=torch.uint8, device=prob.device)
    counts = torch.empty(H, dtype=torch.int32, device=prob.device)
    _lib.check(_lib.lib().gptst_mask_labels(_p(prob), _p(label), _p(counts), n, H, _stream()), "gptst_mask_labels")
    return label, counts


def mask_select_adaptive(label, counts, plan, u1, u2, i0: int, all_type: bool):
    """Phase-2 mask (GPTST.py:344-413) from the class labels, the host-made plan and the two uniform draws; (n, i0) int64."""
    n, H = label.numel(), counts.numel()
    if plan.dtype != torch.int64 or plan.numel() != H + 2 or u1.numel() != n or u2.numel() != n:
        raise RuntimeError("mask_select_adaptive: inconsistent arguments")
    final = torch.empty((n, i0), dtype=torch.int64, device=label.device)
    m_ada = torch.empty(n, dtype=torch.uint8, device=label.device)
    rc = _lib.lib().gptst_mask_select(_p(label), _p(counts), _p(plan), _p(u1.contiguous()), _p(u2.contiguous()), _p(m_ada), _p(final),
                                      n, H, i0, int(all_type), 2, _stream())
    _lib.check(rc, "gptst_mask_select")
    return final


def mask_select_random(u, k_dev):
    """Phase-1 mask (GPTST.py:316-323): zero the k largest entries of u (ties in index order); k_dev is a 1-element int64
    device tensor.  Returns (n,) int64."""
    n = u.numel()
    final = torch.empty((n, 1), dtype=torch.int64, device=u.device)
    rc = _lib.lib().gptst_mask_select(None, None, _p(k_dev), _p(u.contiguous()), None, None, _p(final), n, 1, 1, 0, 1, _stream())
    _lib.check(rc, "gptst_mask_select")
    return final.view(-1)


def mask_adaptive(prob, label_override, plan, u1, u2, i0: int, all_type: bool):
    """Phase-2 mask through the multi-CTA pipeline: (n, i0) int64, 1 = keep.  prob (..., H); label_override (n,) int or None."""
    prob = prob.contiguous()
    _chk(prob)
    H = prob.shape[-1]
    n = prob.numel() // H
    dev = prob.device
    L = _lib.lib()
    if plan.dtype != torch.int64 or plan.numel() != H + 2 or u1.numel() != n or u2.numel() != n:
        raise RuntimeError("mask_adaptive: inconsistent arguments")
    lab_in = None if label_override is None else label_override.reshape(-1).to(torch.uint8).contiguous()
    label = torch.empty(n, dtype=torch.uint8, device=dev)
    m_ada = torch.empty(n, dtype=torch.uint8, device=dev)
    ws = torch.empty(L.gptst_mask_ws_ints(), dtype=torch.int32, device=dev)
    final = torch.empty((n, i0), dtype=torch.int64, device=dev)
    st = _stream()
    _count(7)   # label + 3 passes per selection + 2 (normally idle) slow-path launches
    rc = L.gptst_mask_adaptive(_p(prob), _p(lab_in), _p(plan), _p(u1.contiguous()), _p(u2.contiguous()), _p(label), _p(m_ada), _p(ws),
                               _p(final), n, H, i0, int(all_type), st)
    _lib.check(rc, "gptst_mask_adaptive")
    return final


def mask_random(u, k_dev):
    """Phase-1 mask through the multi-CTA pipeline: zero the k largest entries of u (ties in index order); (n,) int64."""
    n = u.numel()
    L = _lib.lib()
    ws = torch.empty(L.gptst_mask_ws_ints(), dtype=torch.int32, device=u.device)
    final = torch.empty(n, dtype=torch.int64, device=u.device)
    st = _stream()
    _count(3)
    _lib.check(L.gptst_mask_random(_p(k_dev), _p(u.contiguous()), _p(ws), _p(final), n, st), "gptst_mask_random")
    return final


# ---------------------------------------------------------------------------------------------------
# fused pre-training loss (row f2): value + analytic gradients in one pass
# ---------------------------------------------------------------------------------------------------
class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow_out, prob, source, inv_mask, hs, mode, mean, std, thr, kl_w):
        o = flow_out.contiguous()
        _chk(o)
        src = source if source.is_contiguous() else source.contiguous()
        inv = inv_mask.contiguous()
        if inv.dtype != torch.int64:
            raise RuntimeError("inv_mask must be int64 (the model returns 1 - mask as int64)")
        ibd = o.shape[-1]
        cells = o.numel() // ibd
        H = hs.shape[-1] if kl_w != 0.0 else 0
        L = _lib.lib()
        st = _stream()
        _count(1)
        d_o = torch.empty_like(o)
        d_p = torch.empty_like(prob) if kl_w != 0.0 else None          # (prob is None without the KL term)
        part = torch.empty(3 * L.gptst_loss_parts(), device=o.device, dtype=torch.float32)
        out = torch.empty(3, device=o.device, dtype=torch.float32)
        pc = prob.contiguous() if kl_w != 0.0 else None
        hc = hs.contiguous() if kl_w != 0.0 else None
        _lib.check(L.gptst_pretrain_loss(_p(o), _p(src), _p(inv), _p(pc), _p(hc), _p(d_o), _p(d_p), _p(part), _p(out), cells, ibd,
                                         src.shape[-1], H, int(mode), float(mean), float(std), float(thr), float(kl_w), st),
                   "gptst_pretrain_loss")
        ctx.save_for_backward(d_o, d_p if d_p is not None else d_o)
        ctx.has_kl = kl_w != 0.0
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g, _g2):
        d_o, d_p = ctx.saved_tensors
        return d_o * g, (d_p * g) if ctx.has_kl else None, None, None, None, None, None, None, None, None


def fused_probe_loss(outs, source, use_kl: bool, kl_weight: float = 0.1):
    """mean|(o - x)*mask| (+ kl_weight * KL(sum)) -- one fused kernel pair; returns the scalar loss."""
    flow_out, _, inv_mask, prob, hs = outs
    # without the KL term the scorer is NOT part of the loss graph (BasicTrainer.py:84-88): its parameters must end the step with
    # grad None -- a zero gradient would make Adam create state and count steps for them during the random-mask phase
    return _FusedLoss.apply(flow_out, prob if use_kl else None, source, inv_mask, hs, 0, 0.0, 1.0, 0.0, kl_weight if use_kl else 0.0)[0]


def fused_mask_mae_loss(outs, source, use_kl: bool, mean: float, std: float, mask_value: float = 0.0, kl_weight: float = 0.1):
    """Reference training loss (Run.py:91-101 + BasicTrainer.py:84-86) as one fused kernel pair."""
    flow_out, _, inv_mask, prob, hs = outs
    return _FusedLoss.apply(flow_out, prob if use_kl else None, source, inv_mask, hs, 1, mean, std, mask_value,
                            kl_weight if use_kl else 0.0)[0]
