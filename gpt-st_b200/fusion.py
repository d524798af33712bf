"""Eval-path glue of the reference's ``Enhance_model`` (model/Model.py:5-18, 106-109): the gate that
blends the pre-trained encoder output with a linear embedding of the raw flow before the downstream predictor,

    z = sigmoid(HS_fc(x) + HT_fc(y)) ;   H = output_fc(z * x + (1 - z) * y) ,      y = lin_test(flow)

with the same parameter names as the reference ``Fusion`` (``HS_fc``, ``HT_fc``, ``output_fc``), so a fine-tuned checkpoint's
``fusion.*`` entries load unchanged.  The three D x D products run through the sm_100a projection kernels (one shared-weight
group); the sigmoid and the blend are the epilogue of the second product (``gptst_gate_fwd``, D = 64; an elementwise kernel on
the two pre-activations for other widths), the backward of the gate is one elementwise kernel followed by the two linear-layer
backward kernels; the one-feature ``lin_test`` goes through the affine kernel.  CUDA only, like the rest of the package.

``TemporalConvGLU`` is STGCN's gated temporal convolution (model/STGCN/stgcn.py:25-53, ``TemporalConvLayer(kt, c_in, c_out,
"GLU")``, with the same parameter names ``conv.*`` / ``align.conv1x1.*``) on the reference's own (B, C, T, N) layout, forward and
backward through ``csrc/glu_conv.cu``.  The rest of the downstream predictors (the reference's model zoo) stays out of scope."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .ops import _p, _stream


class _FusionGate(torch.autograd.Function):
    """h = z * x + (1 - z) * y,  z = sigmoid(x W_S^T + b_S + y W_T^T + b_T)      (model/Model.py:13-16)"""

    @staticmethod
    def forward(ctx, x, y, WS, bS, WT, bT, prec):
        x, y = x.contiguous(), y.contiguous()
        D = x.shape[-1]
        rows = x.numel() // D
        L = _lib.lib()
        xs = ops.gproj_fwd(x.view(1, 1, rows, D), WS.t().contiguous().unsqueeze(0), bS.contiguous().view(1, D), None,
                           node_grouped=False, act=False, prec=prec)
        h, z = torch.empty_like(x), torch.empty_like(x)
        WTt = WT.t().contiguous()
        rc = L.gptst_gate_fwd(_p(x), _p(y), _p(xs), _p(WTt), _p(bT.contiguous()), _p(h), _p(z), rows, D, prec, _stream())
        if rc == -2:      # width the fused epilogue does not cover: plain second product + elementwise gate
            xt = ops.gproj_fwd(y.view(1, 1, rows, D), WTt.unsqueeze(0), bT.contiguous().view(1, D), None, node_grouped=False,
                               act=False, prec=prec)
            rc = L.gptst_gate_blend(_p(xs), _p(xt), _p(x), _p(y), _p(h), _p(z), x.numel(), _stream())
        _lib.check(rc, "gptst_gate_fwd")
        ctx.save_for_backward(x, y, z, WS, WT)
        ctx.prec = prec
        return h

    @staticmethod
    def backward(ctx, dh):
        x, y, z, WS, WT = ctx.saved_tensors
        D = x.shape[-1]
        rows = x.numel() // D
        L = _lib.lib()
        dh = dh.contiguous()
        dpre, dx, dy = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        _lib.check(L.gptst_gate_bwd(_p(dh), _p(z), _p(x), _p(y), _p(dpre), _p(dx), _p(dy), x.numel(), _stream()), "gptst_gate_bwd")
        outs = []
        for inp, W, dacc in ((x, WS, dx), (y, WT, dy)):
            if D == 64:      # dX accumulated in place on top of the direct path, weight in nn.Linear layout
                parts = L.gptst_linear_bwd_acc_splits(rows, D)
                dWp = torch.empty((parts, D, D), device=x.device, dtype=torch.float32)
                dbp = torch.empty((parts, D), device=x.device, dtype=torch.float32)
                _lib.check(L.gptst_linear_bwd_acc(_p(dpre), _p(inp), _p(W.contiguous()), _p(dacc), _p(dWp), _p(dbp), rows, D, ctx.prec,
                                                  parts, _stream()), "gptst_linear_bwd_acc")
                dW, db = ops.sum_partials(dWp, dbp)
            else:
                dX, dWt, db, _ = ops.gproj_bwd(dpre.view(1, 1, rows, D), None, inp.view(1, 1, rows, D), W.t().contiguous().unsqueeze(0),
                                               node_grouped=False, act=False, prec=ctx.prec, want_dres=False)
                dacc += dX.view_as(dacc)
                dW, db = dWt.view(D, D).t(), db.view(D)
            outs.append((dacc, dW.view(D, D), db.view(D)))
        (dx, dWS, dbS), (dy, dWT, dbT) = outs
        return dx, dy, dWS, dbS, dWT, dbT, None


def fusion_gate(x, y, WS, bS, WT, bT, prec=None):
    return _FusionGate.apply(x, y, WS, bS, WT, bT, ops.default_precision() if prec is None else prec)


class _GluTConv(torch.autograd.Function):
    """out = (conv(x)[:, :Cout] + align(x)) * sigmoid(conv(x)[:, Cout:])        (model/STGCN/stgcn.py:37-48)"""

    @staticmethod
    def forward(ctx, x, W, b, aw, ab):
        x = x.contiguous()
        B, Cin, T, N = x.shape
        C2, _, kt, _ = W.shape
        Cout = C2 // 2
        L = _lib.lib()
        out = torch.empty((B, Cout, T, N), device=x.device, dtype=torch.float32)
        need = any(ctx.needs_input_grad)
        P = torch.empty_like(out) if need else None
        S = torch.empty_like(out) if need else None
        Wc = W.contiguous()
        _lib.check(L.gptst_glu_tconv_fwd(_p(x), _p(Wc), _p(b.contiguous()), _p(aw.contiguous()) if aw is not None else None,
                                         _p(ab.contiguous()) if ab is not None else None, _p(out), _p(P), _p(S), B, Cin, Cout, T, N, kt,
                                         _stream()), "gptst_glu_tconv_fwd")
        ctx.save_for_backward(x, Wc, aw, P, S)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W, aw, P, S = ctx.saved_tensors
        B, Cin, T, N = x.shape
        C2, _, kt, _ = W.shape
        Cout, pad = C2 // 2, (kt - 1) // 2
        L = _lib.lib()
        st = _stream()
        dconv = torch.empty((B, C2, T, N), device=x.device, dtype=torch.float32)
        _lib.check(L.gptst_glu_gate_bwd(_p(dout.contiguous()), _p(P), _p(S), _p(dconv), B, Cout, T, N, st), "gptst_glu_gate_bwd")
        # dx = "same" convolution of dconv with the flipped, transposed taps; the Align term sits on the centre tap of the P half
        Weff = W[..., 0].clone()                                     # (C2, Cin, kt)
        if aw is not None:
            Weff[:Cout, :, pad] += aw.reshape(Cout, Cin)
        else:
            m = min(Cin, Cout)
            Weff[:m, :m, pad] += torch.eye(m, device=x.device, dtype=torch.float32)
        Wt = Weff.flip(2).permute(1, 0, 2).contiguous()               # (Cin, C2, kt)
        dx = torch.empty_like(x)
        _lib.check(L.gptst_tconv_fwd(_p(dconv), _p(Wt), None, _p(dx), B, C2, Cin, T, N, kt, st), "gptst_tconv_fwd")
        splits = L.gptst_tconv_dw_splits(B, C2, Cin, N)
        dWp = torch.empty((splits, C2, Cin, kt), device=x.device, dtype=torch.float32)
        dbp = torch.empty((splits, C2), device=x.device, dtype=torch.float32)
        _lib.check(L.gptst_tconv_dw(_p(dconv), _p(x), _p(dWp), _p(dbp), B, C2, Cin, T, N, kt, splits, st), "gptst_tconv_dw")
        dW, db = ops.sum_partials(dWp, dbp)
        daw = dab = None
        if aw is not None:
            daw = dW[:Cout, :, pad].reshape(aw.shape).contiguous()
            dab = db[:Cout].contiguous()
        return dx, dW.unsqueeze(-1), db, daw, dab


class _Align(nn.Module):
    """Parameter container with the reference's names (stgcn.py:10-16): ``align.conv1x1.{weight,bias}`` exist iff c_in > c_out."""

    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        if c_in > c_out:
            self.conv1x1 = nn.Conv2d(c_in, c_out, 1)


class TemporalConvGLU(nn.Module):
    """``TemporalConvLayer(kt, c_in, c_out, "GLU")`` of the reference STGCN (stgcn.py:25-48); same state_dict keys."""

    def __init__(self, kt: int, c_in: int, c_out: int):
        super().__init__()
        if kt % 2 == 0:
            raise ValueError("an even kt shortens the sequence: the reference's residual add fails for it as well")
        self.kt, self.c_in, self.c_out = kt, c_in, c_out
        self.align = _Align(c_in, c_out)
        self.conv = nn.Conv2d(c_in, c_out * 2, (kt, 1), 1, padding=[(kt - 1) // 2, 0])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("gptst_b200.fusion.TemporalConvGLU runs on CUDA only (no CPU fallback)")
        aw = ab = None
        if self.c_in > self.c_out:
            aw, ab = self.align.conv1x1.weight, self.align.conv1x1.bias
        return _GluTConv.apply(x, self.conv.weight, self.conv.bias, aw, ab)


class Fusion(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        if dim not in (64, 128):
            raise ValueError("gptst_b200 kernels are built for hidden_dim 64 or 128")
        self.HS_fc = nn.Linear(dim, dim, bias=True)
        self.HT_fc = nn.Linear(dim, dim, bias=True)
        self.output_fc = nn.Linear(dim, dim, bias=True)

    def forward(self, flow_eb: torch.Tensor, time_eb: torch.Tensor) -> torch.Tensor:
        if not flow_eb.is_cuda:
            raise RuntimeError("gptst_b200.fusion.Fusion runs on CUDA only (no CPU fallback)")
        h = fusion_gate(flow_eb, time_eb, self.HS_fc.weight, self.HS_fc.bias, self.HT_fc.weight, self.HT_fc.bias)
        return ops.shared_linear(h, self.output_fc.weight, self.output_fc.bias)


class EvalGlue(nn.Module):
    """``lin_test`` + ``Fusion`` of Enhance_model.forward_pretrain (model/Model.py:106-109): returns the embedding handed to the
    predictor.  Parameter names: ``fusion.*``, ``lin_test.*`` as in the reference."""

    def __init__(self, input_base_dim: int, hidden_dim: int):
        super().__init__()
        self.input_base_dim = input_base_dim
        self.fusion = Fusion(hidden_dim)
        self.lin_test = nn.Linear(input_base_dim, hidden_dim)

    def forward(self, source: torch.Tensor, x_pretrain_flow: torch.Tensor) -> torch.Tensor:
        flow = source[..., :self.input_base_dim]
        if self.input_base_dim == 1:
            x_t1 = ops.affine1(flow, self.lin_test.weight, self.lin_test.bias)
        else:
            x_t1 = self.lin_test(flow)
        return self.fusion(x_pretrain_flow, x_t1)
