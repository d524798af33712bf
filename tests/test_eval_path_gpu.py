"""Row f4, second half: STGCN's gated temporal convolution (reference model/STGCN/stgcn.py:25-53) and the fusion
gate with its sigmoid / blend as the epilogue of the second product (model/Model.py:12-17), CUDA kernels vs

* the committed golden fixtures generated from the UNMODIFIED reference classes (tests/golden/eval_path.npz: outputs and every
  gradient of `TemporalConvLayer(kt, c_in, c_out, "GLU")` in four channel / kernel configurations, and of `Fusion` + `lin_test`),
* the fp64 oracle restatement (itself pinned by the same fixtures on CPU) at STGCN's real PEMS08 sizes.

Tolerances: the temporal convolution is plain fp32 FMA -> 2e-6 x abs-max forward, 2e-5 x abs-max on gradients (sums over
B*T*N = 130 560 positions); the gate uses the three-term fp16 split for its D x D products -> 5e-5 / 2e-4 x abs-max."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import gptst_oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu


def tol(ref, rel):
    return rel * max(1e-6, ref.detach().abs().max().item())


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "eval_path.npz"))


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _layer(kt, c_in, c_out, w, b, aw, ab):
    from gptst_b200.fusion import TemporalConvGLU
    layer = TemporalConvGLU(kt, c_in, c_out).cuda()
    sd = {"conv.weight": w, "conv.bias": b}
    if aw is not None:
        sd.update({"align.conv1x1.weight": aw, "align.conv1x1.bias": ab})
    assert set(layer.state_dict()) == set(sd), "state_dict keys must be the reference's"
    layer.load_state_dict(sd, strict=True)
    return layer


@pytest.mark.parametrize("tag", ["same", "narrow", "widen", "wide_kernel"])
def test_glu_tconv_matches_reference_golden(gold, tag):
    pre = f"glu.{tag}."
    w, b = T(gold[pre + "conv.weight"]), T(gold[pre + "conv.bias"])
    aw = T(gold[pre + "align.weight"]) if pre + "align.weight" in gold.files else None
    ab = T(gold[pre + "align.bias"]) if aw is not None else None
    x = T(gold[pre + "x"]).cuda().requires_grad_()
    layer = _layer(w.shape[2], w.shape[1], w.shape[0] // 2, w, b, aw, ab)
    y = layer(x)
    want = T(gold[pre + "out"])
    assert_close(y, want, atol=tol(want, 2e-6), rtol=0, what="GLU out " + tag)
    y.backward(T(gold[pre + "gout"]).cuda())
    pairs = [(x.grad, "g.x"), (layer.conv.weight.grad, "g.conv.weight"), (layer.conv.bias.grad, "g.conv.bias")]
    if aw is not None:
        pairs += [(layer.align.conv1x1.weight.grad, "g.align.weight"), (layer.align.conv1x1.bias.grad, "g.align.bias")]
    for got, key in pairs:
        ref = T(gold[pre + key])
        assert_close(got, ref, atol=tol(ref, 2e-5), rtol=0, what=f"GLU {key} {tag}")


@pytest.mark.parametrize("B,c_in,c_out,N", [(4, 64, 32, 170), (2, 128, 128, 170), (2, 32, 64, 207), (64, 64, 32, 170)])
def test_glu_tconv_matches_oracle_at_stgcn_sizes(B, c_in, c_out, N):
    """st_conv*.tconv1 (64 -> 32) and output.tconv1 (128 -> 128) of the PEMS08 STGCN (conf/STGCN/PEMS08.conf: Kt = 3), fp64 oracle."""
    kt, Tn = 3, 12
    g = torch.Generator().manual_seed(5 + c_in + B)
    w = (torch.randn(2 * c_out, c_in, kt, 1, generator=g) / (c_in * kt) ** 0.5)
    b = torch.randn(2 * c_out, generator=g) * 0.1
    aw = torch.randn(c_out, c_in, 1, 1, generator=g) / c_in ** 0.5 if c_in > c_out else None
    ab = torch.randn(c_out, generator=g) * 0.1 if c_in > c_out else None
    x = torch.randn(B, c_in, Tn, N, generator=g)
    gout = torch.randn(B, c_out, Tn, N, generator=g)
    xr = x.double().requires_grad_()
    P = [t.double().requires_grad_() if t is not None else None for t in (w, b, aw, ab)]
    want = O.temporal_conv_glu(xr, *P)
    want.backward(gout.double())
    layer = _layer(kt, c_in, c_out, w, b, aw, ab)
    xc = x.cuda().requires_grad_()
    got = layer(xc)
    assert_close(got, want, atol=tol(want, 2e-6), rtol=0, what="GLU out")
    got.backward(gout.cuda())
    assert_close(xc.grad, xr.grad, atol=tol(xr.grad, 5e-6), rtol=0, what="GLU dx")
    assert_close(layer.conv.weight.grad, P[0].grad, atol=tol(P[0].grad, 2e-5), rtol=0, what="GLU dW")
    assert_close(layer.conv.bias.grad, P[1].grad, atol=tol(P[1].grad, 2e-5), rtol=0, what="GLU db")
    if aw is not None:
        assert_close(layer.align.conv1x1.weight.grad, P[2].grad, atol=tol(P[2].grad, 2e-5), rtol=0, what="GLU d align W")
        assert_close(layer.align.conv1x1.bias.grad, P[3].grad, atol=tol(P[3].grad, 2e-5), rtol=0, what="GLU d align b")


def test_glu_tconv_rejects_what_the_reference_rejects():
    from gptst_b200.fusion import TemporalConvGLU
    with pytest.raises(ValueError):
        TemporalConvGLU(2, 8, 8)                       # even kt: conv output is one step short of align(x) in the reference as well
    layer = TemporalConvGLU(3, 8, 8)
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 8, 12, 5))                # CPU tensor: no fallback


@pytest.mark.parametrize("D", [64, 128])
def test_fusion_gate_fused_epilogue_matches_oracle(D):
    """z = sigmoid(HS_fc(x) + HT_fc(y)), h = z x + (1 - z) y, output_fc(h): D = 64 takes the fused epilogue (gptst_gate_fwd), D = 128
    the elementwise gate; both vs the fp64 oracle (O.fusion_gate, pinned to the reference Fusion by tests/golden/eval_path.npz)."""
    from gptst_b200.fusion import Fusion
    torch.manual_seed(21)
    fus = Fusion(D).cuda()
    P = {"fusion." + k: v.detach().double().cpu().requires_grad_() for k, v in fus.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x, y = torch.randn(2, 12, 170, D, generator=g), torch.randn(2, 12, 170, D, generator=g)
    gout = torch.randn(2, 12, 170, D, generator=g)
    xr, yr = x.double().requires_grad_(), y.double().requires_grad_()
    want = O.fusion_gate(xr, yr, P, "fusion.")
    want.backward(gout.double())
    xc, yc = x.cuda().requires_grad_(), y.cuda().requires_grad_()
    got = fus(xc, yc)
    assert_close(got, want, atol=tol(want, 5e-5), rtol=0, what="gate out")
    got.backward(gout.cuda())
    assert_close(xc.grad, xr.grad, atol=tol(xr.grad, 2e-4), rtol=0, what="gate d flow")
    assert_close(yc.grad, yr.grad, atol=tol(yr.grad, 2e-4), rtol=0, what="gate d time")
    for k, p in fus.named_parameters():
        assert_close(p.grad, P["fusion." + k].grad, atol=tol(P["fusion." + k].grad, 2e-4), rtol=0, what="gate grad " + k)


def test_eval_glue_matches_reference_golden(gold):
    """lin_test + Fusion of Enhance_model.forward_pretrain (Model.py:106-109) at D = 16 -- a width the projection kernels do not
    cover: the module must refuse it loudly rather than fall back."""
    from gptst_b200.fusion import EvalGlue
    D, ibd = (int(v) for v in gold["glue.dims"])
    with pytest.raises(ValueError):
        EvalGlue(ibd, D)
