"""Pin the CPU oracle (oracle/gptst_oracle.py) against outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by oracle/make_golden.py from
/root/reference/model/Pretrain_model/GPTST.py (fp32, CPU).  Tolerances: the oracle re-associates
a few contractions (no (B,T,H,N,D) outer product, fused two-hop) so results differ from the
reference by fp32 rounding only: 2e-5 abs + 2e-5 rel on values of O(1..10); masks are bit-exact.
"""
import ast
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import gptst_oracle as O

ATOL, RTOL = 2e-5, 2e-5


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, atol=ATOL, rtol=RTOL, what=""):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    lim = atol + rtol * b.abs()
    assert bool((err <= lim).all()), f"{what}: max err {err.max().item():.3e} (ref scale {b.abs().max().item():.3e})"


@pytest.fixture(scope="module")
def blocks():
    return np.load(os.path.join(GOLDEN, "blocks_small.npz"))


def test_hypertem_block(blocks):
    g = blocks
    ins = {k: T(g["ht." + k]).requires_grad_() for k in ("eb", "node_emb", "time_eb", "adj", "weights_pool", "bias_pool")}
    y = O.hypertem(**ins)
    close(y, T(g["ht.out"]), what="hyperTem out")
    y.backward(T(g["ht.gout"]))
    for k, v in ins.items():
        close(v.grad, T(g["ht.g." + k]), atol=1e-4, rtol=1e-4, what="hyperTem grad " + k)


def test_cap_block(blocks):
    g = blocks
    names = ("x", "node_emb", "time_eb_spg", "teb", "ln_p.weight", "ln_p.bias", "adj", "t_adj", "weights_spa", "bias_spa")
    ins = {k: T(g["cap." + k]).requires_grad_() for k in names}
    y, c, dyn = O.cap(ins["x"], ins["node_emb"], ins["time_eb_spg"], ins["teb"], ins["ln_p.weight"], ins["ln_p.bias"],
                      ins["adj"], ins["t_adj"], ins["weights_spa"], ins["bias_spa"], int(g["dims"][9]))
    close(y, T(g["cap.out"]), what="cap out")
    close(c, T(g["cap.c"]), what="cap c")
    close(dyn, T(g["cap.dyn"]), what="cap dyn")
    y.backward(T(g["cap.gout"]))
    for k, v in ins.items():
        close(v.grad, T(g["cap.g." + k]), atol=1e-4, rtol=1e-4, what="cap grad " + k)


def test_mlp_rl_block(blocks):
    g = blocks
    P = {"m." + k[len("mlp.p."):]: T(g[k]).requires_grad_() for k in g.files if k.startswith("mlp.p.")}
    fl, te, ne = (T(g["mlp." + k]).requires_grad_() for k in ("flow", "time_eb", "node_eb"))
    y = O.mlp_rl(fl, te, ne, P, "m.")
    close(y, T(g["mlp.out"]), what="MLP_RL out")
    y.backward(T(g["mlp.gout"]))
    close(fl.grad, T(g["mlp.g.flow"]), atol=1e-4, rtol=1e-4, what="dflow")
    close(te.grad, T(g["mlp.g.time_eb"]), atol=1e-4, rtol=1e-4, what="dtime_eb")
    close(ne.grad, T(g["mlp.g.node_eb"]), atol=1e-4, rtol=1e-4, what="dnode_eb")
    for k, v in P.items():
        close(v.grad, T(g["mlp.g.p." + k[2:]]), atol=1e-4, rtol=1e-4, what="MLP_RL grad " + k)


def load_case(name):
    g = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    cfg = types.SimpleNamespace(**ast.literal_eval(str(g["cfg"][0])))
    P = {k[3:]: T(g[k]) for k in g.files if k.startswith("sd.")}
    return g, cfg, P


def test_param_shapes_match_reference_state_dict():
    g, cfg, P = load_case("pre_phase1")
    shapes = dict(O.param_shapes(cfg))
    ref = {k: tuple(v.shape) for k, v in P.items() if "mask_template" not in k}
    assert shapes == ref
    # registration order (== named_parameters order, what Run.py's init loop and Adam see)
    assert [k for k, _ in O.param_shapes(cfg)] == [k for k in P if "mask_template" not in k]


def test_model_eval_mode():
    g, cfg, P = load_case("eval")
    outs = O.model_forward(P, cfg, T(g["source"]))
    close(outs[0], T(g["out.enc"]), what="eval encoder output")
    assert all(o is outs[0] for o in outs)


@pytest.mark.parametrize("name", ["pre_phase1", "pre_phase2_all", "pre_phase2_half", "pre_phase1_ibd2", "pre_phase2_ibd2"])
def test_model_pretrain_forward_backward(name):
    g, cfg, P = load_case(name)
    epoch = int(g["epoch"][0])
    for k in P:
        if "mask_template" not in k:
            P[k].requires_grad_()
    draws = O.Draws(T(g["draw.u1"]), [int(i) for i in g["draw.order"]] if "draw.order" in g.files else None,
                    T(g["draw.u2"]) if "draw.u2" in g.files else None)
    src = T(g["source"])
    outs = O.model_forward(P, cfg, src, epoch, draws)
    flow_out, dec, inv_mask, prob, hs1 = outs
    assert inv_mask.dtype == torch.int64
    assert torch.equal(inv_mask, T(g["out.inv_mask"])), "mask must be bit-identical"
    close(flow_out, T(g["out.flow_out"]), what="flow_out")
    close(dec, T(g["out.flow_decode"]), what="flow_decode")
    close(prob, T(g["out.prob"]), what="probability")
    close(hs1, T(g["out.hs1"]), what="HS1")
    loss = O.synthetic_loss(outs, src, epoch, cfg.change_epoch)
    close(loss, T(g["loss"]), what="loss")
    loss.backward()
    have = {k for k in P if P[k].grad is not None}
    want = {k[5:] for k in g.files if k.startswith("grad.")}
    assert have == want, (have ^ want)
    for k in sorted(want):
        close(P[k].grad, T(g["grad." + k]), atol=2e-5, rtol=2e-4, what="grad " + k)


def test_draws_replay_matches_torch_global_rng():
    """Draws.sample consumes the generator exactly like the reference's rand_like calls."""
    import random
    torch.manual_seed(7)
    a = torch.rand_like(torch.empty(100))
    b = torch.rand_like(torch.empty(100))
    d = O.Draws.sample(100, 100, 10, True, torch.Generator().manual_seed(7), random.Random(7))
    assert torch.equal(a, d.u1) and torch.equal(b, d.u2)
    random.seed(7)
    o = list(range(10))
    random.shuffle(o)
    assert o == d.class_order


def load_pems08_weights():
    w = np.load(os.path.join(GOLDEN, "pems08_weights.npz"))
    return {str(k): torch.from_numpy(w[str(k)]) for k in w["__order__"]}


def test_pems08_weights_fixture_is_the_shipped_checkpoint():
    """tests/golden/pems08_weights.npz: 159 tensors in the reference's registration order; identical to the .pth when it is here."""
    from oracle.ref_import import checkpoint_path
    P = load_pems08_weights()
    assert len(P) == 159 and sum(v.numel() for v in P.values()) == 1036579
    ck = checkpoint_path("PEMS08")
    if ck is not None:
        sd = torch.load(ck, map_location="cpu")
        assert list(sd.keys()) == list(P.keys())
        assert all(torch.equal(sd[k], P[k]) for k in sd)


@pytest.mark.parametrize("tag,thr", [("thr0", 0.0), ("thr1e-3", 0.001)])
def test_loss_oracle_matches_reference_trainer_loss(tag, thr):
    """O.masked_mae / O.kl_sum against the reference's own loss closure (Run.py:91-101 -> lib/metrics.py:11-18, KL of
    BasicTrainer.py:84-86) on the fixture generated by oracle/make_golden.py `losses`."""
    g = np.load(os.path.join(GOLDEN, "losses.npz"))
    pred, prob = T(g["pred"]).requires_grad_(), T(g["prob"]).requires_grad_()
    true, inv, hs = T(g["true"]), torch.from_numpy(g["inv_mask"]), T(g["hs"])
    mean, std = (float(v) for v in g["scaler"])
    mae = O.masked_mae(pred, true, inv, mean, std, thr)
    kl = 0.1 * O.kl_sum(prob, hs)
    assert abs(mae.item() - float(g[f"{tag}.mae"][0])) <= 1e-5 * float(g[f"{tag}.mae"][0])
    assert abs(kl.item() - float(g[f"{tag}.kl"][0])) <= 1e-5 * float(g[f"{tag}.kl"][0])
    gp, gq = torch.autograd.grad(mae + kl, [pred, prob])
    close(gp, T(g[f"{tag}.g.pred"]), atol=1e-7, rtol=1e-5, what="d loss / d pred")
    close(gq, T(g[f"{tag}.g.prob"]), atol=1e-7, rtol=1e-5, what="d loss / d prob")


def test_pems08_checkpoint_golden():
    """Shipped PEMS08 checkpoint, first 8 test windows (SURVEY.md §8c numbers)."""
    g = np.load(os.path.join(GOLDEN, "pems08_ckpt.npz"))
    P = load_pems08_weights()                      # the shipped checkpoint as a committed fixture: runs without /root/reference
    x = T(g["x"])
    assert abs(float(x[0, 0, 0, 0]) - 1.3344) < 1e-4

    def cfg(mode):
        return types.SimpleNamespace(num_nodes=170, input_base_dim=1, input_extra_dim=2, hidden_dim=64, output_dim=1,
                                     horizon=12, lag=12, embed_dim=16, embed_dim_spa=4, HS=10, HT=16, HT_Tem=8, num_route=2,
                                     mode=mode, scaler_zeros=float(g["scaler_zeros"][0]), mask_ratio=0.25,
                                     ada_mask_ratio=0.5, ada_type="all", change_epoch=10, epochs=300)

    with torch.no_grad():
        o = O.model_forward(P, cfg("eval"), x)[0]
    close(o[:, :, ::10, ::4], T(g["eval.sample"]), atol=2e-5, rtol=1e-5, what="eval encoder sample")
    mom = g["eval.moments"]
    assert abs(o.mean().item() - 0.34829369) < 1e-6 and abs(o.mean().item() - mom[0]) < 1e-6
    assert abs(o.abs().mean().item() - 0.34927550) < 1e-6
    assert abs(o.std().item() - 0.58143628) < 1e-6
    close(o[0, 0, 0, :6], torch.tensor([-0.0025578002, 0.6543606520, 0.8402512074, -0.0014474761, 0.5384114385,
                                        -0.0006552326]), atol=5e-6, rtol=0, what="survey golden o[0,0,0,:6]")
    close(o[7, 11, 169, -4:], torch.tensor([-0.0002394654, 0.2032192647, 0.3642835319, 0.0327602550]), atol=5e-6, rtol=0,
          what="survey golden o[7,11,169,-4:]")
    std = float(g["std"][0])
    for ep in (1, 300):
        dr = O.Draws(T(g[f"pre{ep}.u1"]), [int(i) for i in g[f"pre{ep}.order"]] if ep > 10 else None,
                     T(g[f"pre{ep}.u2"]) if ep > 10 else None)
        with torch.no_grad():
            fo, _, inv, prob, hs = O.model_forward(P, cfg("pretrain"), x, ep, dr)
        want = np.unpackbits(g[f"pre{ep}.inv_mask_packed"])[: inv.numel()].reshape(inv.shape)
        assert np.array_equal(inv.numpy().astype(np.uint8), want), f"epoch {ep}: mask differs from reference"
        close(fo[:, :, ::5, 0], T(g[f"pre{ep}.flow_out_sample"]), atol=5e-5, rtol=1e-5, what=f"flow_out epoch {ep}")
        m = inv.bool()
        mae = ((fo - x[..., :1]) * std).abs()[m].mean().item()
        kl = O.kl_sum(prob, hs).item()
        st = g[f"pre{ep}.stats"]
        assert int(m.sum()) == int(st[1]) == 4080
        assert abs(mae - st[0]) < 2e-3 and abs(kl - st[2]) < 2e-3 * max(1.0, abs(st[2]))


# ---------------------------------------------------------------------------------------------------
# eval path next to the encoder: reference Fusion gate + lin_test, STGCN's GLU temporal convolution
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eval_path():
    return np.load(os.path.join(GOLDEN, "eval_path.npz"))


def test_eval_glue_oracle_matches_reference_fusion(eval_path):
    g = eval_path
    D, ibd = (int(v) for v in g["glue.dims"])
    P = {k[len("glue.p."):]: T(g[k]).requires_grad_() for k in g.files if k.startswith("glue.p.")}
    x_pre = T(g["glue.x_pre"]).requires_grad_()
    y = O.eval_glue(T(g["glue.source"]), x_pre, P, ibd)
    close(y, T(g["glue.out"]), what="eval glue out")
    y.backward(T(g["glue.gout"]))
    close(x_pre.grad, T(g["glue.g.x_pre"]), atol=1e-4, rtol=1e-4, what="eval glue d x_pre")
    assert len(P) == 8
    for k, v in P.items():
        close(v.grad, T(g["glue.g." + k]), atol=1e-4, rtol=1e-4, what="eval glue grad " + k)


@pytest.mark.parametrize("tag", ["same", "narrow", "widen", "wide_kernel"])
def test_temporal_conv_glu_oracle_matches_reference_stgcn(eval_path, tag):
    g, pre = eval_path, f"glu.{tag}."
    x = T(g[pre + "x"]).requires_grad_()
    w, b = T(g[pre + "conv.weight"]).requires_grad_(), T(g[pre + "conv.bias"]).requires_grad_()
    aw = T(g[pre + "align.weight"]).requires_grad_() if pre + "align.weight" in g.files else None
    ab = T(g[pre + "align.bias"]).requires_grad_() if aw is not None else None
    y = O.temporal_conv_glu(x, w, b, aw, ab)
    close(y, T(g[pre + "out"]), what="GLU out " + tag)
    y.backward(T(g[pre + "gout"]))
    close(x.grad, T(g[pre + "g.x"]), atol=1e-4, rtol=1e-4, what="GLU dx " + tag)
    close(w.grad, T(g[pre + "g.conv.weight"]), atol=1e-4, rtol=1e-4, what="GLU dW " + tag)
    close(b.grad, T(g[pre + "g.conv.bias"]), atol=1e-4, rtol=1e-4, what="GLU db " + tag)
    if aw is not None:
        close(aw.grad, T(g[pre + "g.align.weight"]), atol=1e-4, rtol=1e-4, what="GLU d align W " + tag)
        close(ab.grad, T(g[pre + "g.align.bias"]), atol=1e-4, rtol=1e-4, what="GLU d align b " + tag)
