"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

What is produced (all fp32 CPU results of the reference's own classes, imported by
oracle/ref_import.py; the reference ships no fixtures of its own):

  blocks_small.npz     reference ``hyperTem`` / ``cap`` / ``MLP_RL`` modules on seeded random inputs at
                       tiny dims, with outputs and every input/parameter gradient for a fixed cotangent.
  model_<case>.npz     whole ``GPTST_Model`` (state_dict + input + the random draws it consumed) with the
                       five forward outputs, the probe loss and every parameter gradient, for
                       eval / pretrain phase 1 / pretrain phase 2 ('all' and 'half', ibd 1 and 2).
  pems08_ckpt.npz      shipped PEMS08 checkpoint on the first 8 PEMS08 test windows: the input windows,
                       a strided sample + moments of the eval-mode encoder output, and the pretrain-mode
                       masked MAE / KL at epoch 1 and 300 (the SURVEY.md §8c numbers).
  pems08_weights.npz   the shipped PEMS08 checkpoint itself (model/SAVE/PEMS08/GPTST_ada.pth, 159 tensors, a reference DATA artefact)
                       as a plain npz, so that the checkpoint parity tests run on a box without /root/reference.
  losses.npz           the trainer's loss (Run.py:91-101 `scaler_mae_loss` closure extracted with ast + lib/metrics.py MAE_torch +
                       lib/normalization.py StandardScaler; KL of Run.py:132 / BasicTrainer.py:84-86) on seeded inputs, with
                       values and gradients, for mask_value (args.mape_thresh) 0.0 and 0.001.
  eval_path.npz        eval-path pieces next to the encoder: the reference ``Fusion`` gate + ``lin_test``
                       and STGCN's ``TemporalConvLayer`` (GLU) in four channel / kernel configurations, with gradients.
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from oracle import gptst_oracle as O  # noqa: E402
from oracle.ref_import import checkpoint_path, load_reference, reference_root  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def small_cfg(**kw):
    base = dict(num_nodes=9, input_base_dim=1, input_extra_dim=2, hidden_dim=16, output_dim=1, horizon=12, lag=12,
                embed_dim=4, embed_dim_spa=3, HS=5, HT=6, HT_Tem=4, num_route=2, mode="pretrain", model="TGCN",
                device="cpu", scaler_zeros=-1.5767, interval=5, week_day=7, mask_ratio=0.25, ada_mask_ratio=0.5,
                ada_type="all", change_epoch=10, epochs=300)
    base.update(kw)
    return types.SimpleNamespace(**base)


def run_init(model, seed):
    torch.manual_seed(seed)
    for p in model.parameters():  # Run.py:79-85
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)
        else:
            torch.nn.init.uniform_(p)


def npd(t):
    return t.detach().cpu().numpy()


def blocks(ref):
    torch.manual_seed(123)
    B, T, N, D, d, ds, H, HT, Ht = 3, 12, 9, 16, 4, 3, 5, 6, 4
    out = {"dims": np.array([B, T, N, D, d, ds, H, HT, Ht, 2])}
    # ---- hyperTem
    m = ref.hyperTem(T, N, D, D, d, Ht)
    run_init(m, 1)
    eb = torch.randn(B, T, N, D, requires_grad=True)
    ne = torch.randn(N, d, requires_grad=True)
    te = torch.randn(B, T, d, requires_grad=True)
    y = m(eb, ne, te)
    g = torch.randn_like(y)
    y.backward(g)
    out.update({"ht.eb": npd(eb), "ht.node_emb": npd(ne), "ht.time_eb": npd(te), "ht.adj": npd(m.adj),
                "ht.weights_pool": npd(m.weights_pool), "ht.bias_pool": npd(m.bias_pool), "ht.out": npd(y),
                "ht.gout": npd(g), "ht.g.eb": npd(eb.grad), "ht.g.node_emb": npd(ne.grad),
                "ht.g.time_eb": npd(te.grad), "ht.g.adj": npd(m.adj.grad),
                "ht.g.weights_pool": npd(m.weights_pool.grad), "ht.g.bias_pool": npd(m.bias_pool.grad)})
    # ---- cap
    m = ref.cap(D, N, T, d, ds, H, HT, 2)
    run_init(m, 2)
    x = torch.randn(B, T, N, D, requires_grad=True)
    ne = torch.randn(N, d, requires_grad=True)
    tes = torch.randn(B, ds, requires_grad=True)
    teb = torch.randn(B, T, ds, requires_grad=True)
    y, c, dyn = m(x, ne, tes, teb)
    g = torch.randn_like(y)
    y.backward(g)
    out.update({"cap.x": npd(x), "cap.node_emb": npd(ne), "cap.time_eb_spg": npd(tes), "cap.teb": npd(teb),
                "cap.ln_p.weight": npd(m.ln_p.weight), "cap.ln_p.bias": npd(m.ln_p.bias), "cap.adj": npd(m.adj),
                "cap.t_adj": npd(m.t_adj), "cap.weights_spa": npd(m.weights_spa), "cap.bias_spa": npd(m.bias_spa),
                "cap.out": npd(y), "cap.c": npd(c.squeeze(-1)), "cap.dyn": npd(dyn), "cap.gout": npd(g),
                "cap.g.x": npd(x.grad), "cap.g.node_emb": npd(ne.grad), "cap.g.time_eb_spg": npd(tes.grad),
                "cap.g.teb": npd(teb.grad), "cap.g.ln_p.weight": npd(m.ln_p.weight.grad),
                "cap.g.ln_p.bias": npd(m.ln_p.bias.grad), "cap.g.adj": npd(m.adj.grad), "cap.g.t_adj": npd(m.t_adj.grad),
                "cap.g.weights_spa": npd(m.weights_spa.grad), "cap.g.bias_spa": npd(m.bias_spa.grad)})
    # ---- MLP_RL
    m = ref.MLP_RL(1, H, D, d, "cpu")
    run_init(m, 3)
    fl = torch.randn(B, T, N, 1, requires_grad=True)
    te = torch.randn(B, T, d, requires_grad=True)
    ne = torch.randn(N, d, requires_grad=True)
    y = m(fl, te, ne)
    g = torch.randn_like(y)
    y.backward(g)
    out.update({"mlp.flow": npd(fl), "mlp.time_eb": npd(te), "mlp.node_eb": npd(ne), "mlp.out": npd(y), "mlp.gout": npd(g),
                "mlp.g.flow": npd(fl.grad), "mlp.g.time_eb": npd(te.grad), "mlp.g.node_eb": npd(ne.grad)})
    for k, p in m.named_parameters():
        out["mlp.p." + k] = npd(p)
        out["mlp.g.p." + k] = npd(p.grad)
    np.savez_compressed(os.path.join(GOLD, "blocks_small.npz"), **out)
    print("blocks_small.npz", sum(v.size for v in out.values()), "elements")


def synth_source(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, cfg.horizon, cfg.num_nodes, cfg.input_base_dim + 2, generator=g)


def model_case(ref, name, cfg, B, epoch, seed):
    model = ref.GPTST_Model(cfg)
    run_init(model, seed)
    src = synth_source(cfg, B, seed + 1)
    out = {"cfg": np.array([repr(vars(cfg))]), "source": npd(src), "epoch": np.array([epoch if epoch else -1])}
    for k, v in model.state_dict().items():
        out["sd." + k] = npd(v)
    torch.manual_seed(seed + 2)
    random.seed(seed + 2)
    res = model(src, src, 1, epoch)
    if cfg.mode == "pretrain":
        # replay the draws the reference consumed (same generator algorithm, same order; see Draws.sample)
        gen = torch.Generator().manual_seed(seed + 2)
        n = B * cfg.horizon * cfg.num_nodes
        dr = O.Draws.sample(n * cfg.input_base_dim, n, cfg.HS, epoch > cfg.change_epoch, gen, random.Random(seed + 2))
        out["draw.u1"] = npd(dr.u1)
        if dr.u2 is not None:
            out["draw.u2"] = npd(dr.u2)
            out["draw.order"] = np.array(dr.class_order)
        loss = O.synthetic_loss(res, src, epoch, cfg.change_epoch)
        loss.backward()
        out["loss"] = npd(loss)
        for k, p in model.named_parameters():
            if p.grad is not None:
                out["grad." + k] = npd(p.grad)
        names = ["flow_out", "flow_decode", "inv_mask", "prob", "hs1"]
        for nme, t in zip(names, res):
            out["out." + nme] = npd(t)
    else:
        out["out.enc"] = npd(res[0])
    np.savez_compressed(os.path.join(GOLD, f"model_{name}.npz"), **out)
    print(f"model_{name}.npz", sum(v.size for v in out.values()), "elements")


def pems08_windows(root):
    """First 8 windows of the PEMS08 test split, through the reference's own lib/ pipeline."""
    npz = None
    for r in (root, os.path.join(REPO, "baseline", "_ref")):
        p = os.path.join(r, "data", "PEMS08", "PEMS08.npz")
        if os.path.isfile(p):
            npz = p
            break
    if npz is None:
        return None
    sys.path.insert(0, root)
    cwd = os.getcwd()
    os.chdir(os.path.join(os.path.dirname(os.path.dirname(npz)), "..", "model"))
    try:
        from lib.dataloader import get_dataloader
        a = types.SimpleNamespace(dataset="PEMS08", val_ratio=0.2, test_ratio=0.2, lag=12, horizon=12, input_base_dim=1,
                                  column_wise=False, batch_size=8)
        _, _, te, sc, _, _, _ = get_dataloader(a, normalizer="std", tod=False, dow=False, weather=False, single=False)
        x = next(iter(te))[0][:8, ..., :3].cpu()
        return x, float(sc.mean), float(sc.std), a.interval, a.week_day
    finally:
        os.chdir(cwd)


def pems08(ref, root):
    ck = checkpoint_path("PEMS08")
    got = pems08_windows(root)
    if ck is None or got is None:
        print("pems08: checkpoint or dataset missing, skipped")
        return
    x, mean, std, interval, week_day = got
    zeros = (0 - mean) / std
    sd = torch.load(ck, map_location="cpu")
    out = {"x": npd(x), "mean": np.array([mean]), "std": np.array([std]), "scaler_zeros": np.array([zeros])}

    def cfg(mode):
        return types.SimpleNamespace(num_nodes=170, input_base_dim=1, input_extra_dim=2, hidden_dim=64, output_dim=1,
                                     horizon=12, lag=12, embed_dim=16, embed_dim_spa=4, HS=10, HT=16, HT_Tem=8, num_route=2,
                                     mode=mode, model="STGCN", device="cpu", scaler_zeros=zeros, interval=interval,
                                     week_day=week_day, mask_ratio=0.25, ada_mask_ratio=0.5, ada_type="all",
                                     change_epoch=10, epochs=300)

    m = ref.GPTST_Model(cfg("eval"))
    print(m.load_state_dict(sd, strict=True))
    with torch.no_grad():
        o = m(x, x)[0]
    out["eval.sample"] = npd(o[:, :, ::10, ::4])
    out["eval.moments"] = np.array([o.mean().item(), o.abs().mean().item(), o.std().item(), o.abs().max().item()])
    out["eval.first6"] = npd(o[0, 0, 0, :6])
    out["eval.last4"] = npd(o[7, 11, 169, -4:])
    print("eval moments", out["eval.moments"], out["eval.first6"])
    m = ref.GPTST_Model(cfg("pretrain"))
    m.load_state_dict(sd, strict=True)
    m.eval()
    for ep in (1, 300):
        torch.manual_seed(12)
        random.seed(12)
        with torch.no_grad():
            fo, _, inv, prob, hs = m(x, x, None, ep)
        gen = torch.Generator().manual_seed(12)
        n = x.shape[0] * 12 * 170
        dr = O.Draws.sample(n, n, 10, ep > 10, gen, random.Random(12))
        out[f"pre{ep}.u1"] = npd(dr.u1)
        if dr.u2 is not None:
            out[f"pre{ep}.u2"] = npd(dr.u2)
            out[f"pre{ep}.order"] = np.array(dr.class_order)
        msk = inv.bool()
        mae = ((fo - x[..., :1]) * std).abs()[msk].mean().item()
        kl = torch.nn.KLDivLoss(reduction="sum")(prob.log(), hs).item()
        out[f"pre{ep}.stats"] = np.array([mae, int(msk.sum()), kl])
        out[f"pre{ep}.inv_mask_packed"] = np.packbits(npd(inv).astype(np.uint8).reshape(-1))
        out[f"pre{ep}.flow_out_sample"] = npd(fo[:, :, ::5, 0])
        print(f"pretrain epoch {ep}: masked MAE {mae:.5f} cells {int(msk.sum())} KL {kl:.4f}")
    np.savez_compressed(os.path.join(GOLD, "pems08_ckpt.npz"), **out)
    print("pems08_ckpt.npz", sum(v.size for v in out.values()), "elements")


def pems08_weights():
    ck = checkpoint_path("PEMS08")
    if ck is None:
        print("pems08_weights: checkpoint missing, skipped")
        return
    sd = torch.load(ck, map_location="cpu")
    out = {"__order__": np.array(list(sd.keys()))}
    out.update({k: npd(v) for k, v in sd.items()})
    np.savez_compressed(os.path.join(GOLD, "pems08_weights.npz"), **out)
    print("pems08_weights.npz", len(sd), "tensors", sum(v.numel() for v in sd.values()), "elements")


def losses(root):
    """Run.py's loss closure + KL on seeded (B,T,N,1) tensors (values in z-score space, as the trainer passes them)."""
    import ast
    import importlib.util

    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    metrics = load(os.path.join(root, "lib", "metrics.py"), "ref_metrics")
    norm = load(os.path.join(root, "lib", "normalization.py"), "ref_norm")
    src = open(os.path.join(root, "model", "Run.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "scaler_mae_loss")
    ns = {"MAE_torch": metrics.MAE_torch, "args": types.SimpleNamespace(mode="pretrain"), "torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "Run.py:scaler_mae_loss", "exec"), ns)
    mean, std = 229.64618311390265, 145.64368009977818                       # PEMS08 scaler
    scaler = norm.StandardScaler(mean, std)
    torch.manual_seed(91)
    B, T, N, H = 3, 12, 11, 5
    pred = torch.randn(B, T, N, 1, requires_grad=True)
    true = torch.randn(B, T, N, 1) * 1.2
    true[0, :, 3] = -mean / std                                               # exact zeros after the inverse transform (sensor gaps)
    true[1, :, 4] = (0.0005 - mean) / std                                     # 0 < value <= 0.001: kept by thresh 0, dropped by 0.001
    inv_mask = (torch.rand(B, T, N, 1) < 0.25).long()
    inv_mask[1, :, 4] = 1
    inv_mask[0, 0:6, 3] = 1
    prob = torch.softmax(torch.randn(B, T, N, H), -1).requires_grad_()
    hs = torch.softmax(torch.randn(B, T, N, H), -1)
    out = {"pred": npd(pred), "true": npd(true), "inv_mask": npd(inv_mask), "prob": npd(prob), "hs": npd(hs),
           "scaler": np.array([mean, std])}
    for tag, thr in (("thr0", 0.0), ("thr1e-3", 0.001)):
        loss_fn = ns["scaler_mae_loss"](scaler, thr)
        mae, _ = loss_fn(pred, true, inv_mask)                                # BasicTrainer.py:83
        kl = torch.nn.KLDivLoss(reduction="sum")(prob.log(), hs) * 0.1        # BasicTrainer.py:85
        tot = mae + kl
        gp, gq = torch.autograd.grad(tot, [pred, prob])
        out.update({f"{tag}.mae": np.array([mae.item()]), f"{tag}.kl": np.array([kl.item()]), f"{tag}.g.pred": npd(gp),
                    f"{tag}.g.prob": npd(gq)})
        print(f"losses {tag}: mae {mae.item():.6f} kl {kl.item():.6f}")
    np.savez_compressed(os.path.join(GOLD, "losses.npz"), **out)


def eval_path(root):
    """Fusion (model/Model.py:5-18, class source extracted with ast: the module itself imports the whole predictor zoo) and
    STGCN's TemporalConvLayer with GLU (model/STGCN/stgcn.py:25-53, imported as is) on seeded inputs, with all gradients."""
    import ast
    import importlib.util
    src = open(os.path.join(root, "model", "Model.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "Fusion")
    ns = {"torch": torch, "nn": torch.nn}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "Model.py:Fusion", "exec"), ns)
    out = {}
    torch.manual_seed(70)
    D, ibd = 16, 1
    fus = ns["Fusion"](D)
    lin = torch.nn.Linear(ibd, D)
    source = torch.randn(3, 12, 9, ibd + 2)
    x_pre = torch.randn(3, 12, 9, D, requires_grad=True)
    y = fus(x_pre, lin(source[..., :ibd]))                    # Model.py:106-109
    g = torch.randn_like(y)
    y.backward(g)
    out.update({"glue.source": npd(source), "glue.x_pre": npd(x_pre), "glue.out": npd(y), "glue.gout": npd(g),
                "glue.g.x_pre": npd(x_pre.grad), "glue.dims": np.array([D, ibd])})
    for k, v in list(fus.named_parameters()) + [("lin_test." + k, v) for k, v in lin.named_parameters()]:
        name = k if k.startswith("lin_test.") else "fusion." + k
        out["glue.p." + name] = npd(v)
        out["glue.g." + name] = npd(v.grad)
    spec = importlib.util.spec_from_file_location("ref_stgcn", os.path.join(root, "model", "STGCN", "stgcn.py"))
    st = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st)
    for tag, (kt, c_in, c_out, T) in {"same": (3, 8, 8, 12), "narrow": (3, 12, 6, 12), "widen": (3, 4, 8, 12), "wide_kernel": (5, 8, 8, 12)}.items():
        torch.manual_seed(80 + kt + c_in)
        layer = st.TemporalConvLayer(kt, c_in, c_out, "GLU")
        x = torch.randn(2, c_in, T, 7, requires_grad=True)
        y = layer(x)
        g = torch.randn_like(y)
        y.backward(g)
        pre = f"glu.{tag}."
        out.update({pre + "x": npd(x), pre + "out": npd(y), pre + "gout": npd(g), pre + "g.x": npd(x.grad),
                    pre + "conv.weight": npd(layer.conv.weight), pre + "conv.bias": npd(layer.conv.bias),
                    pre + "g.conv.weight": npd(layer.conv.weight.grad), pre + "g.conv.bias": npd(layer.conv.bias.grad)})
        if c_in > c_out:
            out.update({pre + "align.weight": npd(layer.align.conv1x1.weight), pre + "align.bias": npd(layer.align.conv1x1.bias),
                        pre + "g.align.weight": npd(layer.align.conv1x1.weight.grad), pre + "g.align.bias": npd(layer.align.conv1x1.bias.grad)})
    np.savez_compressed(os.path.join(GOLD, "eval_path.npz"), **out)
    print("eval_path.npz", sum(v.size for v in out.values()), "elements")


def main():
    os.makedirs(GOLD, exist_ok=True)
    root = reference_root()
    ref = load_reference("cpu", root)
    torch.set_num_threads(4)
    blocks(ref)
    model_case(ref, "eval", small_cfg(mode="eval"), 3, None, 10)
    model_case(ref, "pre_phase1", small_cfg(), 3, 1, 20)
    model_case(ref, "pre_phase2_all", small_cfg(), 3, 200, 30)
    model_case(ref, "pre_phase2_half", small_cfg(ada_type="half"), 3, 299, 40)
    model_case(ref, "pre_phase1_ibd2", small_cfg(input_base_dim=2), 2, 5, 50)
    model_case(ref, "pre_phase2_ibd2", small_cfg(input_base_dim=2), 2, 120, 60)
    pems08(ref, root)
    pems08_weights()
    losses(root)
    eval_path(root)


if __name__ == "__main__":
    main()
