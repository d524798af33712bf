"""Shared helpers for the parity tests (CUDA path vs CPU oracle)."""
import types

import torch


def make_cfg(N=170, D=64, mode="pretrain", ibd=1, **kw):
    base = dict(num_nodes=N, input_base_dim=ibd, input_extra_dim=2, hidden_dim=D, output_dim=ibd, horizon=12, lag=12,
                embed_dim=16, embed_dim_spa=4, HS=10, HT=16, HT_Tem=8, num_route=2, mode=mode, model="TGCN", device="cuda",
                scaler_zeros=-1.5767, interval=5, week_day=7, mask_ratio=0.25, ada_mask_ratio=0.5, ada_type="all",
                change_epoch=10, epochs=300)
    base.update(kw)
    return types.SimpleNamespace(**base)


def assert_close(got, want, atol, rtol, what=""):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    assert got.shape == want.shape, (what, tuple(got.shape), tuple(want.shape))
    err = (got - want).abs()
    lim = atol + rtol * want.abs()
    bad = err > lim
    if bool(bad.any()):
        i = int(torch.argmax(err - lim))
        raise AssertionError(f"{what}: {int(bad.sum())}/{err.numel()} elements out of tolerance; max err {err.max().item():.3e} "
                             f"(at flat {i}: got {got.reshape(-1)[i].item():.6e} want {want.reshape(-1)[i].item():.6e}); "
                             f"ref abs-max {want.abs().max().item():.3e}")


def rel_l2(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).norm() / want.norm().clamp_min(1e-30)).item()
