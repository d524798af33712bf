"""Losses of the pre-training step as the reference trainer applies them.

    masked MAE   Run.py:91-101 + lib/metrics.py:11-18   (inverse z-score, x int mask, keep true > thresh, mean |.|)
    KL           Run.py:132 + BasicTrainer.py:84-86     (0.1 * KLDivLoss(sum)(prob.log(), HS1), epoch > change_epoch)
    probe loss   the synthetic-data loss of the driver's GPU probe

Plain PyTorch on purpose: O(B*T*N) work on (B,T,N,1) / (B,T,N,H) tensors; the heavy blocks are in ops.py.
"""
from __future__ import annotations

import torch


def kl_sum(prob: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.kl_div(prob.log(), target, reduction="sum")


def masked_mae(pred, true, mask, mean: float, std: float, mask_value=0.0):
    p = (pred * std + mean) * mask
    t = (true * std + mean) * mask
    if mask_value is not None:
        sel = t > mask_value
        p, t = torch.masked_select(p, sel), torch.masked_select(t, sel)
    return (t - p).abs().mean()


def pretrain_loss(outs, source, epoch: int, change_epoch: int, mean: float, std: float, output_dim: int = 1,
                  mask_value: float = 0.0):
    flow_out, _, inv_mask, prob, hs1 = outs
    loss = masked_mae(flow_out, source[..., :output_dim], inv_mask, mean, std, mask_value)
    if epoch > change_epoch:
        loss = loss + 0.1 * kl_sum(prob, hs1)
    return loss


def probe_loss(outs, source, epoch: int, change_epoch: int = 10):
    o, _, inv_mask, prob, hs = outs
    loss = ((o - source[..., :o.shape[-1]]) * inv_mask).abs().mean()
    if epoch > change_epoch:
        loss = loss + 0.1 * kl_sum(prob, hs)
    return loss


def masked_mae_syncfree(pred, true, mask, mean: float, std: float, mask_value=0.0):
    """Same value as masked_mae without the data-dependent `masked_select` (no host sync, CUDA-graph safe):
    mean over {true*mask > thresh} of |true - pred| == sum(|.| * sel) / sum(sel)."""
    p = (pred * std + mean) * mask
    t = (true * std + mean) * mask
    if mask_value is None:
        return (t - p).abs().mean()
    sel = (t > mask_value).to(p.dtype)
    return ((t - p).abs() * sel).sum() / sel.sum()


def pretrain_loss_syncfree(outs, source, epoch: int, change_epoch: int, mean: float, std: float, output_dim: int = 1,
                           mask_value: float = 0.0):
    """mask_value = args.mape_thresh of the reference (Run.py:116-122: 0.0 for PEMS08 / METR_LA, 0.001 for NYC_BIKE / NYC_TAXI)."""
    flow_out, _, inv_mask, prob, hs1 = outs
    loss = masked_mae_syncfree(flow_out, source[..., :output_dim], inv_mask, mean, std, mask_value)
    if epoch > change_epoch:
        loss = loss + 0.1 * kl_sum(prob, hs1)
    return loss
