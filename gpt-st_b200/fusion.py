"""Eval-path glue of the reference's ``Enhance_model`` (model/Model.py:5-18, 106-109): the gate that
blends the pre-trained encoder output with a linear embedding of the raw flow before the downstream predictor,

    z = sigmoid(HS_fc(x) + HT_fc(y)) ;   H = output_fc(z * x + (1 - z) * y) ,      y = lin_test(flow)

with the same parameter names as the reference ``Fusion`` (``HS_fc``, ``HT_fc``, ``output_fc``), so a fine-tuned checkpoint's
``fusion.*`` entries load unchanged.  The three D x D products run through the sm_100a projection kernels (one shared-weight
group), the one-feature ``lin_test`` through the affine kernel; CUDA only, like the rest of the package.  The downstream
predictors (STGCN, ...) are the reference's own model zoo and stay out of scope."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


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
        xs = ops.shared_linear(flow_eb, self.HS_fc.weight, self.HS_fc.bias)
        xt = ops.shared_linear(time_eb, self.HT_fc.weight, self.HT_fc.bias)
        z = torch.sigmoid(xs + xt)
        h = torch.addcmul(time_eb, z, flow_eb - time_eb)           # z*x + (1-z)*y
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
