"""The drop-in's phase-2 mask (sync-free device formulation, runs on any device) against the oracle's restatement
of the reference's class-selection loop (GPTST.py:344-413): bit-identical masks for equal draws."""
import random

import pytest
import torch

from oracle import gptst_oracle as O
from util import make_cfg


@pytest.mark.parametrize("ada_type", ["all", "half"])
@pytest.mark.parametrize("epoch", [11, 50, 200, 299, 300])
def test_adaptive_mask_matches_oracle(ada_type, epoch):
    from gptst_b200.GPTST import Hypergraph_encoder
    for seed in range(6):
        cfg = make_cfg(N=13, D=64, ada_type=ada_type)
        enc = Hypergraph_encoder(cfg)
        B = 3
        n = B * 12 * 13
        probs = torch.rand(B, 12, 13, 10, generator=torch.Generator().manual_seed(seed))
        if seed == 5:
            probs[..., 3] += 5          # one dominant class: the first picked class may already exceed the budget
        label = torch.sort(probs, dim=-1, descending=True)[1][..., 0]
        src = torch.zeros(B, 12, 13, 3)
        random.seed(seed)
        torch.manual_seed(seed)
        m = enc._adaptive_mask(src, probs, epoch)
        order = list(range(10))
        random.Random(seed).shuffle(order)
        torch.manual_seed(seed)
        u1, u2 = torch.rand(n), torch.rand(n)
        _, ada, rnd = O.mask_budgets(n, 0.25, 0.5, epoch, 10, 300)
        want = O.adaptive_mask(label, order, u1, u2, ada, rnd, ada_type)
        assert torch.equal(m[..., 0], want), (ada_type, epoch, seed)
        assert int((1 - m).sum()) == int(n * 0.25)


def test_plan_override_is_used():
    from gptst_b200.GPTST import Hypergraph_encoder
    enc = Hypergraph_encoder(make_cfg(N=13, D=64))
    probs = torch.rand(2, 12, 13, 10, generator=torch.Generator().manual_seed(0))
    src = torch.zeros(2, 12, 13, 3)
    random.seed(3)
    plan = enc.mask_plan(2 * 12 * 13, 200)
    torch.manual_seed(1)
    a = enc._adaptive_mask(src, probs, 200) if False else None
    enc.plan_override = plan
    torch.manual_seed(1)
    b = enc._adaptive_mask(src, probs, 12345)          # epoch ignored: budgets come from the plan
    enc.plan_override = None
    random.seed(3)
    torch.manual_seed(1)
    c = enc._adaptive_mask(src, probs, 200)
    assert torch.equal(b, c)
