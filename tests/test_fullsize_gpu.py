"""Parity at BASELINE.json's full size (configs[1]: batch 64, N=170, T=12, D=64), where the fp64 oracle is too slow to be the
checker: size-independent properties of the path.

* sample independence: no op of the path mixes samples, so the forward of the full batch restricted to a few
  samples equals -- bit for bit -- the forward of those samples alone;
* the incidence c is a softmax over the hyperedges: rows sum to one; the mask zeroes exactly int(0.25 n) cells;
* the backward is linear in the cotangent, and every gradient scale inside the kernels is a power of two: grad(4 g) == 4 grad(g)
  exactly;
* determinism: two runs give identical bits."""
import pytest
import torch

pytestmark = pytest.mark.gpu

B, N, D, T, H, HT = 64, 170, 64, 12, 10, 16


def _cap_inputs(seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    R = lambda *s, sc=1.0: torch.randn(*s, device="cuda", generator=g) * sc
    return dict(x=R(B, T, N, D), Wp=R(D, D, sc=D ** -0.5), bp=R(D, sc=0.3), dadj=R(B, T, H, N), dyn=R(B, HT, T * H, sc=0.3),
                Wn=R(N, D, D, sc=D ** -0.5), bn=R(N, D, sc=0.3))


def _cap(i, sl=slice(None)):
    from gptst_b200 import ops
    return ops.cap_core(i["x"][sl], i["Wp"], i["bp"], i["dadj"][sl], i["dyn"][sl], i["Wn"], i["bn"], 2, 3)


def test_cap_full_size_sample_independence_and_softmax():
    i = _cap_inputs()
    with torch.no_grad():
        out, c = _cap(i)
        out2, c2 = _cap(i)
        assert torch.equal(out, out2) and torch.equal(c, c2)                       # determinism
        sub_out, sub_c = _cap(i, slice(17, 21))
        assert torch.equal(out[17:21], sub_out) and torch.equal(c[17:21], sub_c)   # sample independence, bit exact
        assert (c.sum(2) - 1).abs().max().item() < 1e-5 and c.min().item() >= 0
        assert torch.isfinite(out).all()


def test_cap_full_size_backward_is_linear_in_the_cotangent():
    i = {k: v.requires_grad_() for k, v in _cap_inputs(1).items()}
    g = torch.randn(B, T, N, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    out, _ = _cap(i)
    g1 = torch.autograd.grad(out, list(i.values()), g, retain_graph=True)
    g4 = torch.autograd.grad(out, list(i.values()), 4 * g)
    for name, a, b in zip(i, g1, g4):
        assert torch.equal(4 * a, b), name


def test_hypertem_full_size_properties():
    from gptst_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    eb = torch.randn(B, T, N, D, device="cuda", generator=g).requires_grad_()
    Mn = (torch.randn(N, T, T, device="cuda", generator=g) * 0.2).requires_grad_()
    W = (torch.randn(B, T, D, D, device="cuda", generator=g) * D ** -0.5).requires_grad_()
    b = torch.rand(B, T, D, device="cuda", generator=g).requires_grad_()
    out = ops.hypertem_core(eb, Mn, W, b, 3)
    with torch.no_grad():
        sub = ops.hypertem_core(eb[40:43], Mn, W[40:43], b[40:43], 3)
    assert torch.equal(out[40:43].detach(), sub)
    go = torch.randn(B, T, N, D, device="cuda", generator=g)
    g1 = torch.autograd.grad(out, [eb, Mn, W, b], go, retain_graph=True)
    g8 = torch.autograd.grad(out, [eb, Mn, W, b], 8 * go, retain_graph=True)
    g1b = torch.autograd.grad(out, [eb, Mn, W, b], go)
    for a, c, d in zip(g1, g8, g1b):
        assert torch.equal(8 * a, c) and torch.equal(a, d)


def test_mask_full_size_exact_count_and_determinism():
    import random
    from gptst_b200 import ops
    n = B * T * N
    g = torch.Generator(device="cuda").manual_seed(3)
    prob = torch.softmax(torch.randn(B, T, N, H, device="cuda", generator=g), -1)
    u1, u2 = torch.rand(n, device="cuda", generator=g), torch.rand(n, device="cuda", generator=g)
    order = list(range(H))
    random.Random(1).shuffle(order)
    total = int(n * 0.25)
    for ada in (0, 1, total // 3, total):
        plan = torch.tensor(order + [ada, total - ada], dtype=torch.int64, device="cuda")
        m = ops.mask_adaptive(prob, None, plan, u1, u2, 1, True)
        assert int((1 - m).sum()) == total, ada
        assert torch.equal(m, ops.mask_adaptive(prob, None, plan, u1, u2, 1, True))
        label, counts = ops.mask_labels(prob)
        assert torch.equal(m, ops.mask_select_adaptive(label, counts, plan, u1, u2, 1, True))   # pipeline == one-CTA specification
