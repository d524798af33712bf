"""Host-side logic of gptst_b200.ops that needs no GPU: partial counts of the deferred gradient sums, the stride-0 expanded views
and their autograd contract, argument checks of the newer C-ABI entry points."""
import ctypes as C

import pytest
import torch


def test_partial_counts_follow_the_library():
    from gptst_b200 import _lib, ops
    L = _lib.lib()
    for (B, T, N, D, H) in [(64, 12, 170, 64, 10), (4, 12, 170, 64, 10), (64, 12, 207, 64, 10), (16, 12, 2048, 128, 10)]:
        p_wn, p_dyn, p_wp = ops.cap_partial_counts(B, T, N, D, H)
        assert p_wn == L.gptst_gproj_splits(N, B * T, D) >= 1
        assert p_dyn == L.gptst_cap_hop_bwd_parts(D) == D // 16
        if L.gptst_cap_route2_supported(N, D, H):
            assert p_wp == L.gptst_linear_bwd_acc_splits(B * T * N, D) >= 1
        else:
            assert p_wp == L.gptst_cap_route_bwd_parts(B, T, N, D, H) >= 1
        want = L.gptst_tmix_bwd_splits(B, N) if D == 64 else L.gptst_tmix_dM_splits(B, N)
        assert ops.hypertem_partial_count(B, N, D) == want >= 1


def test_expanded_views_cost_nothing_and_sum_their_gradient_partials(monkeypatch):
    """The blocks return (P, *shape) raw partials as the gradient of a stride-0 view; the view's backward is the sum."""
    from gptst_b200 import ops
    monkeypatch.setenv("GPTST_B200_EXPAND", "native")          # torch's own expand: runs on CPU tensors too
    w = torch.randn(5, 7, requires_grad=True)
    b = torch.randn(7, requires_grad=True)
    we, be = ops.expand_partials_many((w, b), (3, 2))
    assert we.shape == (3, 5, 7) and we.stride()[0] == 0 and we.data_ptr() == w.data_ptr()
    assert be.shape == (2, 7) and be.stride()[0] == 0
    gw, gb = torch.randn(3, 5, 7), torch.randn(2, 7)
    torch.autograd.backward([we, be], [gw, gb])
    assert torch.allclose(w.grad, gw.sum(0)) and torch.allclose(b.grad, gb.sum(0))
    monkeypatch.delenv("GPTST_B200_EXPAND")
    we2, = ops.expand_partials_many((w.detach(),), (4,))       # the fused Function's forward is device-agnostic
    assert we2.shape == (4, 5, 7) and we2.stride()[0] == 0
    assert ops.expand_partials(w.detach(), 1).shape == (1, 5, 7)


def test_blocks_reject_unexpanded_or_cpu_inputs():
    from gptst_b200 import ops
    x = torch.randn(2, 12, 5, 64)
    with pytest.raises(RuntimeError):
        ops.hypertem_core(x, torch.randn(5, 12, 12), torch.randn(2, 12, 64, 64), torch.randn(2, 12, 64))
    with pytest.raises(RuntimeError):
        ops.cap_core(x, torch.randn(64, 64), torch.randn(64), torch.randn(2, 12, 10, 5), torch.randn(2, 16, 120),
                     torch.randn(5, 64, 64), torch.randn(5, 64), 2)
    with pytest.raises(RuntimeError):
        ops.proj_out(x, torch.randn(1, 64), torch.randn(1))


def test_new_entry_points_check_their_arguments_without_a_gpu():
    from gptst_b200 import _lib
    L = _lib.lib()
    one = C.c_void_p(16)                                        # any non-null pointer: the checks come before the launch
    assert L.gptst_proj_out_fwd(None, None, None, None, 10, 64, 1, None) == -1
    assert L.gptst_proj_out_fwd(one, one, one, one, 10, 96, 1, None) == -2
    assert L.gptst_proj_out_fwd(one, one, one, one, 10, 64, 5, None) == -2
    assert L.gptst_proj_out_bwd(None, None, None, None, None, 10, 64, 1, 1, None) == -1
    assert L.gptst_proj_out_bwd(one, one, one, None, one, 10, 64, 0, 1, None) == -2
    assert L.gptst_proj_out_bwd_parts(130560) == 296 and L.gptst_proj_out_bwd_parts(10) == 1
    assert L.gptst_sum_partials(None, None, None, None, 1, None) == -1
    assert L.gptst_cap_recon_hop_fused(None, None, None, None, None, None, 1, 12, 5, 64, 10, 16, None) == -1
