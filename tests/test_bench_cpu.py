"""Host logic of bench.py and of this round's new C entry points (no GPU): workload partitioning, identical config dicts on both
arms, the reference arm's behaviour on non-zero ranks, argument checks of the new entry points."""
import argparse
import ctypes as C
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_weak_and_strong_partitioning():
    assert bench.per_gpu_batch("pems08", 1) == 64 and bench.per_gpu_batch("pems08", 8) == 64          # weak: 64 per GPU
    assert bench.per_gpu_batch("metr_la", 4) == 64
    assert [bench.per_gpu_batch("synthetic2048", w) for w in (1, 2, 4, 8)] == [128, 64, 32, 16]       # strong: 128 global
    with pytest.raises(SystemExit):
        bench.per_gpu_batch("synthetic2048", 3)
    assert bench.per_gpu_batch("pems08", 2, override=16) == 16


def test_config_dict_is_the_same_function_of_the_arguments_on_both_arms():
    a = bench.workload_config("pems08", 64, 8, 200)
    b = bench.workload_config("pems08", 64, 8, 200)
    assert a == b and a["global_batch"] == 512 and a["per_gpu_batch"] == 64 and a["parallelism"] == "dp8"
    assert a["num_nodes"] == 170 and a["hidden_dim"] == 64 and a["mask_phase"] == "adaptive+KL"
    assert "configs[1]" in a["workload"] and "model" not in a
    assert bench.workload_config("synthetic2048", 16, 8, 1)["mask_phase"] == "random"


def test_reference_arm_is_silent_on_non_zero_ranks(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "1")
    args = argparse.Namespace(workload="pems08", gpus=2, batch=0, epoch=200, steps=1, warmup=0)
    assert bench.run_reference_arm(args) == 0
    assert capsys.readouterr().out == ""


def test_round2_entry_points_check_their_arguments_without_a_gpu():
    from gptst_b200 import _lib
    L = _lib.lib()
    one = C.c_void_p(16)                                        # any non-null pointer: the checks come before the launch
    # routing forward that stores Z / backward that reads it (GPTST.py:102-123)
    assert L.gptst_cap_route_fwd_z(None, None, None, None, None, None, None, 2, 12, 30, 64, 10, 2, 3, None) == -1
    assert L.gptst_cap_route_fwd_z(one, one, one, one, one, one, one, 2, 12, 300, 64, 10, 2, 3, None) == -2     # N > 256
    assert L.gptst_cap_route_fwd_z(one, one, one, one, one, one, one, 2, 12, 30, 128, 10, 2, 3, None) == -2     # D != 64
    assert L.gptst_cap_route_bwd_dz_z(None, None, None, None, None, None, 2, 12, 30, 64, 10, 3, None) == -1
    assert L.gptst_cap_route_bwd_dz_z(one, one, one, one, one, one, 2, 12, 30, 64, 10, 2, None) == -2           # precision 2
    # masked input embedding (GPTST.py:419-421)
    assert L.gptst_masked_affine1_fwd(None, 3, None, C.c_float(0.0), None, None, None, None, 10, 64, None) == -1
    assert L.gptst_masked_affine1_fwd(one, 3, one, C.c_float(0.0), one, one, one, one, 10, 6, None) == -2       # D % 4
    assert L.gptst_masked_affine1_fwd(one, 0, one, C.c_float(0.0), one, one, one, one, 10, 64, None) == -1      # stride
    # partial sums: more than 8 segments are rejected, mixed calls are legal
    assert L.gptst_sum_partials(one, one, one, one, 9, None) == -2
