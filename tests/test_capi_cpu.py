"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, and the Python module mirrors the reference's state_dict contract."""
import copy
import os
import re

import pytest
import torch

from conftest import REPO
from oracle import gptst_oracle as O
from util import make_cfg


def header_symbols():
    src = open(os.path.join(REPO, "include", "gptst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gptst_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gptst_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/gptst_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert b"sm_100a" in L.gptst_version()


def test_pure_host_entry_points():
    from gptst_b200 import _lib
    L = _lib.lib()
    assert L.gptst_gproj_splits(768, 170, 64) == 1          # one CTA per (b,t) group: deterministic dW
    assert L.gptst_gproj_splits(170, 768, 64) >= 2
    assert L.gptst_tmix_dM_splits(64, 170) >= 1
    assert L.gptst_cap_route_bwd_parts(64, 12, 170, 64, 10) >= 148


def test_argument_checks_return_codes_without_gpu():
    from gptst_b200 import _lib
    L = _lib.lib()
    assert L.gptst_tmix(None, None, None, 1, 12, 1, 64, 0, 0, None) == -1
    assert L.gptst_gproj_fwd(None, None, None, None, None, 1, 1, 1, 1, 64, 1, 3, None) == -1
    assert L.gptst_cap_route_fwd(None, None, None, None, None, None, 1, 12, 1, 64, 10, 2, 3, None) == -1
    with pytest.raises(_lib.GptstLibraryError):
        _lib.check(-2, "x")


def test_state_dict_contract_matches_reference_layout():
    from gptst_b200.GPTST import GPTST_Model
    cfg = make_cfg(N=170, D=64)
    m = GPTST_Model(cfg)
    names = [k for k, _ in m.named_parameters()]
    want = O.param_shapes(cfg)
    assert names == [k for k, _ in want]
    assert {k: tuple(v.shape) for k, v in m.named_parameters()} == dict(want)
    sd = m.state_dict()
    assert len(sd) == 159 and sum(v.numel() for v in sd.values()) == 1036579
    assert [k for k in sd if "mask_template" in k] == [
        "encoder.STHCN_encode.cap1.mask_template", "encoder.STHCN_encode.cap2.mask_template",
        "decoder.STHCN_decode.cap1.mask_template", "decoder.STHCN_decode.cap2.mask_template"]
    for p in m.parameters():                                  # Run.py:79-85 (the pools are registered uninitialised)
        torch.nn.init.xavier_uniform_(p) if p.dim() > 1 else torch.nn.init.uniform_(p)
    m2 = copy.deepcopy(m)                                     # BasicTrainer.py:179-180 deep-copies the model
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_shipped_checkpoint_loads_strict():
    from gptst_b200.GPTST import GPTST_Model
    from oracle.ref_import import checkpoint_path
    ck = checkpoint_path("PEMS08")
    if ck is None:
        pytest.skip("reference checkpoint not available")
    m = GPTST_Model(make_cfg(N=170, D=64, mode="eval"))
    sd = torch.load(ck, map_location="cpu")
    assert list(sd.keys()) == list(m.state_dict().keys())     # same order as the reference's state_dict
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_no_cpu_fallback():
    from gptst_b200.GPTST import GPTST_Model
    m = GPTST_Model(make_cfg(N=7, D=64, mode="eval"))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 12, 7, 3), None)
