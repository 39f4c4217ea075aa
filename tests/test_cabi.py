"""The C-ABI library loads and exports every symbol include/jacobi_b200.h declares; without a GPU the product path
fails loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "jacobi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jf_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    g.build()
    from jacobiforcing_b200 import _cabi
    lib = _cabi.load()
    syms = declared_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/jacobi_b200.h but not exported"
    assert sorted(_cabi.EXPORTS) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jacobiforcing_b200.runtime import JacobiRuntime
    from oracle.hf_model import make_config
    with pytest.raises(RuntimeError, match="no CUDA device"):
        JacobiRuntime(make_config("tiny"))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jacobiforcing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} reaches into oracle/"


def test_knobs_struct_matches_policy_header():
    from jacobiforcing_b200 import _cabi
    from jacobiforcing_b200.policy import MODES, PolicyKnobs
    kn = PolicyKnobs(mode=MODES["multiblock"], n=32, K=2, r=0.85, n_gram_pool_size=4, pad=0)
    ints = kn.as_ints()
    assert len(ints) == len(_cabi.PolicyKnobsC._fields_)
    assert ints[3] == 28                       # ceil(0.85 * 32), unified.py:262
    assert PolicyKnobs(mode=1, n=64, r=0.85).as_ints()[3] == 55
    assert PolicyKnobs(mode=1, n=8, lookahead_start_ratio=0.5).as_ints()[10] == 4
    with pytest.raises(NotImplementedError):
        PolicyKnobs(mode=1, n=8, K=3).validate()
    with pytest.raises(ValueError):
        PolicyKnobs(mode=2, n=1).validate()
