"""The C-ABI library loads and exports every symbol include/r2p2_b200.h declares; the product path fails loudly
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import r2p2_b200
from r2p2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "r2p2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(r2p2_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export " + n
        assert n in _lib.EXPORTS, "python binding missing for " + n
    assert sorted(_lib.EXPORTS) == names


def test_version_and_loud_failure_without_gpu():
    L = r2p2_b200.lib()
    assert L.r2p2_version() >= 100
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(r2p2_b200.R2P2Error) as ei:
        r2p2_b200.BatchSim(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under r2p2_b200/ may import, load or include it."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|libr2p2_oracle|orc_run)")
    for dp, _, files in os.walk(os.path.join(ROOT, "r2p2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dp, f)).read()), (dp, f)
