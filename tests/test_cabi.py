"""The C-ABI library loads on a CPU-only box and exports every symbol include/fisr_b200.h declares (no compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "fisr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fisr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from fisr_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fisr_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_param_inventory_matches_oracle():
    import fisr_b200
    from oracle import fisrnet_oracle as O
    inv = fisr_b200.param_inventory()
    ref = O.init_params(0)
    assert list(inv) == list(ref)
    assert all(tuple(ref[k].shape) == inv[k] for k in inv)
    assert sum(int(np.prod(s)) for s in inv.values()) == 48_316_251


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fisr_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.fisr_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU path" in lib.fisr_last_error(None)
    import fisr_b200
    with pytest.raises(fisr_b200.FisrError):
        fisr_b200.Engine(0)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fisr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
