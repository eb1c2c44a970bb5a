"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, its host
tables equal the oracle's, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import carta1_b200
from carta1_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = carta1_b200.load()
    hdr = open(os.path.join(ROOT, "include", "carta1_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(carta1_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert getattr(L, name) is not None
    assert L.carta1_abi_version() == 1


def test_default_tables_equal_oracle(oracle):
    t = carta1_b200.default_tables()
    o = oracle.default_tables()
    for name, _ in _lib.Tables._fields_:
        a = np.ctypeslib.as_array(getattr(t, name)).view(np.uint64)
        b = np.ctypeslib.as_array(getattr(o, name)).view(np.uint64)
        assert np.array_equal(a, b), name


def test_aea_header_roundtrip(oracle):
    h = carta1_b200.aea_write_header("encoded by carta1", 1724, 2)
    assert np.array_equal(h, oracle.aea_header("encoded by carta1", 1724, 2))
    assert carta1_b200.aea_parse_header(h) == ("encoded by carta1", 1724, 2)
    long_title = "x" * 400
    h = carta1_b200.aea_write_header(long_title, 1, 1)
    assert carta1_b200.aea_parse_header(h)[0] == "x" * 255
    with pytest.raises(ValueError, match="Header must be 2048 bytes"):
        carta1_b200.aea_parse_header(np.zeros(5, np.uint8))
    bad = h.copy()
    bad[1] = 9
    with pytest.raises(ValueError, match="Invalid AEA file"):
        carta1_b200.aea_parse_header(bad)


def test_frame_count():
    L = carta1_b200.load()
    assert [L.carta1_frame_count(n) for n in (0, 1, 512, 513, 1280)] == [0, 1, 1, 2, 3]


def test_no_cpu_fallback():
    """Without a CUDA device the context cannot be created and says why."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(carta1_b200.Carta1Error, match="needs a CUDA device"):
        carta1_b200.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "carta1_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", ".js", ".mjs")):
                with open(os.path.join(dirpath, f), errors="replace") as fh:
                    src = fh.read().lower()
                assert "oracle" not in src and "refpin" not in src, os.path.join(dirpath, f)
