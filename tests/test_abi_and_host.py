"""CPU: the C-ABI library loads and exports every symbol include/cup3d_b200.h
declares (no compute calls: there is no GPU here), fails loudly without a
device, and the host-side mesh helpers reproduce the reference's block order."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import ALL_CASES, case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cup3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cup_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built):
    from cup3d_b200 import capi
    names = header_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in capi.SYMBOLS, "not bound in capi.py: " + n
    assert sorted(capi.SYMBOLS) == names


def test_no_cpu_fallback(built):
    """without a CUDA device the product must refuse to run, not fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cup3d_b200
    with pytest.raises(cup3d_b200.CupError):
        cup3d_b200.Context(0, 8)


def test_product_does_not_use_oracle():
    """oracle/ is test infrastructure: nothing under cup3d_b200/ may import, link or load it"""
    bad = ("import oracle", "from oracle", "cup_oracle", "libcup3d_ref", "refbind", "portbind")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cup3d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                for b in bad:
                    assert b not in txt, (f, b)


@pytest.mark.parametrize("name", ALL_CASES)
def test_hilbert_block_order_matches_reference(name):
    """mesh.uniform_blocks orders blocks exactly like the reference's blk_sort (main.c:2807)"""
    from cup3d_b200 import mesh
    c = case(name)
    lvl = int(c.ib[0, 0])
    if not all(b == c.bpd[0] and (b & (b - 1)) == 0 for b in c.bpd):
        # non-cubic / non-power-of-two bases use the reference's Zsave remap: set equality only
        ib, rb = mesh.uniform_blocks(lvl, c.bpd)
        assert sorted(map(tuple, ib)) == sorted(map(tuple, c.ib))
        return
    ib, rb = mesh.uniform_blocks(lvl, c.bpd)
    assert np.array_equal(ib, c.ib)
    assert np.allclose(rb, c.rb, rtol=0, atol=1e-15)


def test_blocks_struct_layout():
    """CupBlk must mirror struct Blk (main.c:59-63): 4 ints, long long, 4 doubles = 56 bytes"""
    from cup3d_b200 import capi
    assert ctypes.sizeof(capi.CupBlk) == 56
    assert capi.CupBlk.h.offset == 24 and capi.CupBlk.origin.offset == 32
