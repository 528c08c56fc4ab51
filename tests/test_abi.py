"""The C-ABI library loads and exports every symbol include/fedg.h declares (no compute calls without a GPU), and
fails loudly -- no CPU fallback -- when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from fe_project_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "fedg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fedg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    L = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(_lib.ABI_SYMBOLS) == syms


def test_rk_tables_match_oracle():
    """Host-only ABI calls (no device needed): the library's tableaux equal the oracle's."""
    from fe_project_b200.dyncore import rk_tables
    import oracle_api
    for name in ("ERK_1s1o", "ERK_4s4o", "ERK_SSP_2s2o", "ERK_SSP_3s3o", "ERK_SSP_4s3o", "ERK_SSP_5s3o_2N2*",
                 "ERK_SSP_10s4o_2N", "IMEX_ARK232", "IMEX_ARK324"):
        a, b = rk_tables(name), oracle_api.rk_tables(name)
        for k in ("nstage", "tend_buf_size", "low_storage", "imex"):
            assert a[k] == b[k], (name, k)
        for k in ("a_ex", "b_ex", "a_im", "b_im", "sig", "gam"):
            assert np.array_equal(a[k], b[k]), (name, k)
    with pytest.raises(_lib.FedgError):
        rk_tables("ERK_NOT_A_SCHEME")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from cases import DensityCurrentCase
    case = DensityCurrentCase(p=3, NeX=2, NeY=1, NeZ=1, intrp_order=3)
    with pytest.raises(_lib.FedgError) as ei:
        case.make_driver(None)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


def test_product_does_not_touch_oracle():
    """Nothing under fe_project_b200/ may import, link or call oracle/."""
    pkg = os.path.join(ROOT, "fe_project_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f)).read()
                for pat in ("oracle_api", "libfeoracle", "fe_oracle.hpp", "import oracle", "from oracle", "oracle/",
                            "feo_"):
                    assert pat not in txt, (os.path.join(dp, f), pat)
