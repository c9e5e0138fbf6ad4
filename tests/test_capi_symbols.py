"""CPU tests: the C-ABI library loads and exports every symbol include/fftop_b200.h declares, the
python binding table matches the header, host-only helpers agree with the oracle, and the product
fails loudly (never falls back) without a CUDA device.  No GPU compute is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fftop_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fop_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ["fop_create", "fop_destroy", "fop_carr_madan", "fop_cos", "fop_fsts",
                 "fop_objfn_cf", "fop_cos_loss", "fop_fft", "fop_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    import fftoptionpricing_b200 as F
    lib = ctypes.CDLL(F.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/fftop_b200.h but not exported"
    assert lib.fop_abi_version() == 1


def test_python_binding_table_matches_header():
    import fftoptionpricing_b200 as F
    assert F.declared_symbols() == header_symbols()


def test_model_param_counts_and_status_strings():
    import fftoptionpricing_b200 as F
    from oracle import oracle as O
    for base in (0, 1, 2):
        for tc in (0, 1, 2):
            assert F.num_params(base, tc) == O.num_params(base, tc)
    assert F.num_params(7, 0) == 0 and F.num_params(0, 9) == 0
    lib = F.load_library()
    assert lib.fop_status_string(0) == b"ok"
    assert b"unsupported" in lib.fop_status_string(4)


def test_host_grid_helpers_match_reference(oracles):
    """getCarrMadanStrikes / getFangOostStrike / getStrikeUnderlying (OptionPricing.h:84-138) are
    host functions of the library; compare with the oracle's (reference headers where built)."""
    import fftoptionpricing_b200 as F
    ora = oracles["reference"] or oracles["port"]
    lib = F.load_library()
    dp = ctypes.POINTER(ctypes.c_double)
    a = np.empty(1024)
    lib.fop_carr_madan_strikes(1024, .25, 50.0, a.ctypes.data_as(dp))
    np.testing.assert_allclose(a, ora.carr_madan_strikes(1024, .25, 50.0), rtol=2e-16, atol=0)
    lib.fop_fang_oost_strikes(-5.0, 5.0, 50.0, 1024, a.ctypes.data_as(dp))
    np.testing.assert_allclose(a, ora.fang_oost_strikes(-5.0, 5.0, 50.0, 1024), rtol=2e-16, atol=0)
    lib.fop_strike_underlying(-5.0, 5.0, 50.0, 1024, a.ctypes.data_as(dp))
    np.testing.assert_allclose(a, ora.strike_underlying(-5.0, 5.0, 50.0, 1024), rtol=2e-16, atol=0)


def test_no_cpu_fallback():
    """Without a CUDA device fop_create must fail (FOP_CUDA_ERROR) -- never a silent CPU path."""
    import torch

    import fftoptionpricing_b200 as F
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(F.FopError) as ei:
        F.Engine(0)
    assert ei.value.status == 2


def test_product_never_imports_oracle():
    """The product package must not import, link or dlopen anything under oracle/ (a product path
    routed through the oracle would void every parity claim)."""
    pkg = os.path.join(ROOT, "fftoptionpricing_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libfftop_oracle|\bora_[a-z_]+\s*\(|oracle_abi\.h",
                     re.M)
    for dp_, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp_, f), errors="ignore").read()
                assert not pat.search(txt), f"{f} references the oracle"
