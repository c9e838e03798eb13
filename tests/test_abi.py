"""CPU: the C-ABI library loads, exports every symbol include/epa_b200.h declares, and refuses to
run without a CUDA device (there is no CPU fallback in the product path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers

ROOT = helpers.ROOT


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            text = open(os.path.join(inc, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(epa_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(os.path.join(ROOT, "epa-ng_b200", "libepa_b200.so"))
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    bound = {n for n, _, _ in built.capi.SYMBOLS} | {n for n, _, _ in built.session.HOST_SYMBOLS}
    assert declared == bound, f"ctypes bindings and headers disagree: {declared ^ bound}"


def test_record_layout_matches_reference_placement(built):
    # src/sample/Placement.hpp:49-53: {size_t branch_id; double likelihood, lwr, pendant, distal}
    assert built.capi.PLACEMENT_DTYPE.itemsize == 40
    assert built.capi.PLACEMENT_DTYPE.names == ("branch_id", "likelihood", "lwr", "pendant_length", "distal_length")


def test_default_options_match_reference(built):
    # src/util/Options.hpp:5-35
    o = built.capi.default_options()
    assert (o.prescoring, o.premasking, o.sliding_blo, o.filter_acc_lwr) == (1, 1, 1, 0)
    assert o.prescoring_threshold == 0.99999 and o.support_threshold == 0.01
    assert (o.filter_min, o.filter_max) == (1, 7)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path needs a GPU-less host")
    case = helpers.cfg1_case()
    with pytest.raises(built.capi.EpaError) as ei:
        helpers.make_context(case)
    assert ei.value.code == built.capi.EPA_ERR_CUDA
    assert "no CPU path" in str(ei.value)
