"""The C-ABI library loads on a CPU-only box and exports every symbol
include/klu.h declares; compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "klu.h")).read()
    return sorted(set(re.findall(r"\b(klu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(klu):
    lib = klu.binding.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(klu.binding.SYMBOLS)


def test_version_and_defaults(klu):
    lib = klu.binding.lib()
    assert lib.klu_version() == 3
    o = klu.binding.KluOpts()
    lib.klu_opts_default(C.byref(o))
    assert o.acoustic_scale == 1.0 and o.graph_scale == 1.0 and o.insertion_penalty == 0.0
    assert o.beam == float("inf") and abs(o.beam_ratio - 0.9) < 1e-7 and o.nbest == 100
    assert o.max_arcs == 2**31 - 1 and o.max_states == 2**31 - 1


def test_no_cpu_fallback(klu):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(klu.KluError):
        klu.Engine(0)
