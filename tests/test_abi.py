"""CPU: the C-ABI shared library loads and exports every symbol include/ies_b200.h
declares, the ctypes table covers them all, and compute entry points fail loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'ies_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(ies_[a-z_0-9]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    from ies_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ies_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ies_b200 import _lib, space
    with pytest.raises(_lib.EngineError):
        space.Basic3D((32, 16, 16), (1e-6, 1e-6, 1e-6), 1e-16, 10, np.float64, np.complex128, method='SHPF')


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'ies_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in re.sub(r'#.*', '', src).replace('"""', ''), fn
