"""GPU parity: the CUDA engine, driven through the reference-style Python API and the
C-ABI, against (a) the golden fields produced by the real reference and (b) the NumPy
oracle on the same seeded inputs.  Tolerance (north_star): rel-L2 <= 1e-10 for fp64
fields, <= 1e-4 for fp32 fields."""
import numpy as np
import pytest

from oracle import cases as C
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', [k['name'] for k in C.CASES])
def test_case_vs_reference_golden(product, name):
    case = C.CASES_BY_NAME[name]
    got = H.run_product(product, case)
    want = H.load_golden(case)
    errs = H.worst_rel_l2(got, want)
    tol = H.tolerance(case)
    assert max(errs.values()) <= tol, (name, errs)
    for n in C.FIELDS:
        assert np.all(np.isfinite(got[n]))
        assert np.asarray(got[n]).dtype == np.dtype(case['dtype'])


@pytest.mark.parametrize('name', ['shpf_f64_xpml', 'fdtd_f64_allpml', 'pstd_f64_allpml'] +
                         [k['name'] for k in C.LIVE_CASES])
def test_case_vs_oracle_live(product, name):
    case = dict(C.CASES_BY_NAME[name])
    if name in ('shpf_f64_xpml', 'fdtd_f64_allpml', 'pstd_f64_allpml'):
        case['steps'] = 17          # a step count no golden was generated for
    got = H.run_product(product, case)
    want = C.run_oracle(case)
    errs = H.worst_rel_l2(got, want)
    assert max(errs.values()) <= H.tolerance(case), (name, errs)


@pytest.mark.parametrize('name', ['shpf_f64_xpml_64', 'shpf_f64_allpml_64_r2', 'shpf_c64_xpml_64'])
def test_two_kernel_path_matches_fused(product, name, monkeypatch):
    """The fused persistent SHPF half-step and the two-kernel path give the same fields."""
    case = C.CASES_BY_NAME[name]
    monkeypatch.setenv('IES_B200_FUSED', '1')
    a = H.run_product(product, case)
    monkeypatch.setenv('IES_B200_FUSED', '0')
    b = H.run_product(product, case)
    for n in C.FIELDS:
        assert np.array_equal(np.asarray(a[n]), np.asarray(b[n])), n
