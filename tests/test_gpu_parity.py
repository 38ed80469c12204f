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


@pytest.mark.parametrize('name', ['shpf_f64_xpml', 'fdtd_f64_allpml', 'pstd_f64_allpml'])
def test_case_vs_oracle_live(product, name):
    case = dict(C.CASES_BY_NAME[name])
    case['steps'] = 17          # a step count no golden was generated for
    got = H.run_product(product, case)
    want = C.run_oracle(case)
    errs = H.worst_rel_l2(got, want)
    assert max(errs.values()) <= H.tolerance(case), (name, errs)
