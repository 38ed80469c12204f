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


ALT_CASES = ['shpf_f64_xpml', 'shpf_f32_xpml', 'shpf_c64_xpml', 'shpf_c128_bloch_yz', 'shpf_f64_allpml_r2',
             'shpf_f64_ypml_only', 'shpf_f64_xpml_64', 'shpf_f32_allpml_64', 'shpf_c128_bloch_yz_128',
             'shpf_f64_xpml_16x64', 'shpf_f64_src_ez', 'shpf_f64_src_ex_hard', 'shpf_f64_src_hy_r2',
             'shpf_c128_src_hx_bloch', 'shpf_f32_src_ez_plane']


@pytest.mark.parametrize('name', ALT_CASES)
def test_alternating_path_matches_two_kernel_path(product, name, monkeypatch):
    """The one-kernel-per-half-step SHPF path (shpf_half.cuh, default) and the z-line +
    y-line kernel pair compute bit-identical fields, including when a source writes the
    components whose derivative was precomputed (scratch refresh)."""
    case = C.CASES_BY_NAME[name]
    monkeypatch.setenv('IES_B200_ALT', '1')
    a = H.run_product(product, case)
    monkeypatch.setenv('IES_B200_ALT', '0')
    b = H.run_product(product, case)
    for n in C.FIELDS:
        assert np.array_equal(np.asarray(a[n]), np.asarray(b[n])), n


def test_alternating_path_field_write_between_half_steps(product, monkeypatch):
    """Writing E_z / H_x through the field API between updates invalidates the precomputed
    derivative planes; the refreshed run equals the two-kernel path bit for bit."""
    case = dict(C.CASES_BY_NAME['shpf_f64_xpml_64'])
    rng = np.random.default_rng(5)
    patch_e = rng.uniform(-1, 1, (3, 64, 64))
    patch_h = rng.uniform(-1, 1, (2, 64, 64)) * 1e-3
    out = {}
    for alt in ('1', '0'):
        monkeypatch.setenv('IES_B200_ALT', alt)
        sp, setter = C.build_api(product, case, 'b200')
        for t in range(4):
            C.step_api(sp, setter, case, t)
        sp.Ez[7:10, :, :] = patch_e            # stale d/dy E_z on planes 7..9
        sp.Ex[20, 3:5, :] = 0.25
        setter.put_src('Ey', 0.5, 'soft')
        sp.updateH(4)
        sp.Hx[30:32, :, :] = patch_h           # stale d/dz H_x on planes 30..31
        sp.updateE(4)
        sp.updateH(5)                          # two H updates in a row: whole scratch is stale
        sp.updateH(6)
        sp.updateE(6)
        out[alt] = {n: np.asarray(getattr(sp, n)[:, :, :]) for n in C.FIELDS}
    for n in C.FIELDS:
        assert np.array_equal(out['1'][n], out['0'][n]), n
