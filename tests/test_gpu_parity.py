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


@pytest.mark.parametrize('name', [k['name'] for k in C.DIGEST_CASES])
def test_large_case_vs_oracle_and_reference_digest(product, name):
    """The benchmarked kernel instantiations (256-point lines: k_zline<.,256,16> /
    k_yline_update<.,256,.> with the shared twiddle table; 512-point three-stage plans) and the
    BASELINE-config shapes: full fields against the oracle live, sub-sample / norms / probe signal
    against the digest of the real reference's run."""
    case = C.CASES_BY_NAME[name]
    got = H.run_product(product, case)
    tol = H.tolerance(case)
    derr = C.digest_errors(got, H.load_digest(case))
    assert max(derr.values()) <= tol, (name, 'digest', derr)
    want = C.run_oracle(case)
    errs = H.worst_rel_l2(got, want)
    assert max(errs.values()) <= tol, (name, 'oracle', errs)
    if 'probe' in want:
        assert C.rel_l2(got['probe'], want['probe']) <= tol


FUSED_CASES = ['shpf_f64_xpml_64', 'shpf_f64_allpml_64_r2', 'shpf_f32_allpml_64', 'shpf_f64_xpml_256',
               'shpf_f64_allpml_256_r2', 'shpf_f32_xpml_256', 'cfg3_shpf_32x256x256', 'cfg5_shpf_24x512x512_sphere']


@pytest.mark.parametrize('name', FUSED_CASES)
@pytest.mark.parametrize('ring', [0, 6])
def test_fused_half_step_matches_two_kernel_path(product, name, ring, monkeypatch):
    """The single-launch SHPF half-step (shpf_fused.cuh: z-line and y-line tiles as two roles of one
    grid, scratch ring in L2) computes bit-identical fields to k_zline + k_yline_update: every
    component sees the same expression and operands.  ring = 0: full-size scratch; 6: a ring of six
    planes with a lead of two, so slots are reused many times within a launch."""
    case = C.CASES_BY_NAME[name]
    monkeypatch.setenv('IES_B200_FUSED', '1')
    monkeypatch.setenv('IES_B200_FUSED_RING', str(ring))
    monkeypatch.setenv('IES_B200_FUSED_LEAD', '2')
    a = H.run_product(product, case)
    monkeypatch.setenv('IES_B200_FUSED', '0')
    b = H.run_product(product, case)
    for n in C.FIELDS:
        assert np.array_equal(np.asarray(a[n]), np.asarray(b[n])), n


PML_SPLIT_CASES = ['shpf_f64_allpml', 'shpf_f64_allpml_r2', 'shpf_f64_ypml_only', 'shpf_f64_xpml', 'fdtd_f64_allpml',
                   'fdtd_f64_allpml_r2', 'fdtd_f64_xpml_pbc', 'pstd_f64_allpml', 'pstd_c64_xpml', 'shpf_f64_allpml_64_r2',
                   'shpf_f64_allpml_256_r2', 'shpf_f64_y512_z32', 'cfg5_shpf_24x512x512_sphere', 'shpf_f32_allpml_64']


@pytest.mark.parametrize('name', PML_SPLIT_CASES)
def test_separate_cpml_pass_matches_in_kernel_cpml(product, name, monkeypatch):
    """CPML corrections applied by the separate pass over the absorber cells (k_pml_terms, default
    when a y or z face carries terms) and inside the update kernels: same statements in the same
    order, bit-identical in double precision.  Single-precision fields round G once more between
    the main update and the corrections in the separate pass -- as the reference does
    (space.py:801-803 then 1153-1162) -- so they agree to fp32 round-off."""
    case = C.CASES_BY_NAME[name]
    a = _run_with_env(product, case, monkeypatch, IES_B200_PML_SPLIT=1)
    b = _run_with_env(product, case, monkeypatch, IES_B200_PML_SPLIT=0)
    single = np.dtype(case['dtype']) in (np.dtype('float32'), np.dtype('complex64'))
    for n in C.FIELDS:
        if single:
            assert max(C.group_rel_l2(a, b).values()) <= 1e-5
        else:
            assert np.array_equal(np.asarray(a[n]), np.asarray(b[n])), n


def test_shpf_x_bloch_with_yz_bloch_is_refused(product):
    """The reference evaluates the y/z Bloch terms after the x ghost copies (space.py:1898-1930); the
    fused multiplier cannot, so the combination raises instead of returning different fields."""
    case = dict(C.CASES_BY_NAME['shpf_c128_bloch_x'])
    case['bbc'] = {'x': True, 'y': True, 'z': False}
    with pytest.raises(NotImplementedError):
        H.run_product(product, case)


def _run_with_env(product, case, monkeypatch, **env):
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))
    return H.run_product(product, case)


@pytest.mark.parametrize('name', ['fdtd_f64_xpml_pbc', 'fdtd_c64_xpml_pbc', 'fdtd_c128_bloch_yz', 'fdtd_f64_allpml',
                                  'fdtd_f64_allpml_r2', 'fdtd_c128_pbcx'])
def test_fdtd_vector_kernel_matches_scalar_kernel(product, name, monkeypatch):
    """k_fdtd_vec (16-byte accesses, tile-level fast path) and the one-thread-per-cell k_fdtd
    compute bit-identical fields."""
    case = C.CASES_BY_NAME[name]
    a = _run_with_env(product, case, monkeypatch, IES_B200_FDTD_VEC=1)
    b = _run_with_env(product, case, monkeypatch, IES_B200_FDTD_VEC=0)
    for n in C.FIELDS:
        assert np.array_equal(np.asarray(a[n]), np.asarray(b[n])), n


@pytest.mark.parametrize('name', ['shpf_f64_xpml', 'shpf_f64_allpml_r2', 'shpf_c128_bloch_yz', 'shpf_f64_xpml_64',
                                  'pstd_f64_allpml', 'pstd_c128_bloch_all'])
def test_uniform_tile_coefficients_match_array(product, name, monkeypatch):
    """Tiles whose coefficient is uniform skip the coefficient array (k_tile_uniform): same bits."""
    case = C.CASES_BY_NAME[name]
    a = _run_with_env(product, case, monkeypatch, IES_B200_CTILE=1)
    b = _run_with_env(product, case, monkeypatch, IES_B200_CTILE=0)
    for n in C.FIELDS:
        assert np.array_equal(np.asarray(a[n]), np.asarray(b[n])), n


@pytest.mark.parametrize('name', ['shpf_f64_xpml_64', 'fdtd_f64_xpml_pbc', 'pstd_f64_allpml'])
def test_two_phase_half_step_equals_whole_half_step(product, name):
    """ies_update_phase(half, 0) + ies_update_phase(half, 1) == ies_update_h / ies_update_e."""
    from ies_b200 import _lib
    lib = _lib.load()
    case = C.CASES_BY_NAME[name]
    out = []
    for phased in (False, True):
        sp, setter = C.build_api(product, case, 'b200')
        for t in range(case['steps']):
            setter.put_src(case['src_field'], C.pulse_value(case, t, sp.dt), case['put'])
            if phased:
                if sp._dirty: sp._finalize()
                for half in (_lib.HALF_H, _lib.HALF_E):
                    _lib.check(lib.ies_update_phase(sp._ctx, half, 0))
                    _lib.check(lib.ies_update_phase(sp._ctx, half, 1))
            else:
                sp.updateH(t); sp.updateE(t)
        out.append({n: np.asarray(getattr(sp, n)[:, :, :]) for n in C.FIELDS})
    for n in C.FIELDS:
        assert np.array_equal(out[0][n], out[1][n]), n
