"""GPU: the canonical RT driver loop (TF/IF/SF spaces + Sx collectors + FieldAtPoint) on the
engine, against (a) the spectra the reference shipped from its own cupy/GPU runs
(graph/simple_2slab_{SHPF,FDTD}, complex64 -> fp32 tolerance 1e-4) and (b) the reference-API
oracle run on the CPU for a short complex128 run (1e-10)."""
import os

import numpy as np
import pytest

from tests import rt_tutorial as RT

pytestmark = pytest.mark.gpu
SHIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'shipped')


def rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b)))


@pytest.mark.parametrize('method,grid,steps', [('SHPF', (360, 16, 32), 15000), ('FDTD', (360, 20, 30), 3000)])
def test_shipped_reference_spectra(product, tmp_path, method, grid, steps):
    checks = tuple(t for t in (3000, 15000) if t <= steps)
    r = RT.run(product, 'b200', method, *grid, steps, dtype=np.complex64, cal_at=checks,
               savedir=str(tmp_path) + '/')
    assert np.allclose(r['freqs'], np.load(os.path.join(SHIP, f'{method}_freqs.npy')), rtol=1e-12)
    for t in checks:
        for name in ('TF_R', 'IF_R', 'SF_L'):
            want = np.load(os.path.join(SHIP, f'{method}_{name}_{t:07d}tstep_area.npy'))
            got = r['out'][(name, t)]['area']
            assert rel(got, want) <= 1e-4, (method, name, t, rel(got, want))
            # file naming / shapes of collector.py:351-360
            assert os.path.exists(os.path.join(str(tmp_path), 'Sx', f'{name}_DFT_Ey_{t:07d}tstep_rank00.npy'))
            assert os.path.exists(os.path.join(str(tmp_path), 'Sx', f'{name}_{t:07d}tstep_area.npy'))
        for name in ('TF_R', 'IF_R', 'SF_L'):
            want = np.load(os.path.join(SHIP, f'{method}_{name}_DFT_Ey_0003000tstep_col00.npy'))
            got = r['out'][(name, 3000)]['Ey'][:, 0, 0]
            assert rel(got, want) <= 1e-4
            assert r['out'][(name, 3000)]['Ey'].shape == (165,) + tuple(grid[1:])
    # energy conservation of the lossless slabs: R + T = 1 (plotter.py:336-398)
    t = checks[-1]
    R = np.abs(r['out'][('SF_L', t)]['area']) / np.abs(r['out'][('IF_R', t)]['area'])
    T = np.abs(r['out'][('TF_R', t)]['area']) / np.abs(r['out'][('IF_R', t)]['area'])
    if t >= 15000:
        assert np.all(np.abs(R + T - 1) < 3e-3)


def test_collectors_and_probe_vs_reference_api_oracle(product, tmp_path):
    """Short complex128 SHPF run: device collectors (incl. lazy SF = TF - IF) and the
    FieldAtPoint recorder against the same loop driven through the oracle on the CPU."""
    from oracle import ies_oracle as O
    steps, grid = 60, (48, 16, 16)
    r = RT.run(product, 'b200', 'SHPF', *grid, steps, dtype=np.complex128, cal_at=(steps,),
               savedir=str(tmp_path) + '/', peak_pos=20, probe=True)
    # oracle replay of the same loop
    Nx, Ny, Nz = grid
    Lx, Ly, Lz = 720e-6, 512e-6, 512e-6
    gap = (Lx / Nx, Ly / Ny, Lz / Nz)
    dt = r['dt']
    mk = lambda: O.OracleSpace(grid, gap, dt, steps, np.complex128, np.complex128, method='SHPF')
    TF, IF = mk(), mk()
    sets = []
    for sp in (TF, IF):
        sp.apply_PML({'x': '+-', 'y': '', 'z': ''}, 10)
        sp.apply_PBC({'x': False, 'y': True, 'z': True})
        sets.append(O.OracleSetter(sp, (Lx * 0.2, 0, 0), (Lx * 0.2, Ly, Lz), (0, 0, 0)))
    for a, b in ((160e-6, 260e-6), (460e-6, 560e-6)):
        TF.eps[round(a / gap[0]):round(b / gap[0])] = 4 * 8.8541878128e-12
    from scipy.constants import epsilon_0
    TF.eps[TF.eps > 2 * epsilon_0] = 4 * epsilon_0
    TF.init_update_constants(); IF.init_update_constants()
    fields = {'TF': lambda n: getattr(TF, n), 'IF': lambda n: getattr(IF, n),
              'SF': lambda n: getattr(TF, n) - getattr(IF, n)}
    cols = {'TF_R': O.OracleSx(fields['TF'], TF, Lx * 0.85, (0, 0), (Ly, Lz), r['freqs']),
            'IF_R': O.OracleSx(fields['IF'], IF, Lx * 0.85, (0, 0), (Ly, Lz), r['freqs']),
            'SF_L': O.OracleSx(fields['SF'], TF, Lx * 0.15, (0, 0), (Ly, Lz), r['freqs'])}
    px, py, pz = round(Lx * 0.5 / gap[0]), round(Ly * 0.5 / gap[1]), round(Lz * 0.5 / gap[2])
    sig = {n: np.zeros(steps, dtype=np.complex128) for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')}
    for t in range(steps + 1):
        p = O.gaussian_pulse_c(t, dt, 100e-6, 0.08, 20)
        for s in sets: s.put_src('Ey', p, 'soft')
        TF.update_h(t); TF.update_e(t); IF.update_h(t); IF.update_e(t)
        for cobj in cols.values(): cobj.do_RFT(t)
        if t < steps:
            for n in sig: sig[n][t] = getattr(TF, n)[px, py, pz]
    for name, cobj in cols.items():
        got = r['out'][(name, steps)]
        for comp in ('Ey', 'Ez', 'Hy', 'Hz'):
            den = max(np.linalg.norm(cobj.DFT['Ey']), 1e-300)
            assert np.linalg.norm(got[comp] - cobj.DFT[comp]) / den <= 1e-10, (name, comp)
        assert rel(got['area'], cobj.get_Sx()) <= 1e-9, name
    for n in ('Ey', 'Hz'):
        assert rel(getattr(r['probe'], n + '_t'), sig[n]) <= 1e-10, n
    # lazy scattered field read back on the host equals TF - IF
    assert rel(np.asarray(r['SF'].Ey), TF.Ey - IF.Ey) <= 1e-10


@pytest.mark.parametrize('ranks', [1, 2])
def test_sy_sz_collectors_vs_stepwise_numpy_dft(product, tmp_path, ranks):
    """Sy / Sz (collector.py:387-801) on the device (time-blocked accumulation, planes spanning the
    x-slabs) against the reference's formula applied step by step in NumPy to the fields read back
    after every step (collector.py:508-527, 716-735)."""
    from oracle import cases as C
    from ies_b200 import comm
    case = dict(C.CASES_BY_NAME['shpf_f64_allpml'])
    steps = 37                                   # two full blocks of 16 + a partial one
    (Lx, Ly, Lz), gap, dt = C.geometry(case)
    grp = comm.LocalGroup(ranks)
    spaces, setters, sys_, szs = [], [], [], []
    wv = np.linspace(60e-6, 90e-6, 7)
    freqs = 299792458.0 / wv
    path = str(tmp_path) + '/'
    for r in range(ranks):
        kw = dict(method='SHPF', engine='b200')
        if ranks > 1: kw['comm'] = grp.comm(r)
        sp = product.space.Basic3D(case['grid'], gap, dt, steps + 1, np.float64, np.complex128, **kw)
        sp.malloc()
        sp.apply_PML(case['pml'], case['npml'])
        s0, s1 = C.source_box(case)
        setters.append(product.source.Setter(sp, s0, s1, case['mmt']))
        for (b0, b1, er, mr) in C.box_list(case):
            product.structure.Box('box', sp, b0, b1, er, mr)
        sp.init_update_constants()
        spaces.append(sp)
        sys_.append(product.collector.Sy('sy', path, sp, 0.4 * Ly, (0.1 * Lx, 0.2 * Lz), (0.9 * Lx, 0.8 * Lz), freqs, 'b200'))
        szs.append(product.collector.Sz('sz', path, sp, 0.6 * Lz, (0.1 * Lx, 0.2 * Lz), (0.9 * Lx, 0.8 * Lz), freqs, 'b200'))
    c0 = sys_[0]
    X0, X1, Z0, Z1, Yc = c0.xsrt, c0.xend, c0.zsrt, c0.zend, c0.ysrt
    d0 = szs[0]
    YS, YE, Zc = d0.ysrt, d0.yend, d0.zsrt
    want = {n: 0 for n in ('yEx', 'yEz', 'yHx', 'yHz', 'zEx', 'zEy', 'zHx', 'zHy')}
    full = lambda n: np.concatenate([np.asarray(getattr(sp, n)[:, :, :]) for sp in spaces], axis=0)
    f = freqs[:, None, None]
    for t in range(steps):
        pv = C.pulse_value(case, t, dt)
        for s in setters: s.put_src(case['src_field'], pv, case['put'])
        for sp in spaces: sp.updateH(t)
        for sp in spaces: sp.updateE(t)
        for c in sys_ + szs: c.do_RFT(t)
        ph = np.exp(2.j * np.pi * f * t * dt) * dt
        F = {n: full(n) for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')}
        want['yEx'] = want['yEx'] + F['Ex'][X0:X1, Yc, Z0:Z1] * ph
        want['yEz'] = want['yEz'] + F['Ez'][X0:X1, Yc, Z0:Z1] * ph
        want['yHx'] = want['yHx'] + F['Hx'][X0:X1, Yc, Z0:Z1] * ph
        want['yHz'] = want['yHz'] + F['Hz'][X0:X1, Yc, Z0:Z1] * ph
        want['zEx'] = want['zEx'] + F['Ex'][X0:X1, YS:YE, Zc] * ph
        want['zEy'] = want['zEy'] + F['Ey'][X0:X1, YS:YE, Zc] * ph
        want['zHx'] = want['zHx'] + F['Hx'][X0:X1, YS:YE, Zc] * ph
        want['zHy'] = want['zHy'] + F['Hy'][X0:X1, YS:YE, Zc] * ph
    # one process plays all ranks: the other ranks save their parts before rank 0 concatenates
    for c in reversed(sys_): c.get_Sy(steps)
    for c in reversed(szs): c.get_Sz(steps)
    got_sy, got_sz = sys_[0], szs[0]                      # rank 0 holds the concatenated result
    Sy = 0.5 * (-(want['yEx'].real * want['yHz'].real) - (want['yEx'].imag * want['yHz'].imag)
                + (want['yEz'].real * want['yHx'].real) + (want['yEz'].imag * want['yHx'].imag))
    Sz = 0.5 * (-(want['zEy'].real * want['zHx'].real) - (want['zEy'].imag * want['zHx'].imag)
                + (want['zEx'].real * want['zHy'].real) + (want['zEx'].imag * want['zHy'].imag))
    assert got_sy.Sy.shape == Sy.shape and got_sz.Sz.shape == Sz.shape
    assert np.linalg.norm(Sy) > 0 and np.linalg.norm(Sz) > 0
    assert rel(got_sy.Sy, Sy) <= 1e-10, rel(got_sy.Sy, Sy)
    assert rel(got_sz.Sz, Sz) <= 1e-10, rel(got_sz.Sz, Sz)
    assert rel(got_sy.Sy_area, Sy.sum(axis=(1, 2)) * spaces[0].dx * spaces[0].dz) <= 1e-10
