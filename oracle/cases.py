"""Parity cases shared by the pin script, the oracle tests and the GPU parity tests
-- TEST INFRASTRUCTURE (see oracle/ies_oracle.py header).

A case is a plain dict.  `run_api(ns, case, engine)` drives any implementation
that exposes the reference's Python API (the reference itself via
oracle/ref_shims.py, or the product package ies_b200); `run_oracle(case)`
drives the NumPy restatement.  Both return {field name: ndarray} of the global
fields after `steps` steps (slabs concatenated along x).
"""
from __future__ import annotations

import numpy as np
from scipy.constants import c

um = 1e-6
FIELDS = ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')


def _case(name, method, dtype, grid, steps=30, pml=None, npml=4, bbc=None, pbc=None,
          mmt=(0., 0., 0.), ranks=1, src='plane', src_field='Ey', boxes=True, put='soft',
          pulse=None, mmtdtype=None, golden=None, sphere=False, probe=False, delta=False,
          nrank_quirk=False):
    dtype = np.dtype(dtype)
    cplx = dtype.kind == 'c'
    if mmtdtype is None:
        mmtdtype = np.complex128 if dtype in (np.dtype('float64'), np.dtype('complex128')) else np.complex64
    return dict(name=name, method=method, dtype=dtype.name, mmtdtype=np.dtype(mmtdtype).name,
                grid=tuple(grid), steps=steps,
                pml=pml if pml is not None else {'x': '+-', 'y': '', 'z': ''}, npml=npml,
                bbc=bbc, pbc=pbc, mmt=tuple(mmt), ranks=ranks, src=src, src_field=src_field,
                boxes=boxes, put=put, pulse=pulse or ('c' if cplx else 're'),
                golden=golden or name, sphere=sphere, probe=probe, delta=delta,
                nrank_quirk=nrank_quirk)


ALLPML = {'x': '+-', 'y': '+-', 'z': '+-'}
NOPML = {'x': '', 'y': '', 'z': ''}
PBC_YZ = {'x': False, 'y': True, 'z': True}
BBC_YZ = {'x': False, 'y': True, 'z': True}
BBC_ALL = {'x': True, 'y': True, 'z': True}
NO = {'x': False, 'y': False, 'z': False}
K1 = (0., 2 * np.pi / (512 * um) * 0.3, 2 * np.pi / (512 * um) * 0.2)
K3 = (2 * np.pi / (720 * um) * 0.25, 2 * np.pi / (512 * um) * 0.3, 2 * np.pi / (512 * um) * 0.2)

CASES = [
    _case('shpf_f64_xpml', 'SHPF', 'float64', (32, 16, 16), pbc=PBC_YZ, bbc=NO),
    _case('shpf_f32_xpml', 'SHPF', 'float32', (32, 16, 16), pbc=PBC_YZ, bbc=NO),
    _case('shpf_c64_xpml', 'SHPF', 'complex64', (32, 16, 32), pbc=PBC_YZ, bbc=NO),
    _case('shpf_c128_bloch_yz', 'SHPF', 'complex128', (32, 16, 16), bbc=BBC_YZ, pbc=NO, mmt=K1, src='point'),
    _case('shpf_f64_allpml', 'SHPF', 'float64', (32, 32, 16), pml=ALLPML, src='point'),
    _case('shpf_f64_allpml_r2', 'SHPF', 'float64', (32, 32, 16), pml=ALLPML, src='point', ranks=2, golden='shpf_f64_allpml'),
    _case('shpf_f64_xpml_r4', 'SHPF', 'float64', (32, 16, 16), pbc=PBC_YZ, bbc=NO, ranks=4, golden='shpf_f64_xpml'),
    _case('shpf_f64_ypml_only', 'SHPF', 'float64', (16, 32, 16), pml={'x': '', 'y': '+', 'z': '-'}, src='point'),
    _case('fdtd_f64_xpml_pbc', 'FDTD', 'float64', (32, 20, 18), pbc=PBC_YZ, bbc=NO),
    _case('fdtd_f32_xpml_pbc', 'FDTD', 'float32', (32, 20, 18), pbc=PBC_YZ, bbc=NO),
    _case('fdtd_c64_xpml_pbc', 'FDTD', 'complex64', (32, 20, 30), pbc=PBC_YZ, bbc=NO),
    _case('fdtd_c128_bloch_yz', 'FDTD', 'complex128', (32, 20, 18), bbc=BBC_YZ, pbc=NO, mmt=K1, src='point'),
    _case('fdtd_f64_allpml', 'FDTD', 'float64', (30, 28, 26), pml=ALLPML, bbc=NO, pbc=NO, src='point'),
    _case('fdtd_f64_allpml_r2', 'FDTD', 'float64', (30, 28, 26), pml=ALLPML, bbc=NO, pbc=NO, src='point', ranks=2),
    _case('fdtd_f64_xpml_pbc_r4', 'FDTD', 'float64', (32, 20, 18), pbc=PBC_YZ, bbc=NO, ranks=4, golden='fdtd_f64_xpml_pbc'),
    _case('fdtd_c128_pbcx', 'FDTD', 'complex128', (24, 20, 18), pml=NOPML, pbc={'x': True, 'y': True, 'z': True}, bbc=NO, src='point'),
    _case('pstd_c128_bloch_all', 'PSTD', 'complex128', (16, 16, 16), pml=NOPML, bbc=BBC_ALL, pbc=NO, mmt=K3, src='point', boxes=True),
    _case('pstd_f64_allpml', 'PSTD', 'float64', (32, 32, 16), pml=ALLPML, src='point'),
    _case('pstd_c64_xpml', 'PSTD', 'complex64', (32, 16, 16), src='plane'),
    _case('pstd_f64_nopml_hard', 'PSTD', 'float64', (16, 16, 32), pml=NOPML, src='point', put='hard'),
    # SHPF with a Bloch / periodic x axis: ghost-plane copies after each update (space.py:1898-1912, 2073-2085)
    _case('shpf_c128_bloch_x', 'SHPF', 'complex128', (16, 16, 16), pml=NOPML, bbc={'x': True, 'y': False, 'z': False},
          pbc=NO, mmt=K3, src='point'),
    # Bloch y/z on two slabs: partition-invariant result == the single-rank golden (the reference's
    # own 2-rank run skips the Bloch term on the slab-edge planes, see DESIGN "waived quirks")
    _case('shpf_c128_bloch_yz_r2', 'SHPF', 'complex128', (32, 16, 16), bbc=BBC_YZ, pbc=NO, mmt=K1, src='point',
          ranks=2, golden='shpf_c128_bloch_yz', nrank_quirk=True),
]
# Cases at the benchmarked kernel instantiations (256-point FFT lines: the only length with the
# shared forward/inverse twiddle table) and at or near the shapes BASELINE.json names.  The real
# reference runs them once in the build container (oracle/pin_against_reference.py --digest);
# the fixture is a DIGEST of its fields (strided sub-sample + norms, tests/golden/digest/), the
# GPU tests compare the full fields with the oracle live and the sub-sample with the digest.
DIGEST_CASES = [
    _case('shpf_f64_xpml_256', 'SHPF', 'float64', (12, 256, 256), steps=6, npml=4, pbc=PBC_YZ, bbc=NO),
    _case('shpf_f64_allpml_256_r2', 'SHPF', 'float64', (12, 256, 256), steps=6, npml=4, pml=ALLPML, src='point', ranks=2),
    _case('shpf_f32_xpml_256', 'SHPF', 'float32', (12, 256, 256), steps=6, npml=4, pbc=PBC_YZ, bbc=NO),
    _case('shpf_c64_xpml_256', 'SHPF', 'complex64', (12, 256, 256), steps=4, npml=3, pbc=PBC_YZ, bbc=NO),
    _case('shpf_c128_bloch_256', 'SHPF', 'complex128', (12, 256, 256), steps=4, npml=3, bbc=BBC_YZ, pbc=NO, mmt=K1, src='point'),
    _case('pstd_f64_256', 'PSTD', 'float64', (16, 256, 256), steps=4, npml=4, src='plane'),
    # config 2: FDTD 256x64x64, CPML x, PBC y/z, fp64
    _case('cfg2_fdtd_256x64x64', 'FDTD', 'float64', (256, 64, 64), steps=24, npml=10, pbc=PBC_YZ, bbc=NO),
    # config 3 slab: the headline grid's yz extent with the headline kernels, x-CPML + plane source
    _case('cfg3_shpf_32x256x256', 'SHPF', 'float64', (32, 256, 256), steps=8, npml=10, pbc=PBC_YZ, bbc=NO),
    # config 4: PSTD 128^3 complex128, Bloch on all axes, Delta dipole + FieldAtPoint probe
    _case('cfg4_pstd_128_bloch', 'PSTD', 'complex128', (128, 128, 128), steps=3, pml=NOPML, bbc=BBC_ALL, pbc=NO,
          mmt=K3, src='point', boxes=True, delta=True, probe=True),
    # config 5 slab: 512-point lines, CPML on all six faces, eps_r = 4 sphere, real fp64
    _case('cfg5_shpf_24x512x512_sphere', 'SHPF', 'float64', (24, 512, 512), steps=4, npml=10, pml=ALLPML,
          src='plane', boxes=False, sphere=True),
]
# Larger live-only cases (no golden file; compared with the oracle at test time).
LIVE_CASES = [
    _case('shpf_f64_xpml_64', 'SHPF', 'float64', (40, 64, 64), steps=10, npml=6, pbc=PBC_YZ, bbc=NO),
    _case('shpf_f64_allpml_64_r2', 'SHPF', 'float64', (40, 64, 64), steps=10, npml=6, pml=ALLPML, src='point', ranks=2),
    _case('shpf_f32_allpml_64', 'SHPF', 'float32', (40, 64, 64), steps=10, npml=6, pml=ALLPML, src='point'),
    _case('shpf_c64_xpml_64', 'SHPF', 'complex64', (40, 64, 64), steps=10, npml=6, pbc=PBC_YZ, bbc=NO),
    _case('shpf_c128_bloch_yz_128', 'SHPF', 'complex128', (28, 128, 128), steps=6, bbc=BBC_YZ, pbc=NO, mmt=K1, src='point'),
    _case('shpf_f64_xpml_16x64', 'SHPF', 'float64', (24, 16, 64), steps=10, pbc=PBC_YZ, bbc=NO),
    _case('pstd_f64_xpml_64', 'PSTD', 'float64', (64, 32, 64), steps=8, npml=6, src='plane'),
    # long lines: three-stage FFT plans (radix 16/16/2 at 512), every power of two in between
    _case('shpf_f64_z512_y128', 'SHPF', 'float64', (12, 128, 512), steps=6, npml=4, pbc=PBC_YZ, bbc=NO, src='point'),
    _case('shpf_f64_y512_z32', 'SHPF', 'float64', (12, 512, 32), steps=6, npml=4, pml=ALLPML, src='point'),
    _case('shpf_c64_z512_y16', 'SHPF', 'complex64', (12, 16, 512), steps=6, npml=4, pbc=PBC_YZ, bbc=NO, src='point'),
    # the split real / imaginary exchange of the 512-point z lines for the other element types
    _case('shpf_f32_z512_y32', 'SHPF', 'float32', (12, 32, 512), steps=6, npml=4, pml=ALLPML, src='point'),
    _case('shpf_c128_z512_y16', 'SHPF', 'complex128', (12, 16, 512), steps=4, npml=3, bbc=BBC_YZ, pbc=NO, mmt=K1, src='point'),
    # spectral axes that are not a power of two (the reference takes any N; the engine applies the derivative
    # as a circulant sum there): the shipped 50^3 set-up's line length, mixed lengths, complex odd lengths
    _case('shpf_f64_50cube', 'SHPF', 'float64', (30, 50, 50), steps=8, npml=4, pbc=PBC_YZ, bbc=NO),
    _case('shpf_f64_allpml_20x36', 'SHPF', 'float64', (24, 20, 36), steps=10, npml=4, pml=ALLPML, src='point'),
    _case('shpf_f32_24x40_r2', 'SHPF', 'float32', (24, 24, 40), steps=10, npml=4, pbc=PBC_YZ, bbc=NO, ranks=2),
    _case('shpf_c128_bloch_odd', 'SHPF', 'complex128', (20, 21, 27), steps=8, bbc=BBC_YZ, pbc=NO, mmt=K1, src='point'),
    _case('pstd_f64_20x24x36', 'PSTD', 'float64', (20, 24, 36), steps=8, npml=4, src='plane'),
    _case('pstd_c128_bloch_odd', 'PSTD', 'complex128', (15, 18, 21), steps=6, pml=NOPML, bbc=BBC_ALL, pbc=NO, mmt=K3, src='point'),
    # sources on every kind of component (E_z, E_x, H_y, H_x; soft / hard; point / plane)
    _case('shpf_f64_src_ez', 'SHPF', 'float64', (24, 32, 32), steps=10, pbc=PBC_YZ, bbc=NO, src='point', src_field='Ez'),
    _case('shpf_f64_src_ex_hard', 'SHPF', 'float64', (24, 32, 32), steps=10, pml=ALLPML, src='point', src_field='Ex', put='hard'),
    _case('shpf_f64_src_hy_r2', 'SHPF', 'float64', (24, 32, 32), steps=10, pbc=PBC_YZ, bbc=NO, src='point', src_field='Hy', ranks=2),
    _case('shpf_c128_src_hx_bloch', 'SHPF', 'complex128', (24, 32, 16), steps=10, bbc=BBC_YZ, pbc=NO, mmt=K1, src='point', src_field='Hx'),
    _case('shpf_f32_src_ez_plane', 'SHPF', 'float32', (24, 16, 64), steps=10, pbc=PBC_YZ, bbc=NO, src='plane', src_field='Ez'),
]
CASES_BY_NAME = {k['name']: k for k in CASES + LIVE_CASES + DIGEST_CASES}


def geometry(case):
    Nx, Ny, Nz = case['grid']
    Lx, Ly, Lz = 720 * um, 512 * um, 512 * um
    dx, dy, dz = Lx / Nx, Ly / Ny, Lz / Nz
    dt = 0.25 * min(dx, dy, dz) / c
    return (Lx, Ly, Lz), (dx, dy, dz), dt


def source_box(case):
    (Lx, Ly, Lz), (dx, dy, dz), dt = geometry(case)
    if case['src'] == 'plane':
        xs = Lx * 0.3
        return (xs, 0, 0), (xs, Ly, Lz)
    # point dipole (src_end = src_srt + one cell, source.py docstring case 2)
    xs, ys, zs = Lx * 0.4, Ly * 0.45, Lz * 0.55
    return (xs, ys, zs), (xs + dx, ys + dy, zs + dz)


def pulse_value(case, step, dt):
    """Gaussian pulse (source.py:278-290) with a short rise so 30 steps see it."""
    if case.get('delta'):
        return 1. if step == 1 else 0.         # source.Delta(pick=1).apply (source.py:479-488)
    wvc, spread, peak = 100 * um, 0.3, 12
    w0 = 2 * np.pi * (c / wvc)
    ws = spread * w0
    tc = peak * dt
    env = np.exp((-.5) * (((step * dt - tc) * ws) ** 2))
    if case['pulse'] == 'c':
        return env * np.exp(-1j * w0 * (step * dt - tc))
    return env * np.cos(w0 * (step * dt - tc))


def box_list(case):
    (Lx, Ly, Lz), d, dt = geometry(case)
    if not case['boxes']:
        return []
    return [((Lx * 0.5, 0, 0), (Lx * 0.65, Ly, Lz), 4., 1.),
            ((Lx * 0.7, Ly * 0.25, Lz * 0.25), (Lx * 0.8, Ly * 0.75, Lz * 0.6), 2.25, 1.5)]


def sphere_spec(case):
    """(center index, radius, eps_r, mu_r) of the case's dielectric sphere or None
    (examples/mie/mie_scattering.py:241-250 style: centre in grid indices, radius in length)."""
    if not case.get('sphere'):
        return None
    Nx, Ny, Nz = case['grid']
    (Lx, Ly, Lz), (dx, dy, dz), dt = geometry(case)
    return (Nx // 2, Ny // 2, Nz // 2), min(5.2 * dx, 0.3 * Ly), 4., 1.


def probe_loc(case):
    (Lx, Ly, Lz), d, dt = geometry(case)
    return (Lx * 0.6, Ly * 0.3, Lz * 0.7)


# ------------------------------------------------------------------ API runner
def build_api(ns, case, engine):
    """Build (space, setter) with the reference's API (tutorials/RT_simple_slabs.py:55-249)."""
    (Lx, Ly, Lz), gap, dt = geometry(case)
    fd = np.dtype(case['dtype']).type
    md = np.dtype(case['mmtdtype']).type
    sp = ns.space.Basic3D(case['grid'], gap, dt, case['steps'] + 1, fd, md,
                          method=case['method'], engine=engine)
    sp.malloc()
    sp.apply_PML(case['pml'], case['npml'])
    if case['bbc'] is not None:
        sp.apply_BBC(case['bbc'])
    if case['pbc'] is not None:
        sp.apply_PBC(case['pbc'])
    s0, s1 = source_box(case)
    setter = ns.source.Setter(sp, s0, s1, case['mmt'])
    for (b0, b1, er, mr) in box_list(case):
        ns.structure.Box('box', sp, b0, b1, er, mr)
    sph = sphere_spec(case)
    if sph is not None:
        ns.structure.Sphere('sphere', sp, *sph)
    sp.init_update_constants()
    return sp, setter


def step_api(sp, setter, case, t):
    setter.put_src(case['src_field'], pulse_value(case, t, sp.dt), case['put'])
    sp.updateH(t)
    sp.updateE(t)


def run_api(ns, case, engine, to_numpy=np.asarray):
    """Single-rank run through the reference-style API."""
    assert case['ranks'] == 1
    sp, setter = build_api(ns, case, engine)
    probe = None
    if case.get('probe'):
        probe = ns.collector.FieldAtPoint('probe', '/tmp', sp, probe_loc(case), engine)
    for t in range(case['steps']):
        step_api(sp, setter, case, t)
        if probe is not None:
            probe.get_time_signal(t)
    out = {n: to_numpy(getattr(sp, n)[:, :, :]) for n in FIELDS}
    if probe is not None:
        out['probe'] = np.stack([to_numpy(getattr(probe, n + '_t'))[:case['steps']] for n in FIELDS])
    return out


# --------------------------------------------------------------- oracle runner
def build_oracle(case):
    from . import ies_oracle as O
    (Lx, Ly, Lz), gap, dt = geometry(case)
    cl = O.OracleCluster(case['ranks'], case['grid'], gap, dt, case['steps'] + 1,
                         np.dtype(case['dtype']).type, np.dtype(case['mmtdtype']).type,
                         method=case['method'])
    s0, s1 = source_box(case)
    setters = []
    for sp in cl.slabs:
        sp.apply_PML(case['pml'], case['npml'])
        if case['bbc'] is not None:
            sp.apply_BBC(case['bbc'])
        if case['pbc'] is not None:
            sp.apply_PBC(case['pbc'])
        setters.append(O.OracleSetter(sp, s0, s1, case['mmt']))
        for (b0, b1, er, mr) in box_list(case):
            oracle_box(sp, b0, b1, er, mr)
        sph = sphere_spec(case)
        if sph is not None:
            oracle_sphere(sp, *sph)
        sp.init_update_constants()
    return cl, setters


def oracle_box(sp, srt, end, eps_r, mu_r):
    """structure.Box (structure.py:108-192) + Structure._get_local_x_loc (17-107)."""
    from . import ies_oracle as O
    from scipy.constants import epsilon_0, mu_0
    xs, ys, zs = (round(srt[0] / sp.dx), round(srt[1] / sp.dy), round(srt[2] / sp.dz))
    xe, ye, ze = (round(end[0] / sp.dx), round(end[1] / sp.dy), round(end[2] / sp.dz))
    assert xs < xe and ys < ye and zs < ze
    g, l = O.local_x_loc(sp, xs, xe)
    if g is not None:
        sp.eps[l[0]:l[1], ys:ye, zs:ze] = eps_r * epsilon_0
        sp.mu[l[0]:l[1], ys:ye, zs:ze] = mu_r * mu_0


def oracle_sphere(sp, center, radius, eps_r, mu_r):
    """structure.Sphere (structure.py:395-480): per x plane of the sphere the disc
    (j-cy)^2 dy^2 + (k-cz)^2 dz^2 <= rr^2, rr = radius*sin(arccos(|i-cx| dx / radius))."""
    from . import ies_oracle as O
    from scipy.constants import epsilon_0, mu_0
    nr = round(radius / sp.dx)
    gs, ge = center[0] - nr, center[0] + nr
    assert gs >= 0 and ge < sp.Nx
    g, l = O.local_x_loc(sp, gs, ge)
    if g is None:
        return
    portion = np.arange(g[0] - center[0] + nr, g[1] - center[0] + nr)
    rx = abs(portion - nr)
    rr = radius * np.sin(np.arccos(rx * sp.dx / radius))
    j = np.arange(sp.Ny); k = np.arange(sp.Nz)
    d2 = (((j - center[1]) * sp.dy) ** 2)[:, None] + (((k - center[2]) * sp.dz) ** 2)[None, :]
    mask = d2[None] <= (rr ** 2)[:, None, None]
    sp.eps[l[0]:l[1]][mask] = eps_r * epsilon_0
    sp.mu[l[0]:l[1]][mask] = mu_r * mu_0


def run_oracle(case):
    cl, setters = build_oracle(case)
    dt = cl.slabs[0].dt
    sig = []
    if case.get('probe'):
        s0 = cl.slabs[0]
        loc = probe_loc(case)
        pidx = (round(loc[0] / s0.dx), round(loc[1] / s0.dy), round(loc[2] / s0.dz))   # collector.py:146-150
    for t in range(case['steps']):
        p = pulse_value(case, t, dt)
        for s in setters:
            s.put_src(case['src_field'], p, case['put'])
        cl.update_h(t)
        cl.update_e(t)
        if case.get('probe'):
            sig.append([cl.gather(n)[pidx] for n in FIELDS])      # collector.py:169-201
    out = {n: cl.gather(n) for n in FIELDS}
    if case.get('probe'):
        out['probe'] = np.array(sig).T.astype(np.dtype(case['dtype']))
    return out


# ------------------------------------------------------------------ digests
DIGEST_STRIDE = (1, 8, 8)


def digest(fields, case):
    """Small fixture of a large run: every field sub-sampled with DIGEST_STRIDE, plus each
    field's L2 norm and plain sum (tests/golden/digest/<case>.npz)."""
    out = {}
    sx, sy, sz = DIGEST_STRIDE
    for n in FIELDS:
        a = np.asarray(fields[n])
        out[n] = np.ascontiguousarray(a[::sx, ::sy, ::sz])
        out[n + '_norm'] = np.float64(np.linalg.norm(a.ravel()))
        out[n + '_sum'] = np.complex128(a.sum(dtype=np.complex128 if a.dtype.kind == 'c' else np.float64))
    if 'probe' in fields:
        out['probe'] = np.asarray(fields['probe'])
    return out


def digest_errors(fields, dg):
    """rel-L2 of the sub-sample, and relative norm difference, per field group (see
    tests/helpers.worst_rel_l2 for the normalisation)."""
    sx, sy, sz = DIGEST_STRIDE
    errs = {}
    for grp in (('Ex', 'Ey', 'Ez'), ('Hx', 'Hy', 'Hz')):
        den = max(np.linalg.norm(np.asarray(dg[n]).ravel()) for n in grp)
        nden = max(float(dg[n + '_norm']) for n in grp)
        for n in grp:
            a = np.asarray(fields[n])
            num = np.linalg.norm((a[::sx, ::sy, ::sz] - dg[n]).ravel())
            errs[n] = float(num / den) if den > 0 else float(num)
            dn = abs(np.linalg.norm(a.ravel()) - float(dg[n + '_norm']))
            errs[n + '_norm'] = float(dn / nden) if nden > 0 else float(dn)
    if 'probe' in dg:
        den = np.linalg.norm(np.asarray(dg['probe']).ravel())
        num = np.linalg.norm((np.asarray(fields['probe']) - dg['probe']).ravel())
        errs['probe'] = float(num / den) if den > 0 else float(num)
    return errs


def group_rel_l2(got, want):
    """Per-field ||got - want|| normalised by the largest field norm of the same kind (E or H), so
    that components that are identically ~0 (e.g. Ex of an x-propagating plane wave) neither divide
    by zero nor turn round-off noise into a large relative error."""
    out = {}
    for grp in (('Ex', 'Ey', 'Ez'), ('Hx', 'Hy', 'Hz')):
        den = max(np.linalg.norm(np.asarray(want[n]).ravel()) for n in grp)
        for n in grp:
            num = np.linalg.norm((np.asarray(got[n]) - np.asarray(want[n])).ravel())
            out[n] = float(num / den) if den > 0 else float(num)
    return out


def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    return float(num / den) if den > 0 else float(num)
