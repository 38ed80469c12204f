#!/usr/bin/env python
"""Throughput of the other methods / dtypes on synthetic grids (development tool, GPU box only):
FDTD, PSTD and SHPF for f32/f64/c64/c128, CPML in x, periodic y/z, random fields.

    python tools/bench_methods.py [--steps 20] [--warmup 3]
"""
import argparse
import ctypes as C
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (ClockSampler, measured_peak)
um = 1e-6
C0 = 299792458.0

CASES = [
    ('FDTD', np.float64, (1024, 256, 256)),
    ('FDTD', np.float32, (1024, 256, 256)),
    ('FDTD', np.complex128, (512, 256, 256)),
    ('FDTD', np.float64, (256, 64, 64)),
    ('SHPF', np.float64, (1024, 256, 256)),
    ('SHPF', np.float32, (1024, 256, 256)),
    ('SHPF', np.complex64, (1024, 256, 256)),
    ('SHPF', np.complex128, (512, 256, 256)),
    ('SHPF', np.float64, (512, 512, 512)),
    ('SHPF', np.float64, (2048, 128, 128)),
    ('SHPF-allpml', np.float64, (1024, 256, 256)),
    ('SHPF-allpml', np.float64, (256, 512, 512)),
    ('FDTD-allpml', np.float64, (1024, 256, 256)),
    ('PSTD', np.complex128, (128, 128, 128)),
    ('PSTD', np.float64, (512, 256, 256)),
]
BYTES = {np.float64: 160, np.float32: 88, np.complex64: 160, np.complex128: 304}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--only', type=int, default=-1, help='index into CASES')
    ap.add_argument('--grep', default='', help='only cases whose label/grid string contains this')
    args = ap.parse_args()
    import ies_b200
    from ies_b200 import _lib
    lib = _lib.load()
    peak, _ = bench.measured_peak()
    rows = []
    for method, dt_, grid in (CASES if args.only < 0 else CASES[args.only:args.only + 1]):
        if args.grep and args.grep not in f'{method} {np.dtype(dt_).name} {grid}': continue
        allpml = method.endswith('-allpml')
        label = method
        method = method.split('-')[0]
        nx, ny, nz = grid
        gap = (720 * um / nx, 512 * um / ny, 512 * um / nz)
        dt = 0.25 * min(gap) / C0
        cplx = np.dtype(dt_).kind == 'c'
        sp = ies_b200.space.Basic3D(grid, gap, dt, 1000, dt_, np.complex128 if np.dtype(dt_).itemsize in (8, 16) else np.complex64,
                                    method=method, engine='b200')
        sp.malloc()
        pml = {'x': '+-', 'y': '+-' if allpml else '', 'z': '+-' if allpml else ''}
        sp.apply_PML(pml, 10)
        if allpml:
            pass
        elif cplx:
            sp.apply_BBC({'x': False, 'y': True, 'z': True}); sp.apply_PBC({'x': False, 'y': False, 'z': False})
        else:
            sp.apply_BBC({'x': False, 'y': False, 'z': False}); sp.apply_PBC({'x': False, 'y': True, 'z': True})
        xs = 0.2 * 720 * um
        setter = ies_b200.source.Setter(sp, (xs, 0, 0), (xs, 512 * um, 512 * um), (0, 0, 0))
        ies_b200.structure.Box('slab', sp, (160 * um, 0, 0), (260 * um, 512 * um, 512 * um), 4., 1.)
        sp.init_update_constants()
        rng = np.random.default_rng(3)
        for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
            a = rng.uniform(-1, 1, sp.loc_grid).astype(np.float32 if np.dtype(dt_).itemsize in (4, 8) and np.dtype(dt_) in (np.dtype('float32'), np.dtype('complex64')) else np.float64)
            getattr(sp, n)[:, :, :] = a.astype(dt_)
        def step(t):
            setter.put_src('Ey', 0.01, 'soft')
            sp.updateH(t); sp.updateE(t)
        # the clock sampler (an nvidia-smi process) is started first and given time to attach: its
        # start-up perturbs kernel launches for tens of ms, longer than a whole small-grid run
        clk = bench.ClockSampler(0)
        clk.start()
        import time
        time.sleep(0.4)
        for t in range(args.warmup): step(t)
        sp.sync()
        ms = C.c_double()
        _lib.check(lib.ies_timer_start(sp._ctx))
        for t in range(3): step(t)
        _lib.check(lib.ies_timer_stop(sp._ctx, C.byref(ms)))
        nsteps = max(args.steps, int(np.ceil(250. / max(ms.value / 3, 1e-3))))     # >= 0.25 s of timed work
        nsteps = min(nsteps, 5000)
        _lib.check(lib.ies_timer_start(sp._ctx))
        for t in range(nsteps): step(t)
        _lib.check(lib.ies_timer_stop(sp._ctx, C.byref(ms)))
        clocks = clk.stop()
        per = ms.value / nsteps
        ncell = nx * ny * nz
        g = ncell / per / 1e6
        finite = bool(np.all(np.isfinite(np.asarray(sp.Ey[nx // 2, :4, :4]))))
        row = dict(method=label, dtype=np.dtype(dt_).name, grid=list(grid), ms_per_step=round(per, 4), gcell_s=round(g, 2),
                   gbs_algorithmic=round(g * BYTES[dt_], 0), frac_of_hbm_peak=round(g * BYTES[dt_] / peak, 3),
                   finite=finite, steps=nsteps, clocks=clocks)
        rows.append(row)
        print(json.dumps(row), flush=True)
        del sp, setter
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    if args.only < 0 and not args.grep:
        json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'methods.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
