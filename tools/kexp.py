#!/usr/bin/env python
"""Kernel experiments on the headline workload (development tool, GPU box only).

    python tools/kexp.py [--nx 1024] [--steps 10] [--warmup 3] [--only NAME[,NAME]] [--list]

Runs the bench workload of bench.py under several engine option sets
(ies_set_option) in ONE process, re-uploading the same initial fields before each
variant, and prints ms/step, Gcell/s and whether the result is bit-identical to the
default path.  Under ncu use --only NAME --steps 1 --warmup 0.
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
import bench  # noqa: E402

VARIANTS = [
    ('base', dict(fused=0)),                            # two-kernel path (k_zline + k_yline_update)
    ('fused', dict(fused=1)),                           # single launch, full-size scratch, defaults
    ('fused_split', dict(fused=1, pml_split=1)),
    ('fused_nosplit', dict(fused=1, pml_split=0)),
    ('fused_nosplit_zb1_l3', dict(fused=1, pml_split=0, fused_zb=1, fused_lead=3)),
    ('base_split', dict(fused=0, pml_split=1)),
    ('fused_pf', dict(fused=1, fused_prefetch=1)),
    ('fused_l3', dict(fused=1, fused_lead=3)),
    ('fused_l10', dict(fused=1, fused_lead=10)),
    ('fused_zb1', dict(fused=1, fused_zb=1)),
    ('fused_zb1_l3', dict(fused=1, fused_zb=1, fused_lead=3)),
    ('fused_r32', dict(fused=1, fused_ring=32)),        # scratch ring of 32 planes
    ('noctile', dict(fused=0, ctile=0)),
    ('palette', dict(fused=0, palette=1, ctile=0)),
]
ALL_OPTS = ('fused', 'fused_ring', 'fused_lead', 'fused_zb', 'fused_prefetch', 'pml_split', 'palette', 'ctile')
DEFAULTS = dict(fused=-1, fused_ring=0, fused_lead=6, fused_zb=2, fused_prefetch=0, pml_split=-1, palette=0, ctile=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nx', type=int, default=0)
    ap.add_argument('--config', default='headline')
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--only', default='')
    ap.add_argument('--list', action='store_true')
    ap.add_argument('--check-steps', type=int, default=2)
    args = ap.parse_args()
    if args.list:
        for n, o in VARIANTS: print(n, o)
        return
    import ies_b200
    from ies_b200 import _lib
    lib = _lib.load()
    ns = types.SimpleNamespace(space=ies_b200.space, source=ies_b200.source,
                               structure=ies_b200.structure, collector=ies_b200.collector)
    wl = bench.WORKLOADS[args.config]
    g1 = bench.grid_of(wl, 1)
    sp, setter, src = bench.build_space(ns, wl, 1, 100000, grid=(args.nx if args.nx else g1[0], g1[1], g1[2]))
    sp.init_update_constants()
    rng = np.random.default_rng(7)
    init = {n: rng.uniform(-1, 1, sp.loc_grid) for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')}
    ncell = sp.myNx * g1[1] * g1[2]

    def upload():
        _lib.check(lib.ies_set_option(sp._ctx, b'reset_psi', 1))
        for n, a in init.items():
            _lib.check(lib.ies_set_field(sp._ctx, _lib.COMP[n], _lib.I3(0, 0, 0), _lib.I3(*sp.loc_grid),
                                         a.ctypes.data_as(C.c_void_p)))

    def setopts(o):
        full = dict(DEFAULTS); full.update(o)
        for k in ALL_OPTS:
            _lib.check(lib.ies_set_option(sp._ctx, k.encode(), int(full[k])))

    def step(t):
        setter.put_src('Ey', src.pulse_re(t), 'soft')
        sp.updateH(t)
        sp.updateE(t)

    def signature():
        out = []
        for n in ('Ey', 'Hz', 'Ex'):
            for i in (0, 5, sp.myNx // 2, sp.myNx - 1):
                out.append(np.asarray(getattr(sp, n)[i, :, :]).copy())
        return np.stack(out)

    only = [x for x in args.only.split(',') if x]
    ref_sig = None
    rows = []
    for name, o in VARIANTS:
        if only and name not in only and name != 'base':
            continue
        try:
            setopts(o)
            sig_ok = None
            if args.check_steps > 0:
                upload()
                for t in range(args.check_steps): step(t)
                sp.sync()
                sig = signature()
                if name == 'base': ref_sig = sig
                sig_ok = bool(np.array_equal(sig, ref_sig)) if ref_sig is not None else None
                if ref_sig is not None and not sig_ok:
                    print('max abs diff vs base', float(np.max(np.abs(sig - ref_sig))), 'per plane',
                          [float(np.max(np.abs(a - b))) for a, b in zip(sig, ref_sig)], flush=True)
            for t in range(args.warmup): step(t)
            sp.sync()
            _lib.check(lib.ies_timer_start(sp._ctx))
            for t in range(args.steps): step(t)
            ms = C.c_double()
            _lib.check(lib.ies_timer_stop(sp._ctx, C.byref(ms)))
            sp.sync()
            per = ms.value / max(args.steps, 1)
            if o.get('fused'):
                _lib.check(lib.ies_set_option(sp._ctx, b'fused_prof', 1))
                for t in range(2): step(t)
                pr = (C.c_uint64 * 16)()
                _lib.check(lib.ies_fused_prof_read(sp._ctx, pr))
                _lib.check(lib.ies_set_option(sp._ctx, b'fused_prof', 0))
                nz_, ny_ = max(pr[8], 1), max(pr[9], 1)
                print('  cycles/tile: z fft %d, z ringwait %d, z store %d | y fft %d, y wait %d, y update %d' % (
                    pr[0] // nz_, pr[1] // nz_, pr[2] // nz_, pr[3] // ny_, pr[4] // ny_, pr[5] // ny_), flush=True)
            row = dict(name=name, ms_per_step=round(per, 4), gcell_s=round(ncell / per / 1e6, 2) if per > 0 else None,
                       bit_identical=sig_ok, opts=o)
        except Exception as e:      # keep sweeping
            row = dict(name=name, error=str(e), opts=o)
        rows.append(row)
        print(json.dumps(row), flush=True)
    setopts({})
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'kexp.json'), 'w') as f:
        json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
