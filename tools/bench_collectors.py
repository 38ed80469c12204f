#!/usr/bin/env python
"""Cost of the running-DFT collectors on the headline grid (development tool, GPU box only):
SHPF fp64 1024x256x256, three Sx collectors of 165 frequencies on 256x256 planes (the RT
tutorial's set-up, RT_simple_slabs.py:138), time per step with and without do_RFT."""
import ctypes as C
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import ies_b200
    from ies_b200 import _lib
    lib = _lib.load()
    ns = types.SimpleNamespace(space=ies_b200.space, source=ies_b200.source,
                               structure=ies_b200.structure, collector=ies_b200.collector)
    sp, setter, src = bench.build_space(ns, bench.WORKLOADS['headline'], 1, 100000)
    sp.init_update_constants()
    rng = np.random.default_rng(7)
    for n in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
        getattr(sp, n)[:, :, :] = rng.uniform(-1, 1, sp.loc_grid)
    um = 1e-6
    wv = np.arange(67, 100, 0.2) * um
    freqs = 299792458.0 / wv
    os.makedirs('/tmp/ies_coll/', exist_ok=True)
    cols = [ns.collector.Sx(f'c{q}', '/tmp/ies_coll/', sp, x * 720 * um, (0, 0), (512 * um, 512 * um), freqs, 'b200')
            for q, x in enumerate((0.15, 0.5, 0.85))]
    out = {}
    for mode in ('off', 'on'):
        def step(t):
            setter.put_src('Ey', src.pulse_re(t), 'soft')
            sp.updateH(t); sp.updateE(t)
            if mode == 'on':
                for c in cols: c.do_RFT(t)
        for t in range(4): step(t)
        sp.sync()
        K = 64
        _lib.check(lib.ies_timer_start(sp._ctx))
        for t in range(K): step(t)
        ms = C.c_double()
        _lib.check(lib.ies_timer_stop(sp._ctx, C.byref(ms)))
        out[mode] = ms.value / K
    nf = len(freqs)
    print(json.dumps(dict(nf=nf, ms_per_step_off=round(out['off'], 4), ms_per_step_on=round(out['on'], 4),
                          collectors_ms_per_step=round(out['on'] - out['off'], 4),
                          unblocked_traffic_gb_per_step=round(3 * 4 * nf * 65536 * 32 / 1e9, 2))))


if __name__ == '__main__':
    main()
