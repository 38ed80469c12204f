#!/usr/bin/env python
"""A/B of two builds of libies_b200.so on the same box (development tool):
    IES_B200_LIB=/path/to/other.so python tools/ab_lib.py [--config headline] [--steps 40]
Only entry points every build has are used."""
import argparse, ctypes as C, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='headline')
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--opt', action='append', default=[], help='name=value engine options')
    a = ap.parse_args()
    from ies_b200 import _lib
    lib = _lib.load()
    ns = bench.product_ns()
    wl = bench.WORKLOADS[a.config]
    sp, setter, src = bench.build_space(ns, wl, 1, 100000)
    sp.init_update_constants()
    for o in a.opt:
        k, v = o.split('=')
        _lib.check(lib.ies_set_option(sp._ctx, k.encode(), int(v)))
    rng = np.random.default_rng(7)
    for n in bench.FIELDS:
        arr = rng.uniform(-1, 1, sp.loc_grid)
        _lib.check(lib.ies_set_field(sp._ctx, _lib.COMP[n], _lib.I3(0, 0, 0), _lib.I3(*sp.loc_grid), arr.ctypes.data_as(C.c_void_p)))

    def step(t):
        setter.put_src('Ey', src.pulse_re(t), 'soft'); sp.updateH(t); sp.updateE(t)
    for t in range(a.warmup): step(t)
    sp.sync()
    ms = C.c_double()
    _lib.check(lib.ies_timer_start(sp._ctx))
    for t in range(a.steps): step(t)
    _lib.check(lib.ies_timer_stop(sp._ctx, C.byref(ms)))
    print(json.dumps(dict(lib=os.environ.get('IES_B200_LIB', 'in-tree'), opts=a.opt, ms_per_step=round(ms.value / a.steps, 4))))


if __name__ == '__main__':
    main()
