"""Helpers for the parity tests: run a case of oracle/cases.py through the product."""
import os

import numpy as np

from oracle import cases as C

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(case):
    z = np.load(os.path.join(GOLD, case['golden'] + '.npz'))
    return {n: z[n] for n in C.FIELDS}


def load_digest(case):
    z = np.load(os.path.join(GOLD, 'digest', case['golden'] + '.npz'))
    return {n: z[n] for n in z.files}


def tolerance(case):
    """north_star: relative L2 <= 1e-10 in fp64, <= 1e-4 in fp32."""
    return 1e-4 if np.dtype(case['dtype']) in (np.dtype('float32'), np.dtype('complex64')) else 1e-10


def run_product(ns, case):
    """Run a case on the GPU through the product's reference-style API.  Multi-rank
    cases run as several x-slabs of one process (comm.LocalGroup) on device 0."""
    if case['ranks'] == 1:
        return C.run_api(ns, case, 'b200')
    from ies_b200 import comm
    grp = comm.LocalGroup(case['ranks'])
    (Lx, Ly, Lz), gap, dt = C.geometry(case)
    fd = np.dtype(case['dtype']).type
    md = np.dtype(case['mmtdtype']).type
    spaces, setters = [], []
    for r in range(case['ranks']):
        sp = ns.space.Basic3D(case['grid'], gap, dt, case['steps'] + 1, fd, md,
                              method=case['method'], engine='b200', comm=grp.comm(r))
        sp.malloc()
        sp.apply_PML(case['pml'], case['npml'])
        if case['bbc'] is not None: sp.apply_BBC(case['bbc'])
        if case['pbc'] is not None: sp.apply_PBC(case['pbc'])
        s0, s1 = C.source_box(case)
        setters.append(ns.source.Setter(sp, s0, s1, case['mmt']))
        for (b0, b1, er, mr) in C.box_list(case):
            ns.structure.Box('box', sp, b0, b1, er, mr)
        if C.sphere_spec(case) is not None:
            ns.structure.Sphere('sphere', sp, *C.sphere_spec(case))
        sp.init_update_constants()
        spaces.append(sp)
    for t in range(case['steps']):
        p = C.pulse_value(case, t, dt)
        for s in setters: s.put_src(case['src_field'], p, case['put'])
        for sp in spaces: sp.updateH(t)
        for sp in spaces: sp.updateE(t)
    return {n: np.concatenate([np.asarray(getattr(sp, n)) for sp in spaces], axis=0) for n in C.FIELDS}


def worst_rel_l2(got, want):
    return C.group_rel_l2(got, want)
