"""Pin the oracle against the REAL reference, and write tests/golden/ fixtures.

Run in the build container only (needs /root/reference):

    python -m oracle.pin_against_reference            # all cases
    python -m oracle.pin_against_reference shpf_f64_xpml

For every case in oracle/cases.py it runs the unmodified reference modules
(oracle/ref_shims.py: cupy->numpy alias, fake mpi4py; N-rank cases on N threads)
and the NumPy restatement (oracle/ies_oracle.py) on identical inputs, prints the
per-field max-abs and relative-L2 differences, fails if any exceeds the bound,
and stores the REFERENCE's fields as tests/golden/<case>.npz.
"""
import json
import os
import sys

import numpy as np

from . import cases as C
from . import ref_shims as R

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def run_reference(case):
    ns = R.load_reference()
    if case['ranks'] == 1:
        return C.run_api(ns, case, 'cupy')

    def work(rank):
        sp, setter = C.build_api(ns, case, 'cupy')
        for t in range(case['steps']):
            C.step_api(sp, setter, case, t)
        return {n: np.asarray(getattr(sp, n)) for n in C.FIELDS}

    outs = R.run_ranks(case['ranks'], work)
    return {n: np.concatenate([o[n] for o in outs], axis=0) for n in C.FIELDS}


def main(argv):
    os.makedirs(GOLD, exist_ok=True)
    names = argv or [k['name'] for k in C.CASES]
    report = {}
    worst = 0.0
    for name in names:
        case = C.CASES_BY_NAME[name]
        ref = run_reference(case)
        ora = C.run_oracle(case)
        errs = {n: C.rel_l2(ora[n], ref[n]) for n in C.FIELDS}
        amax = max(float(np.abs(ref[n]).max()) for n in C.FIELDS)
        e = max(errs.values())
        # Q3: the reference's slab-edge coefficient quirk makes N-rank != 1-rank when an
        # interface sits on a slab edge; multi-rank cases are compared with a looser bound.
        single = np.dtype(case['dtype']) in (np.dtype('float32'), np.dtype('complex64'))
        bound = 1e-5 if single else 1e-13
        if case['ranks'] > 1 and not single:
            bound = 1e-13
        status = 'OK' if e <= bound else 'FAIL'
        print(f"{name:28s} max rel-L2 {e:.3e}  (|field|max {amax:.3e})  {status}")
        report[name] = dict(rel_l2=errs, field_absmax=amax, bound=bound, status=status)
        worst = max(worst, e / bound)
        if case['golden'] != name:
            # N-rank reference run must equal the single-rank golden bit for bit
            base = np.load(os.path.join(GOLD, case['golden'] + '.npz'))
            same = all(np.array_equal(base[n], ref[n]) for n in C.FIELDS)
            report[name]['equals_golden'] = case['golden'] if same else False
            print(f"{'':28s} N-rank reference == golden '{case['golden']}': {same}")
            if not same:
                worst = 2.0
            continue
        np.savez_compressed(os.path.join(GOLD, name + '.npz'),
                            **{n: ref[n].astype(np.dtype(case['dtype'])) if case['method'] != 'PSTD'
                               else ref[n] for n in C.FIELDS})
    with open(os.path.join(GOLD, 'pin_report.json'), 'w') as f:
        json.dump(report, f, indent=1, sort_keys=True)
    if worst > 1.0:
        sys.exit(1)


if __name__ == '__main__':
    main(sys.argv[1:])
