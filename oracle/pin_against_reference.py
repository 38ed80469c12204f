"""Pin the oracle against the REAL reference, and write tests/golden/ fixtures.

Run in the build container only (needs /root/reference):

    python -m oracle.pin_against_reference            # all small golden cases
    python -m oracle.pin_against_reference shpf_f64_xpml
    python -m oracle.pin_against_reference --digest   # the large DIGEST_CASES (minutes)

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
    assert not case.get('probe')

    outs = R.run_ranks(case['ranks'], work)
    return {n: np.concatenate([o[n] for o in outs], axis=0) for n in C.FIELDS}


def main(argv):
    os.makedirs(GOLD, exist_ok=True)
    os.makedirs(os.path.join(GOLD, 'digest'), exist_ok=True)
    if argv and argv[0] == '--digest':
        names = argv[1:] or [k['name'] for k in C.DIGEST_CASES]
    else:
        names = argv or [k['name'] for k in C.CASES]
    digest_names = {k['name'] for k in C.DIGEST_CASES}
    rep_path = os.path.join(GOLD, 'pin_report.json')
    report = json.load(open(rep_path)) if os.path.exists(rep_path) else {}
    worst = 0.0
    for name in names:
        case = C.CASES_BY_NAME[name]
        ref = run_reference(case)
        ora = C.run_oracle(case)
        if case.get('nrank_quirk'):
            # the reference's own N-rank run deviates from its single-rank run on the slab-edge
            # planes (waived quirk): the oracle is pinned to the SINGLE-rank golden instead
            base = np.load(os.path.join(GOLD, case['golden'] + '.npz'))
            dev = max(C.rel_l2(ref[n], base[n]) for n in C.FIELDS)
            e = max(C.rel_l2(ora[n], base[n]) for n in C.FIELDS)
            status = 'OK' if e <= 1e-13 else 'FAIL'
            print(f"{name:28s} oracle(N rank) vs 1-rank golden {e:.3e}  {status}; reference(N rank) vs 1-rank golden {dev:.3e} (waived)")
            report[name] = dict(oracle_vs_single_rank_golden=e, reference_nrank_vs_single_rank=dev, status=status)
            worst = max(worst, e / 1e-13)
            continue
        errs = C.group_rel_l2(ora, ref) if name in digest_names else {n: C.rel_l2(ora[n], ref[n]) for n in C.FIELDS}
        if 'probe' in ref:
            errs['probe'] = C.rel_l2(ora['probe'], ref['probe'])
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
        if name in digest_names:
            np.savez_compressed(os.path.join(GOLD, 'digest', name + '.npz'), **C.digest(ref, case))
            continue
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
