"""Multi-process worker: one rank per GPU, x-slab decomposition, halo planes over CUDA IPC
(comm.IpcComm: copy-engine pushes into peer-mapped buffers, stream-ordered flags); rank 0 gathers
the fields and compares them with the N-rank oracle.  Launched by tests/test_gpu_multi.py through
`python -m torch.distributed.run` (used as a process launcher only: it sets RANK / WORLD_SIZE /
LOCAL_RANK / MASTER_PORT; the worker imports neither torch nor MPI)."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ies_b200
    from ies_b200 import comm
    from oracle import cases as C
    world = int(os.environ['WORLD_SIZE'])
    rank = int(os.environ['RANK'])
    ns = types.SimpleNamespace(space=ies_b200.space, source=ies_b200.source,
                               structure=ies_b200.structure, collector=ies_b200.collector)
    worst = 0.0
    for name in sys.argv[1:]:
        case = dict(C.CASES_BY_NAME[name])
        case['ranks'] = 1                      # build_api builds THIS rank's slab; size comes from the comm
        sp, setter = C.build_api(ns, case, 'b200')
        assert sp.MPIsize == world and isinstance(sp.MPIcomm, comm.IpcComm)
        for t in range(case['steps']):
            C.step_api(sp, setter, case, t)
        g = ies_b200.plotter.Graphtool(sp, 'g', '/tmp/ies_mgpu/')
        got = {n: g.gather(n) for n in C.FIELDS}
        if rank == 0:
            ocase = dict(C.CASES_BY_NAME[name]); ocase['ranks'] = world
            want = C.run_oracle(ocase)
            err = max(C.group_rel_l2(got, want).values())
            print(f"MGPU {name} world={world} rel-L2 {err:.3e}", flush=True)
            worst = max(worst, err)
        sp.MPIcomm.Barrier()
    if worst > 1e-10:
        sys.exit(3)


if __name__ == '__main__':
    main()
