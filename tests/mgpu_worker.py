"""torchrun worker: one rank per GPU, x-slab decomposition over NCCL (comm.TorchComm);
rank 0 gathers the fields and compares them with the oracle.  Launched by
tests/test_gpu_multi.py."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import ies_b200
    from oracle import cases as C
    ns = types.SimpleNamespace(space=ies_b200.space, source=ies_b200.source,
                               structure=ies_b200.structure, collector=ies_b200.collector)
    worst = 0.0
    for name in sys.argv[1:]:
        case = dict(C.CASES_BY_NAME[name])
        case['ranks'] = 1                      # build_api builds THIS rank's slab; size comes from the comm
        sp, setter = C.build_api(ns, case, 'b200')
        assert sp.MPIsize == dist.get_world_size()
        for t in range(case['steps']):
            C.step_api(sp, setter, case, t)
        g = ies_b200.plotter.Graphtool(sp, 'g', '/tmp/ies_mgpu/')
        got = {n: g.gather(n) for n in C.FIELDS}
        if dist.get_rank() == 0:
            ocase = dict(C.CASES_BY_NAME[name]); ocase['ranks'] = dist.get_world_size()
            want = C.run_oracle(ocase)
            den = max(np.linalg.norm(want[n].ravel()) for n in C.FIELDS)
            err = max(np.linalg.norm((got[n] - want[n]).ravel()) for n in C.FIELDS) / den
            print(f"MGPU {name} world={dist.get_world_size()} rel-L2 {err:.3e}", flush=True)
            worst = max(worst, err)
    dist.barrier()
    dist.destroy_process_group()
    if worst > 1e-10:
        sys.exit(3)


if __name__ == '__main__':
    main()
