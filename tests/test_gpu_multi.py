"""GPU, >= 2 devices: the real multi-process path (one rank per GPU, CUDA IPC halo exchange)
against the N-rank oracle.  Skipped on a single-GPU box; the slab logic itself
is also covered on one GPU by the LocalGroup cases of test_gpu_parity.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('world', [2, 4])
def test_multiprocess_ipc_halo_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    # the 256-point-line case has 12 planes and a 4-cell x absorber: two slabs of 6
    cases = ['shpf_f64_allpml_64_r2', 'fdtd_f64_xpml_pbc_r4'] + (['shpf_f64_allpml_256_r2'] if world == 2 else [])
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(29700 + world),
           os.path.join(ROOT, 'tests', 'mgpu_worker.py')] + cases
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('MGPU') == len(cases), r.stdout[-2000:]
