"""GPU: BASELINE config 1 as stated -- the reference's tutorials/RT_simple_slabs.py executed
UNCHANGED with only the engine string switched (`SHPF b200 1000 128 32 32`), its written
Sx/*.npy compared with the outputs of the same script on the real reference
(oracle/make_tutorial_golden.py -> tests/golden/tutorial_rt_128_32_32_1000.npz).

The script text travels to the GPU box in the git-ignored oracle/_ref/ (oracle/make_ref.py); it
hard-codes `/root/SHPF/` as library and output root (lines 10, 128), so the test points that path
at a scratch directory for the duration of the run.  Field dtype is complex64 in the script
(lines 55-57): tolerance 1e-4 (north_star, fp32)."""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, 'oracle', '_ref', 'tutorials', 'RT_simple_slabs.py')
GOLD = os.path.join(ROOT, 'tests', 'golden', 'tutorial_rt_128_32_32_1000.npz')


def _rel(a, b):
    den = np.linalg.norm(np.asarray(b).ravel())
    return float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / den) if den > 0 else float(np.abs(a).max())


@pytest.mark.parametrize('tsteps', [1000, 3000])
def test_reference_tutorial_script_runs_unchanged_and_matches_reference_outputs(tsteps):
    """1000 steps = BASELINE config 1 as stated (the Gaussian peaks at step 2000, so the spectra are the
    pulse's leading tail: a deterministic vector, physically empty); 3000 steps = the same script once
    the pulse has passed the collectors (R + T spectra with weight)."""
    gold_path = GOLD.replace('_1000.npz', f'_{tsteps}.npz')
    if not os.path.isfile(gold_path):
        pytest.skip(f'no golden for {tsteps} steps')
    if not os.path.isfile(SCRIPT):
        pytest.skip('oracle/_ref/ (the reference script text) did not travel to this box')
    made_link = False
    scratch = tempfile.mkdtemp(prefix='ies_shpf_')
    try:
        if os.path.lexists('/root/SHPF'):
            if not os.path.isdir('/root/SHPF'):
                pytest.skip('/root/SHPF exists and is not a directory')
            out_root = '/root/SHPF'
        else:
            os.symlink(scratch, '/root/SHPF')
            made_link, out_root = True, scratch
        r = subprocess.run([sys.executable, '-m', 'ies_b200.compat.run', SCRIPT, 'SHPF', 'b200', str(tsteps), '128', '32', '32'],
                           cwd=ROOT, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
        run_dir = glob.glob(os.path.join(out_root, 'graph', 'simple_2slab_SHPF', f'*_0128_0032_0032_{tsteps:07d}_*'))
        assert len(run_dir) == 1, run_dir
        sx = os.path.join(run_dir[0], 'Sx')
        gold = np.load(gold_path)
        assert os.path.isfile(os.path.join(run_dir[0], 'sim_data.json'))
        checked = 0
        for key in gold.files:
            got = np.load(os.path.join(sx, key + '.npy'))
            want = gold[key]
            if '_DFT_' in key:
                assert got.shape == (want.shape[0], 32, 32) and got.dtype == np.complex128
                got = got[:, ::8, ::8]
            else:
                assert got.shape == want.shape
            assert _rel(got, want) <= 1e-4, (key, _rel(got, want))
            checked += 1
        assert checked == 15 * (tsteps // 1000)        # 3 collectors x (4 DFT planes + area) per cal_per
    finally:
        if made_link:
            os.unlink('/root/SHPF')
        shutil.rmtree(scratch, ignore_errors=True)
