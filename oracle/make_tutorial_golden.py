"""Golden outputs of the reference's own tutorial script -- TEST INFRASTRUCTURE, build container only.

    python -m oracle.make_tutorial_golden [tsteps Nx Ny Nz]      (default: BASELINE config 1 = 1000 128 32 32)

Executes /root/reference/tutorials/RT_simple_slabs.py on the UNMODIFIED reference modules
(oracle/ref_shims.py: cupy->NumPy alias, single-rank mpi4py stub, matplotlib stub; argv
`SHPF cupy <tsteps> <Nx> <Ny> <Nz>`, the only correct SHPF branch of the reference).
The one change to the script text is its hard-coded output root `/root/SHPF/` (lines 10, 128),
re-pointed to a scratch directory inside this repository because nothing outside it may be
written here; tests/test_gpu_tutorial.py runs the text unchanged on the GPU box.

Stores tests/golden/tutorial_rt_<Nx>_<Ny>_<Nz>_<tsteps>.npz: per collector the Poynting spectrum
`*_area.npy` and a sub-sample of the DFT planes written at the last cal_per step.
"""
import glob
import os
import shutil
import sys

import numpy as np

from . import ref_shims as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(R.REF, 'tutorials', 'RT_simple_slabs.py')


def collect(sx_dir):
    out = {}
    for f in sorted(glob.glob(os.path.join(sx_dir, '*_area.npy'))):
        out[os.path.basename(f)[:-4]] = np.load(f)
    for f in sorted(glob.glob(os.path.join(sx_dir, '*_DFT_*_rank00.npy'))):
        out[os.path.basename(f)[:-4]] = np.load(f)[:, ::8, ::8]
    return out


def main(argv):
    tsteps, nx, ny, nz = (argv + ['1000', '128', '32', '32'])[:4] if len(argv) < 4 else argv[:4]
    scratch = os.path.join(ROOT, 'gpurun_out', '_tutorial_ref') + os.sep
    shutil.rmtree(scratch, ignore_errors=True)
    os.makedirs(scratch)
    ns = R.load_reference(extra=('plotter', 'recorder'))
    for name in ('space', 'source', 'collector', 'structure', 'plotter', 'recorder'):
        sys.modules[name] = getattr(ns, name)
    src = open(SCRIPT).read().replace('/root/SHPF/', scratch)
    old_argv = sys.argv
    sys.argv = [SCRIPT, 'SHPF', 'cupy', str(tsteps), str(nx), str(ny), str(nz)]
    try:
        exec(compile(src, SCRIPT, 'exec'), {'__name__': '__main__', '__file__': SCRIPT})
    finally:
        sys.argv = old_argv
    sx = glob.glob(os.path.join(scratch, 'graph', 'simple_2slab_SHPF', '*', 'Sx'))[0]
    data = collect(sx)
    dst = os.path.join(ROOT, 'tests', 'golden', f'tutorial_rt_{int(nx)}_{int(ny)}_{int(nz)}_{int(tsteps)}.npz')
    np.savez_compressed(dst, **data)
    print('wrote', dst, {k: v.shape for k, v in data.items()})
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == '__main__':
    main(sys.argv[1:])
