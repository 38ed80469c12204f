"""Copy the small pieces of the reference's SHIPPED result fixtures that the tests use
(graph/simple_2slab_{SHPF,FDTD}/.../Sx/*_area.npy, freqs.npy, sim_data.json) into
tests/golden/shipped/.  Run in the build container only; these are outputs of the
reference's own cupy/GPU runs, not source code."""
import os
import shutil

import numpy as np

REF = '/root/reference/graph'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'shipped')
RUNS = {
    'SHPF': 'simple_2slab_SHPF/0720um0512um0512um_0360_0016_0032_0015000_100um_200um_100um',
    'FDTD': 'simple_2slab_FDTD/0720um0512um0512um_0360_0020_0030_0015000_100um_200um_100um',
}

if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    for m, d in RUNS.items():
        src = os.path.join(REF, d)
        shutil.copy(os.path.join(src, 'sim_data.json'), os.path.join(OUT, f'{m}_sim_data.json'))
        shutil.copy(os.path.join(src, 'freqs.npy'), os.path.join(OUT, f'{m}_freqs.npy'))
        for name in ('TF_R', 'IF_R', 'SF_L'):
            for t in (1000, 3000, 15000):
                f = os.path.join(src, 'Sx', f'{name}_{t:07d}tstep_area.npy')
                shutil.copy(f, os.path.join(OUT, f'{m}_{name}_{t:07d}tstep_area.npy'))
            # one column of the DFT planes (the problem is uniform in y,z): (165,) complex128
            for comp in ('Ey', 'Hz'):
                a = np.load(os.path.join(src, 'Sx', f'{name}_DFT_{comp}_0003000tstep_rank00.npy'))
                np.save(os.path.join(OUT, f'{m}_{name}_DFT_{comp}_0003000tstep_col00.npy'), a[:, 0, 0])
    print(sorted(os.listdir(OUT)))
