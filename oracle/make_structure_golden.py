"""Rasterise Box / Sphere / Cylinder3D with the REAL reference (container only) and save
tests/golden/structures.npz for tests/test_host_logic.py."""
import os

import numpy as np

from . import ref_shims as R

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

if __name__ == '__main__':
    ns = R.load_reference()
    grid, gap = (48, 20, 18), (15e-6, 25.6e-6, 28.4e-6)
    dt = 0.25 * min(gap) / 299792458.0
    sp = ns.space.Basic3D(grid, gap, dt, 10, np.float64, np.complex128, method='FDTD', engine='cupy')
    sp.malloc()
    ns.structure.Box('b', sp, (60e-6, 0, 0), (150e-6, 300e-6, 512e-6), 4., 1.)
    ns.structure.Sphere('s', sp, (24, 10, 9), 120e-6, 2.25, 1.5)
    ns.structure.Cylinder3D('c', sp, 'x', 90e-6, (400e-6, 560e-6), (256e-6, 256e-6), 6., 1.)
    ns.structure.Cylinder3D('d', sp, 'y', 60e-6, (100e-6, 400e-6), (620e-6, 300e-6), 3., 2.)
    np.savez_compressed(os.path.join(GOLD, 'structures.npz'), eps=sp.eps_Ex, mu=sp.mu_Hx,
                        grid=np.array(grid), gap=np.array(gap))
    print('eps values:', np.unique(sp.eps_Ex / 8.8541878128e-12).round(3))
