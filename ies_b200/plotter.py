"""plotter.Graphtool for the b200 engine (reference: plotter.py:10-262).

`gather(what)` is the part scripts depend on: it returns the global field on
rank 0 (device -> host copy of each slab, gathered over the communicator).
`plot2D3D` draws a simple slice figure when matplotlib is importable and is a
no-op otherwise (pure visualisation, out of the hot-path scope).
"""
import os

import numpy as np


class Graphtool(object):

    def __init__(self, Space, name, path):
        self.Space = Space
        self.name = name
        self.savedir = path
        if self.Space.MPIrank == 0:
            if not os.path.exists(self.savedir): os.makedirs(self.savedir, exist_ok=True)

    def gather(self, what):
        """plotter.py:31-82."""
        local = np.asarray(getattr(self.Space, what))
        self.what = what
        if self.Space.MPIsize == 1:
            gathered = [local]
        else:
            gathered = self.Space.MPIcomm.gather(local, root=0)
        if self.Space.MPIrank == 0:
            self.integrated = np.zeros((self.Space.grid), dtype=self.Space.field_dtype)
            for MPIrank in range(self.Space.MPIsize):
                self.integrated[self.Space.myNx_slices[MPIrank], :, :] = gathered[MPIrank]
            return self.integrated
        return None

    def plot2D3D(self, integrated, tstep, xidx=None, yidx=None, zidx=None, **kwargs):
        """plotter.py:84-262 (reduced: one 2-D slice image)."""
        if self.Space.MPIrank != 0 or integrated is None:
            return
        try:
            import matplotlib
            matplotlib.use('Agg')
            import matplotlib.pyplot as plt
        except Exception:
            return
        if xidx is not None: plane = integrated[xidx, :, :]
        elif yidx is not None: plane = integrated[:, yidx, :]
        elif zidx is not None: plane = integrated[:, :, zidx]
        else: raise ValueError("Plane is not defined. Please insert one of x,y or z index of the plane.")
        fig, ax = plt.subplots(1, 1, figsize=kwargs.get('figsize', (8, 6)))
        im = ax.imshow(np.real(plane).T, cmap=kwargs.get('colordeep', 'bwr'), origin='lower', aspect='auto')
        fig.colorbar(im)
        ax.set_title(f"{self.what} at tstep {tstep}")
        folder = os.path.join(self.savedir, 'plot2D3D')
        os.makedirs(folder, exist_ok=True)
        fig.savefig(os.path.join(folder, f"{self.name}_{self.what}_{tstep:07d}.png"))
        plt.close(fig)
