"""structure.Box / Sphere / Cylinder3D for the b200 engine (reference: structure.py).

Structures rasterise eps_r / mu_r into the HOST material arrays of the space
(space.eps_E*, space.mu_H*, one shared array each -- space.py:207-218) before
init_update_constants() uploads the coefficients.  The inclusion predicates are
the reference's (structure.py:184-190, 469, 789, 831); the reference's O(N^3)
Python loops are evaluated as whole-array NumPy expressions with the same
floating-point operations, so the rasters are identical cell for cell.
"""
import numpy as np
from scipy.constants import c, mu_0, epsilon_0

try:
    from . import comm as _comm
except ImportError:  # sys.path drop-in
    import comm as _comm


class Structure:

    def __init__(self, name, space):
        self.name = name
        self.space = space

    def _get_local_x_loc(self, gxsrts, gxends):
        """structure.py:17-107: the one shared statement of the rule lives in comm.local_x_loc."""
        return _comm.local_x_loc(self.space, gxsrts, gxends)

    def _fill(self, index, mask=None):
        sp = self.space
        if mask is None:
            sp.eps[index] = self.eps_r * epsilon_0
            sp.mu[index] = self.mu_r * mu_0
        else:
            sp.eps[index][mask] = self.eps_r * epsilon_0
            sp.mu[index][mask] = self.mu_r * mu_0
        sp._dirty = True


class Box(Structure):
    """structure.py:108-192."""

    def __init__(self, name, space, srt, end, eps_r, mu_r):
        self.eps_r = eps_r
        self.mu_r = mu_r
        Structure.__init__(self, name, space)
        assert len(srt) == 3, "Only 3D material is possible."
        assert len(end) == 3, "Only 3D material is possible."
        xsrt = round(srt[0] / self.space.dx)
        ysrt = round(srt[1] / self.space.dy)
        zsrt = round(srt[2] / self.space.dz)
        xend = round(end[0] / self.space.dx)
        yend = round(end[1] / self.space.dy)
        zend = round(end[2] / self.space.dz)
        assert xsrt < xend
        assert ysrt < yend
        assert zsrt < zend
        self.gxloc, self.lxloc = Structure._get_local_x_loc(self, xsrt, xend)
        self.ysrt, self.yend = ysrt, yend
        self.zsrt, self.zend = zsrt, zend
        if self.gxloc != None:
            lxsrt, lxend = self.lxloc
            self._fill((slice(lxsrt, lxend), slice(ysrt, yend), slice(zsrt, zend)))
        return


class Sphere(Structure):
    """structure.py:395-480: `center` is in grid indices, `radius` in length units."""

    def __init__(self, name, space, center, radius, eps_r, mu_r):
        Structure.__init__(self, name, space)
        assert len(center) == 3, "Please insert x,y,z coordinate of the center."
        assert type(eps_r) == float, "Only isotropic media is possible. eps_r must be a single float."
        assert type(mu_r) == float, "Only isotropic media is possible.  mu_r must be a single float."
        self.eps_r = eps_r
        self.mu_r = mu_r
        self.radius = radius
        self.center_idx = center
        dx, dy, dz = self.space.dx, self.space.dy, self.space.dz
        gxsrt = center[0] - round(radius / dx)
        gxend = center[0] + round(radius / dx)
        assert gxsrt >= 0
        assert gxend < self.space.Nx
        self.gxloc, self.lxloc = Structure._get_local_x_loc(self, gxsrt, gxend)
        if self.gxloc != None:
            portion_srt = self.gxloc[0] - center[0] + round(radius / dx)
            portion_end = self.gxloc[1] - center[0] + round(radius / dx)
            self.portion = np.arange(portion_srt, portion_end)
            rx = abs(self.portion - round(radius / dx))
            theta = np.arccos(rx * dx / radius)
            rr = radius * np.sin(theta)
            j = np.arange(self.space.Ny)
            k = np.arange(self.space.Nz)
            d2 = (((j - center[1]) * dy) ** 2)[:, None] + (((k - center[2]) * dz) ** 2)[None, :]
            mask = d2[None, :, :] <= (rr ** 2)[:, None, None]
            self._fill((slice(self.lxloc[0], self.lxloc[1]), slice(None), slice(None)), mask)
        return


class Cylinder3D(Structure):
    """structure.py:722-845 (axis 'x' and 'y'; 'z' raises like the reference)."""

    def __init__(self, name, space, axis, radius, height, center, eps_r, mu_r):
        Structure.__init__(self, name, space)
        self.axis = axis
        self.radius = radius
        self.height = height
        self.center = center
        self.eps_r = eps_r
        self.mu_r = mu_r
        dx, dy, dz = self.space.dx, self.space.dy, self.space.dz
        self.rx = self.ry = self.rz = None
        j = np.arange(self.space.Ny)
        k = np.arange(self.space.Nz)
        if axis == 'x':
            self.ry = center[0] / dy
            self.rz = center[1] / dz
            gxsrts = round(height[0] / dx)
            gxends = round(height[1] / dx)
            self.gxloc, self.lxloc = Structure._get_local_x_loc(self, gxsrts, gxends)
            if self.gxloc != None:
                mask = ((((j - self.ry) * dy) ** 2)[:, None] + (((k - self.rz) * dz) ** 2)[None, :]) <= (radius ** 2)
                nxl = self.lxloc[1] - self.lxloc[0]
                self._fill((slice(self.lxloc[0], self.lxloc[1]), slice(None), slice(None)),
                           np.broadcast_to(mask[None], (nxl,) + mask.shape))
        elif axis == 'y':
            gxsrt = round(center[0] / dx) - round(radius / dx)
            gxend = round(center[0] / dx) + round(radius / dx)
            if gxsrt < 0: gxsrt = 0
            if gxend >= self.space.Nx: gxend = self.space.Nx - 1
            self.gxloc, self.lxloc = Structure._get_local_x_loc(self, gxsrt, gxend)
            if self.gxloc != None:
                portion_srt = self.gxloc[0] - round(center[0] / dx) + round(radius / dx)
                portion_end = self.gxloc[1] - round(center[0] / dx) + round(radius / dx)
                self.portion = np.arange(portion_srt, portion_end)
                rx = abs(self.portion - round(radius / dx)) * dx
                theta = np.arccos(rx / radius)
                rz = radius * np.sin(theta)
                inz = ((k * dz - center[1]) ** 2)[None, :] <= (rz ** 2)[:, None]
                iny = ((j * dy) >= height[0]) & ((j * dy) <= height[1])
                mask = inz[:, None, :] & iny[None, :, None]
                self._fill((slice(self.lxloc[0], self.lxloc[1]), slice(None), slice(None)), mask)
        elif axis == 'z':
            raise ValueError("Cylinder parallel to 'z' axis is not developed yet. Sorry.")
        return
