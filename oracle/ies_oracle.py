"""CPU oracle for the IES updateH/updateE hot path -- TEST INFRASTRUCTURE ONLY.

This is a NumPy restatement of the reference's time-stepping algorithm
(/root/reference/space.py, source.py, collector.py).  It is NOT part of the
product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  The product path (ies_b200/) never
imports anything from oracle/.

Pinning: `oracle/pin_against_reference.py` imports the real reference modules
(cupy->numpy alias, single-rank mpi4py stub) in the build container and checks
this restatement against them step by step (max-abs difference 0.0 or 1 ulp on
every configuration listed there); it also writes the golden fixtures in
tests/golden/.  The shipped reference goldens (graph/simple_2slab_{SHPF,FDTD})
are reproduced in tests/test_oracle_golden.py.

Every function cites the reference lines it follows.  Layout and structure are
deliberately different from the reference (table-driven CPML, explicit halo
arguments, one coefficient array per half-step) -- this is a restatement of
the arithmetic, not a copy of the code.
"""
from __future__ import annotations

import numpy as np
from scipy.constants import c, mu_0, epsilon_0

_REAL = (np.float32, np.float64)
_CPLX = (np.complex64, np.complex128)


def _is_complex(dt):
    return np.dtype(dt).kind == 'c'


class OracleSpace:
    """One x-slab of the simulation space (reference: space.Basic3D, space.py:7-141).

    `rank`/`size` select the slab; halo planes are passed explicitly to
    update_h / update_e (reference: blocking MPI send/recv, space.py:645-670,
    863-887).
    """

    def __init__(self, grid, gridgap, dt, tsteps, field_dtype, mmtdtype,
                 method='SHPF', rank=0, size=1, courant=0.25):
        self.field_dtype = np.dtype(field_dtype).type
        self.mmtdtype = np.dtype(mmtdtype).type
        self.rank, self.size = rank, size
        self.grid = tuple(grid)
        self.Nx, self.Ny, self.Nz = self.grid
        self.dx, self.dy, self.dz = gridgap
        # space.py:73-75
        self.Lx = (self.Nx - 1) * self.dx
        self.Ly = (self.Ny - 1) * self.dy
        self.Lz = (self.Nz - 1) * self.dz
        self.dt = dt
        self.tsteps = tsteps
        self.method = method
        assert method in ('FDTD', 'SHPF', 'PSTD')
        if method == 'PSTD':
            assert size == 1                       # space.py:89
        assert float(self.Nx) % size == 0.         # space.py:108
        self.myNx = round(self.Nx / size)          # space.py:114
        self.loc_grid = (self.myNx, self.Ny, self.Nz)
        self.x0 = rank * self.myNx
        self.myNx_indice = [(r * self.myNx, (r + 1) * self.myNx) for r in range(size)]
        z = lambda: np.zeros(self.loc_grid, dtype=self.field_dtype)
        self.Ex, self.Ey, self.Ez = z(), z(), z()
        self.Hx, self.Hy, self.Hz = z(), z(), z()
        self.BBC_called = False
        self.PBC_called = False
        self.bbc = {'x': False, 'y': False, 'z': False}
        self.pbc = {'x': False, 'y': False, 'z': False}
        self.PMLregion = {}
        self.npml = 0
        self.mmt = None
        self._malloc()

    # ------------------------------------------------------------------ setup
    def _malloc(self):
        """FFT tables and materials (space.py:143-237, the engine=='cupy' branch
        168-181, which is the only correct one at this commit -- SURVEY Q1)."""
        cplx = _is_complex(self.field_dtype)
        if cplx:
            self._fft = lambda a, ax: np.fft.fftn(a, axes=(ax,))
            self._ifft = lambda a, ax: np.fft.ifftn(a, axes=(ax,))
            fftfreq = np.fft.fftfreq
        else:
            self._fft = lambda a, ax: np.fft.rfftn(a, axes=(ax,))
            self._ifft = lambda a, ax: np.fft.irfftn(a, axes=(ax,))
            fftfreq = np.fft.rfftfreq
        self.kx = fftfreq(self.Nx, self.dx) * 2 * np.pi
        self.ky = fftfreq(self.Ny, self.dy) * 2 * np.pi
        self.kz = fftfreq(self.Nz, self.dz) * 2 * np.pi
        md = self.mmtdtype
        self.ikx = (1j * self.kx[:, None, None]).astype(md)
        self.iky = (1j * self.ky[None, :, None]).astype(md)
        self.ikz = (1j * self.kz[None, None, :]).astype(md)
        self.xpshift = np.exp(self.ikx * +self.dx / 2).astype(md)
        self.xmshift = np.exp(self.ikx * -self.dx / 2).astype(md)
        self.ypshift = np.exp(self.iky * +self.dy / 2).astype(md)
        self.ymshift = np.exp(self.iky * -self.dy / 2).astype(md)
        self.zpshift = np.exp(self.ikz * +self.dz / 2).astype(md)
        self.zmshift = np.exp(self.ikz * -self.dz / 2).astype(md)
        # persistent derivative buffers: entries outside a method's sub-volume
        # are never written and stay 0 (space.py:193-205)
        self.d = {n: np.zeros(self.loc_grid, dtype=self.field_dtype) for n in
                  ('xEy', 'xEz', 'yEx', 'yEz', 'zEx', 'zEy',
                   'xHy', 'xHz', 'yHx', 'yHz', 'zHx', 'zHy')}
        # one shared eps and one shared mu array (space.py:207-218)
        self.eps = np.ones(self.loc_grid, dtype=np.float64) * epsilon_0
        self.mu = np.ones(self.loc_grid, dtype=np.float64) * mu_0

    def apply_PML(self, region, npml):
        """CPML profiles (space.py:239-361)."""
        self.PMLregion = dict(region)
        self.npml = npml
        g = 2 * npml
        rc0, imp, gO, sO = 1.e-16, np.sqrt(mu_0 / epsilon_0), 3., 3.
        self.pml = {}
        self.psi = {}
        for ax, dd in (('x', self.dx), ('y', self.dy), ('z', self.dz)):
            if region.get(ax, '') == '':
                continue
            bdw = (g - 1) * (dd / 2)
            smax = -(gO + 1) * np.log(rc0) / (2 * imp * bdw)
            loc = np.arange(g) / (g - 1)
            sigma = smax * (loc ** gO)
            kappa = 1 + ((7. - 1) * (loc ** gO))
            alpha = 0.05 * ((1 - loc) ** sO)
            b = np.exp(-(sigma / kappa + alpha) * self.dt / epsilon_0)
            a = sigma / (sigma * kappa + alpha * kappa ** 2) * (b - 1.)
            self.pml[ax] = dict(sigma=sigma, kappa=kappa, alpha=alpha, b=b, a=a)
            shape = {'x': (npml, self.Ny, self.Nz), 'y': (self.myNx, npml, self.Nz),
                     'z': (self.myNx, self.Ny, npml)}[ax]
            # psi names follow the reference: psi_<comp><axis>_<p|m> (space.py:293-335)
            comps = {'x': ('ey', 'ez', 'hy', 'hz'), 'y': ('ex', 'ez', 'hx', 'hz'),
                     'z': ('ex', 'ey', 'hx', 'hy')}[ax]
            for cmp_ in comps:
                for side in 'pm':
                    self.psi[f'{cmp_}{ax}_{side}'] = np.zeros(shape, dtype=self.field_dtype)

    def apply_BBC(self, region):
        """space.py:555-615."""
        self.bbc = {k: bool(region.get(k)) for k in 'xyz'}
        if True in region.values():
            self.BBC_called = True
            assert _is_complex(self.field_dtype)
        if self.bbc['x']:
            assert self.size == 1

    def apply_PBC(self, region):
        """space.py:617-637."""
        self.PBC_called = True
        self.pbc = {k: bool(region.get(k)) for k in 'xyz'}
        if self.pbc['x']:
            assert self.size == 1

    def init_update_constants(self):
        """space.py:445-553 with econ=mcon=0 (never written anywhere in the
        reference, space.py:223-234): C1 == 1 exactly, CH2 = -2dt/(2mu),
        CE2 = 2dt/(2eps).  Kept full-size; the per-method trims of the reference
        are applied by slicing at use."""
        self.CH2 = (-2 * self.dt) / (2. * self.mu)
        self.CE2 = (2. * self.dt) / (2. * self.eps)

    # ------------------------------------------------------------ derivatives
    def _spec(self, F, mult, ax):
        return self._ifft(mult * self._fft(F, ax), ax)

    def _ghost(self, F3, halo, first):
        """FDTD periodic / Bloch ghost-cell copies (space.py:1714-1858, 1981-2033).
        F3 = the three components; halo = (recv1, recv2) planes or None."""
        fx, fy, fz = F3
        for ax, L, dd in ((0, self.Lx, self.dx), (1, self.Ly, self.dy), (2, self.Lz, self.dz)):
            key = 'xyz'[ax]
            if self.bbc[key]:
                newL = L - 2 * dd
            elif self.pbc[key]:
                newL = 0
            else:
                continue
            k = self.mmt[ax]
            pp = np.exp(+1j * k * newL)
            pm = np.exp(-1j * k * newL)
            if not _is_complex(self.field_dtype):
                # reference multiplies a real array by exp(0j)=1+0j and assigns
                # into a real array (numpy discards the zero imaginary part with
                # a ComplexWarning); only k*newL == 0 is meaningful here.
                pp, pm = pp.real, pm.real
            sl = [slice(None)] * 3
            def ix(i):
                s = list(sl); s[ax] = i; return tuple(s)
            for f in (fx, fy, fz):
                f[ix(-1)] = f[ix(1)] * pp
            for f in (fy, fx, fz):
                f[ix(0)] = f[ix(-2)] * pm
            if ax > 0 and halo is not None:
                a2 = ax - 1
                def ix2(i):
                    s = [slice(None)] * 2; s[a2] = i; return tuple(s)
                for r in halo:
                    r[ix2(-1)] = r[ix2(1)] * pp
                for r in halo:
                    r[ix2(0)] = r[ix2(-2)] * pm

    def _ghost_x(self, F3, newL):
        """_exchange_BBCx with mpi=False (space.py:1714-1727): F[-1] = F[1] e^{+ikL'},
        F[0] = F[-2] e^{-ikL'} on the three components."""
        k = self.mmt[0]
        pp, pm = np.exp(+1j * k * newL), np.exp(-1j * k * newL)
        for f in F3:
            f[-1] = f[1] * pp
        for f in F3:
            f[0] = f[-2] * pm

    # ---------------------------------------------------------------- updateH
    def update_h(self, tstep, halo_E=None):
        """space.py:639-840.  halo_E = (Ey[0], Ez[0]) of rank+1 or None."""
        m = self.method
        Ex, Ey, Ez = self.Ex, self.Ey, self.Ez
        d = self.d
        dx, dy, dz = self.dx, self.dy, self.dz
        last = self.rank == self.size - 1
        if not last:
            assert halo_E is not None
            rEy, rEz = (np.array(h, copy=True) for h in halo_E)
        if m == 'FDTD':
            self._ghost((Ex, Ey, Ez), None if last else (rEy, rEz), first=False)
            # space.py:760-779
            d['yEz'][:, :-1, :-1] = (Ez[:, 1:, :-1] - Ez[:, :-1, :-1]) / dy
            d['zEy'][:, :-1, :-1] = (Ey[:, :-1, 1:] - Ey[:, :-1, :-1]) / dz
            d['zEx'][:-1, :, :-1] = (Ex[:-1, :, 1:] - Ex[:-1, :, :-1]) / dz
            d['xEz'][:-1, :, :-1] = (Ez[1:, :, :-1] - Ez[:-1, :, :-1]) / dx
            d['yEx'][:-1, :-1, :] = (Ex[:-1, 1:, :] - Ex[:-1, :-1, :]) / dy
            d['xEy'][:-1, :-1, :] = (Ey[1:, :-1, :] - Ey[:-1, :-1, :]) / dx
            if not last:
                d['zEx'][-1, :, :-1] = (Ex[-1, :, 1:] - Ex[-1, :, :-1]) / dz
                d['xEz'][-1, :, :-1] = (rEz[:, :-1] - Ez[-1, :, :-1]) / dx
                d['xEy'][-1, :-1, :] = (rEy[:-1, :] - Ey[-1, :-1, :]) / dx
                d['yEx'][-1, :-1, :] = (Ex[-1, 1:, :] - Ex[-1, :-1, :]) / dy
        elif m == 'SHPF':
            # space.py:709-727
            d['yEz'] = self._spec(Ez, self.iky * self.ypshift, 1)
            d['zEy'] = self._spec(Ey, self.ikz * self.zpshift, 2)
            d['zEx'] = self._spec(Ex, self.ikz * self.zpshift, 2)
            d['xEz'][:-1] = (Ez[1:] - Ez[:-1]) / dx
            d['yEx'] = self._spec(Ex, self.iky * self.ypshift, 1)
            d['xEy'][:-1] = (Ey[1:] - Ey[:-1]) / dx
            if not last:
                d['xEz'][-1] = (rEz - Ez[-1]) / dx
                d['xEy'][-1] = (rEy - Ey[-1]) / dx
        else:  # PSTD, space.py:732-741
            d['yEz'] = self._spec(Ez, self.iky, 1)
            d['zEy'] = self._spec(Ey, self.ikz, 2)
            d['zEx'] = self._spec(Ex, self.ikz, 2)
            d['xEz'] = self._spec(Ez, self.ikx, 0)
            d['yEx'] = self._spec(Ex, self.iky, 1)
            d['xEy'] = self._spec(Ey, self.ikx, 0)

        C = self.CH2
        if m == 'PSTD':
            # space.py:795-797 (whole-array rebinding; dtype promotes)
            self.Hx = 1. * self.Hx + C * (d['yEz'] - d['zEy'])
            self.Hy = 1. * self.Hy + C * (d['zEx'] - d['xEz'])
            self.Hz = 1. * self.Hz + C * (d['xEy'] - d['yEx'])
        elif m == 'SHPF':
            # space.py:801-811.  Plane myNx-1 of Hy/Hz exists only when a halo
            # plane was received; it uses that plane's OWN coefficient
            # (partition-invariant; the reference's CHy1[-1] quirk Q3 is waived).
            sx = slice(None) if not last else slice(None, -1)
            self.Hx[:] = self.Hx + C * (d['yEz'] - d['zEy'])
            self.Hy[sx] = self.Hy[sx] + C[sx] * (d['zEx'][sx] - d['xEz'][sx])
            self.Hz[sx] = self.Hz[sx] + C[sx] * (d['xEy'][sx] - d['yEx'][sx])
        else:
            # space.py:815-825
            sx = slice(None) if not last else slice(None, -1)
            s = (slice(None), slice(None, -1), slice(None, -1))
            self.Hx[s] = self.Hx[s] + C[s] * (d['yEz'][s] - d['zEy'][s])
            s = (sx, slice(None), slice(None, -1))
            self.Hy[s] = self.Hy[s] + C[s] * (d['zEx'][s] - d['xEz'][s])
            s = (sx, slice(None, -1), slice(None))
            self.Hz[s] = self.Hz[s] + C[s] * (d['xEy'][s] - d['yEx'][s])

        if self.BBC_called and m != 'FDTD':
            self._bloch_h()
        self._pml_h()

    def _bloch_h(self):
        """space.py:1893-1950."""
        dt, mu, mm = self.dt, self.mu, self.mmt
        if self.method == 'SHPF':
            # x axis: ghost-plane copies of E after the H update (space.py:1898-1912;
            # both branches are plain `if`s there)
            if self.bbc['x']:
                self._ghost_x((self.Ex, self.Ey, self.Ez), self.Lx - 2 * self.dx)
            if self.pbc['x']:
                self._ghost_x((self.Ex, self.Ey, self.Ez), 0)
            # [:-1] on the last rank; a rank with a successor has updated its plane
            # myNx-1 from the halo, and the partition-invariant result (== single rank)
            # needs the Bloch term there too.  The reference applies [:-1] on EVERY rank
            # (space.py:1896), a slab-edge quirk of the same kind as Q3 -- waived.
            s2 = slice(None, -1) if self.rank == self.size - 1 else slice(None)
            if self.bbc['y']:
                ez = self._spec(self.Ez, self.ypshift, 1)
                ex = self._spec(self.Ex, self.ypshift, 1)
                self.Hx[:] += -dt / mu * 1j * (-mm[1] * ez)
                self.Hz[s2] += -dt / mu[s2] * 1j * (+mm[1] * ex[s2])
            if self.bbc['z']:
                ey = self._spec(self.Ey, self.zpshift, 2)
                ex = self._spec(self.Ex, self.zpshift, 2)
                self.Hx[:] += -dt / mu * 1j * (+mm[2] * ey)
                self.Hy[s2] += -dt / mu[s2] * 1j * (-mm[2] * ex[s2])
        else:  # PSTD
            if self.bbc['x']:
                self.Hy = self.Hy + -dt / mu * 1j * (+mm[0] * self.Ez)
                self.Hz = self.Hz + -dt / mu * 1j * (-mm[0] * self.Ey)
            if self.bbc['y']:
                self.Hx = self.Hx + -dt / mu * 1j * (-mm[1] * self.Ez)
                self.Hz = self.Hz + -dt / mu * 1j * (+mm[1] * self.Ex)
            if self.bbc['z']:
                self.Hx = self.Hx + -dt / mu * 1j * (+mm[2] * self.Ey)
                self.Hy = self.Hy + -dt / mu * 1j * (-mm[2] * self.Ex)

    # ---------------------------------------------------------------- updateE
    def update_e(self, tstep, halo_H=None):
        """space.py:842-1052.  halo_H = (Hy[-1], Hz[-1]) of rank-1 or None."""
        m = self.method
        Hx, Hy, Hz = self.Hx, self.Hy, self.Hz
        d = self.d
        dx, dy, dz = self.dx, self.dy, self.dz
        first = self.rank == 0
        if not first:
            assert halo_H is not None
            rHy, rHz = (np.array(h, copy=True) for h in halo_H)
        if m == 'FDTD':
            self._ghost((Hx, Hy, Hz), None if first else (rHy, rHz), first=True)
            # space.py:975-994
            d['yHz'][:, 1:, 1:] = (Hz[:, 1:, 1:] - Hz[:, :-1, 1:]) / dy
            d['zHy'][:, 1:, 1:] = (Hy[:, 1:, 1:] - Hy[:, 1:, :-1]) / dz
            d['zHx'][1:, :, 1:] = (Hx[1:, :, 1:] - Hx[1:, :, :-1]) / dz
            d['xHz'][1:, :, 1:] = (Hz[1:, :, 1:] - Hz[:-1, :, 1:]) / dx
            d['yHx'][1:, 1:, :] = (Hx[1:, 1:, :] - Hx[1:, :-1, :]) / dy
            d['xHy'][1:, 1:, :] = (Hy[1:, 1:, :] - Hy[:-1, 1:, :]) / dx
            if not first:
                d['xHz'][0, :, 1:] = (Hz[0, :, 1:] - rHz[:, 1:]) / dx
                d['zHx'][0, :, 1:] = (Hx[0, :, 1:] - Hx[0, :, :-1]) / dz
                d['xHy'][0, 1:, :] = (Hy[0, 1:, :] - rHy[1:, :]) / dx
                d['yHx'][0, 1:, :] = (Hx[0, 1:, :] - Hx[0, :-1, :]) / dy
        elif m == 'SHPF':
            # space.py:953-970
            d['yHz'] = self._spec(Hz, self.iky * self.ymshift, 1)
            d['zHy'] = self._spec(Hy, self.ikz * self.zmshift, 2)
            d['zHx'] = self._spec(Hx, self.ikz * self.zmshift, 2)
            d['xHz'][1:] = (Hz[1:] - Hz[:-1]) / dx
            d['yHx'] = self._spec(Hx, self.iky * self.ymshift, 1)
            d['xHy'][1:] = (Hy[1:] - Hy[:-1]) / dx
            if not first:
                d['xHz'][0] = (Hz[0] - rHz) / dx
                d['xHy'][0] = (Hy[0] - rHy) / dx
        else:  # PSTD, space.py:903-912
            d['yHz'] = self._spec(Hz, self.iky, 1)
            d['zHy'] = self._spec(Hy, self.ikz, 2)
            d['zHx'] = self._spec(Hx, self.ikz, 2)
            d['xHz'] = self._spec(Hz, self.ikx, 0)
            d['yHx'] = self._spec(Hx, self.iky, 1)
            d['xHy'] = self._spec(Hy, self.ikx, 0)

        C = self.CE2
        if m == 'PSTD':
            # space.py:1011-1013
            self.Ex = 1. * self.Ex + C * (d['yHz'] - d['zHy'])
            self.Ey = 1. * self.Ey + C * (d['zHx'] - d['xHz'])
            self.Ez = 1. * self.Ez + C * (d['xHy'] - d['yHx'])
        elif m == 'SHPF':
            # space.py:1017-1025 (own-plane coefficient, see update_h)
            sx = slice(None) if not first else slice(1, None)
            self.Ex[:] = self.Ex + C * (d['yHz'] - d['zHy'])
            self.Ey[sx] = self.Ey[sx] + C[sx] * (d['zHx'][sx] - d['xHz'][sx])
            self.Ez[sx] = self.Ez[sx] + C[sx] * (d['xHy'][sx] - d['yHx'][sx])
        else:
            # space.py:1029-1037
            sx = slice(None) if not first else slice(1, None)
            s = (slice(None), slice(1, None), slice(1, None))
            self.Ex[s] = self.Ex[s] + C[s] * (d['yHz'][s] - d['zHy'][s])
            s = (sx, slice(None), slice(1, None))
            self.Ey[s] = self.Ey[s] + C[s] * (d['zHx'][s] - d['xHz'][s])
            s = (sx, slice(1, None), slice(None))
            self.Ez[s] = self.Ez[s] + C[s] * (d['xHy'][s] - d['yHx'][s])

        if self.BBC_called and m != 'FDTD':
            self._bloch_e()
        self._pml_e()

    def _bloch_e(self):
        """space.py:2068-2122."""
        dt, eps, mm = self.dt, self.eps, self.mmt
        if self.method == 'SHPF':
            # x axis: ghost-plane copies of H after the E update (space.py:2073-2085, if / elif)
            if self.bbc['x']:
                self._ghost_x((self.Hx, self.Hy, self.Hz), self.Lx - 2 * self.dx)
            elif self.pbc['x']:
                self._ghost_x((self.Hx, self.Hy, self.Hz), 0)
            s2 = slice(1, None) if self.rank == 0 else slice(None)      # see _bloch_h
            if self.bbc['y']:
                hz = self._spec(self.Hz, self.ymshift, 1)
                hx = self._spec(self.Hx, self.ymshift, 1)
                self.Ex[:] += dt / eps * 1j * (-mm[1] * hz)
                self.Ez[s2] += dt / eps[s2] * 1j * (+mm[1] * hx[s2])
            if self.bbc['z']:
                hy = self._spec(self.Hy, self.zmshift, 2)
                hx = self._spec(self.Hx, self.zmshift, 2)
                self.Ex[:] += dt / eps * 1j * (+mm[2] * hy)
                self.Ey[s2] += dt / eps[s2] * 1j * (-mm[2] * hx[s2])
        else:
            if self.bbc['x']:
                self.Ey = self.Ey + dt / eps * 1j * (+mm[0] * self.Hz)
                self.Ez = self.Ez + dt / eps * 1j * (-mm[0] * self.Hy)
            if self.bbc['y']:
                self.Ex = self.Ex + dt / eps * 1j * (-mm[1] * self.Hz)
                self.Ez = self.Ez + dt / eps * 1j * (+mm[1] * self.Hx)
            if self.bbc['z']:
                self.Ex = self.Ex + dt / eps * 1j * (+mm[2] * self.Hy)
                self.Ey = self.Ey + dt / eps * 1j * (-mm[2] * self.Hx)

    # ------------------------------------------------------------------- CPML
    # Table of the reference's slice conventions, one row per (face, half,
    # method-family).  Each row: profile slice, then per component
    # (target, diff, sign, field slices (x,y,z), psi slices (x,y,z)).
    # 'L' = "exclude last x plane on the last rank", 'F' = "exclude first x
    # plane on rank 0" (space.py:1315-1320, 1367-1372 and analogues).
    def _face_rows(self, half, face):
        P = self.npml
        m = self.method
        S = slice
        A = S(None)
        ax = face[0]
        plus = face[1] == 'p'
        last = self.rank == self.size - 1
        first = self.rank == 0
        xL = A if not last else S(0, -1)       # H components updated on [:-1] in x
        xF = A if not first else S(1, None)    # E components updated on [1:] in x
        spectral = m in ('SHPF', 'PSTD')
        rows = []
        if half == 'H':
            if ax == 'x':
                # space.py:1110-1162 (x+), 1198-1250 (x-)
                if plus:
                    if m == 'PSTD':
                        prof, f, p = S(0, None, 2), S(-P, None), S(0, None)
                    else:
                        prof, f, p = S(1, -1, 2), S(-P, -1), S(0, -1)
                else:
                    prof = S(-1, None, -2) if m == 'PSTD' else S(-2, None, -2)
                    f, p = S(0, P), S(0, P)
                zy = A if spectral else S(0, -1)
                rows.append(('Hy', 'xEz', -1, (f, A, zy), (p, A, zy), 'hyx'))
                rows.append(('Hz', 'xEy', +1, (f, zy, A), (p, zy, A), 'hzx'))
            elif ax == 'y':
                # space.py:1296-1346 (y+), 1400-1452 (y-)
                if plus:
                    if m == 'PSTD':
                        prof, f, p = S(0, None, 2), S(-P, None), S(0, None)
                    elif m == 'SHPF':
                        prof, f, p = S(1, None, 2), S(-P, None), S(0, None)
                    else:
                        prof, f, p = S(1, -1, 2), S(-P, -1), S(0, -1)
                else:
                    prof = S(-1, None, -2) if m == 'PSTD' else S(-2, None, -2)
                    f, p = S(0, P), S(0, P)
                xz = A if m == 'PSTD' else xL
                zx = A if spectral else S(0, -1)
                rows.append(('Hx', 'yEz', +1, (A, f, zx), (A, p, zx), 'hxy'))
                rows.append(('Hz', 'yEx', -1, (xz, f, A), (xz, p, A), 'hzy'))
            else:
                # space.py:1506-1556 (z+), 1610-1660 (z-)
                if plus:
                    if m == 'PSTD':
                        prof, f, p = S(0, None, 2), S(-P, None), S(0, P)
                    elif m == 'SHPF':
                        prof, f, p = S(1, None, 2), S(-P, None), S(0, P)
                    else:
                        prof, f, p = S(1, -1, 2), S(-P, -1), S(0, P - 1)
                else:
                    prof = S(-2, None, -2)
                    f, p = S(0, P), S(0, P)
                xy = A if m == 'PSTD' else xL
                yx = A if spectral else S(0, -1)
                rows.append(('Hx', 'zEy', -1, (A, yx, f), (A, yx, p), 'hxz'))
                if m == 'FDTD' and not plus and last:
                    # space.py:1647-1648: the last rank also trims y on Hy at z-
                    rows.append(('Hy', 'zEx', +1, (xy, S(0, -1), f), (xy, S(0, -1), p), 'hyz'))
                else:
                    rows.append(('Hy', 'zEx', +1, (xy, A, f), (xy, A, p), 'hyz'))
        else:
            if ax == 'x':
                # space.py:1164-1196 (x+), 1252-1294 (x-)
                if plus:
                    prof, f, p = S(0, None, 2), S(-P, None), S(0, None)
                else:
                    if m == 'PSTD':
                        prof, f, p = S(-1, None, -2), S(0, P), S(0, P)
                    else:
                        prof, f, p = S(-3, None, -2), S(1, P), S(1, P)
                zy = A if spectral else S(1, None)
                rows.append(('Ey', 'xHz', -1, (f, A, zy), (p, A, zy), 'eyx'))
                rows.append(('Ez', 'xHy', +1, (f, zy, A), (p, zy, A), 'ezx'))
            elif ax == 'y':
                # space.py:1348-1398 (y+), 1454-1504 (y-)
                if plus:
                    prof, f, p = S(0, None, 2), S(-P, None), S(0, P)
                else:
                    if spectral:
                        prof, f, p = S(-1, None, -2), S(0, P), S(0, P)
                    else:
                        prof, f, p = S(-3, None, -2), S(1, P), S(1, P)
                xz = A if m == 'PSTD' else xF
                zx = A if spectral else S(1, None)
                rows.append(('Ex', 'yHz', +1, (A, f, zx), (A, p, zx), 'exy'))
                rows.append(('Ez', 'yHx', -1, (xz, f, A), (xz, p, A), 'ezy'))
            else:
                # space.py:1558-1608 (z+), 1662-1712 (z-)
                if plus:
                    prof, f, p = S(0, None, 2), S(-P, None), S(0, P)
                else:
                    if m == 'PSTD':
                        prof, f, p = S(-2, None, -2), S(0, P), S(0, P)
                    elif m == 'SHPF':
                        prof, f, p = S(-1, None, -2), S(0, P), S(0, P)
                    else:
                        prof, f, p = S(-3, None, -2), S(1, P), S(1, P)
                xy = A if m == 'PSTD' else xF
                yx = A if spectral else S(1, None)
                rows.append(('Ex', 'zHy', -1, (A, yx, f), (A, yx, p), 'exz'))
                rows.append(('Ey', 'zHx', +1, (xy, A, f), (xy, A, p), 'eyz'))
        return prof, rows

    def _pml_face(self, half, face):
        ax = face[0]
        side = face[1]
        prof, rows = self._face_rows(half, face)
        pr = self.pml[ax]
        bshape = {'x': (slice(None), None, None), 'y': (None, slice(None), None),
                  'z': (None, None, slice(None))}[ax]
        b = pr['b'][prof][bshape]
        a = pr['a'][prof][bshape]
        kap = pr['kappa'][prof][bshape]
        C = self.CH2 if half == 'H' else self.CE2
        for tgt, dn, sgn, fs, ps, psn in rows:
            F = getattr(self, tgt)
            D = self.d[dn]
            psi = self.psi[f'{psn}_{side}']
            C2 = C[fs]
            psi[ps] = (b * psi[ps]) + (a * D[fs])
            if sgn > 0:
                F[fs] += C2 * (+((1. / kap - 1.) * D[fs]) + psi[ps])
            else:
                F[fs] += C2 * (-((1. / kap - 1.) * D[fs]) - psi[ps])

    def _pml_faces(self):
        """Dispatch order of space.py:1054-1108."""
        out = []
        r = self.PMLregion
        for ax in 'yz':
            if ax in r:
                if '+' in r[ax]: out.append(ax + 'p')
                if '-' in r[ax]: out.append(ax + 'm')
        if 'x' in r:
            if self.rank == 0:
                if '+' in r['x'] and self.size == 1: out.append('xp')
                if '-' in r['x']: out.append('xm')
            elif self.rank == self.size - 1:
                if '+' in r['x']: out.append('xp')
        return out

    def _pml_h(self):
        for face in self._pml_faces():
            self._pml_face('H', face)

    def _pml_e(self):
        for face in self._pml_faces():
            self._pml_face('E', face)


# --------------------------------------------------------------------- sources
def _pyround(x):
    return round(x)


class OracleSetter:
    """source.Setter (source.py:8-253): index logic and soft/hard injection."""

    def __init__(self, space, src_srt, src_end, mmt):
        self.space = space
        sp = space
        self.src_xsrt = _pyround(src_srt[0] / sp.dx)
        self.src_xend = _pyround(src_end[0] / sp.dx)
        self.src_ysrt = _pyround(src_srt[1] / sp.dy)
        self.src_yend = _pyround(src_end[1] / sp.dy)
        self.src_zsrt = _pyround(src_srt[2] / sp.dz)
        self.src_zend = _pyround(src_end[2] / sp.dz)
        self.who_put_src = None
        for rank in range(sp.size):
            my_xsrt, my_xend = sp.myNx_indice[rank]
            if self.src_xsrt == self.src_xend:
                self.src_xsrt = self.src_xend - 1      # source.py:88
            if self.src_xsrt == self.src_xend - 1:
                if self.src_xsrt >= my_xsrt and self.src_xend <= my_xend:
                    self.who_put_src = rank
                    if sp.rank == rank:
                        self.my_src_xsrt = self.src_xsrt - my_xsrt
                        self.my_src_xend = self.src_xend - my_xsrt
            elif self.src_xsrt < self.src_xend:
                assert sp.size == 1
                self.who_put_src = 0
                self.my_src_xsrt = self.src_xsrt
                self.my_src_xend = self.src_xend
            else:
                raise ValueError("src_end[0] should be bigger than src_srt[0]")
        sp.mmt = mmt
        if sp.rank == self.who_put_src:
            kx, ky, kz = mmt
            self.px = np.exp(+1j * kx * np.arange(self.my_src_xsrt, self.my_src_xend) * sp.dx)
            self.py = np.exp(+1j * ky * np.arange(self.src_ysrt, self.src_yend) * sp.dy)
            self.pz = np.exp(+1j * kz * np.arange(self.src_zsrt, self.src_zend) * sp.dz)
            if self.my_src_xend - self.my_src_xsrt == 1:
                self.px = np.exp(1j * kx * np.arange(1) * sp.dx)
            if self.src_yend - self.src_ysrt == 1:
                self.py = np.exp(1j * ky * np.arange(1) * sp.dy)
            if self.src_zend - self.src_zsrt == 1:
                self.pz = np.exp(1j * kz * np.arange(1) * sp.dz)

    def put_src(self, where, pulse, put_type):
        sp = self.space
        if sp.rank != self.who_put_src:
            return
        x = slice(self.my_src_xsrt, self.my_src_xend)
        y = slice(self.src_ysrt, self.src_yend)
        z = slice(self.src_zsrt, self.src_zend)
        if sp.BBC_called:
            pulse = pulse * (self.px[:, None, None] * self.py[None, :, None] * self.pz[None, None, :])
        name = where[0].upper() + where[1].lower()
        F = getattr(sp, name)
        if put_type == 'soft':
            F[x, y, z] += pulse
        elif put_type == 'hard':
            F[x, y, z] = pulse
        else:
            raise ValueError("Please insert 'soft' or 'hard'")


def gaussian_pulse_c(step, dt, center_wv, spread, pick_pos):
    """source.Gaussian.pulse_c (source.py:256-283)."""
    w0 = 2 * np.pi * (c / center_wv)
    ws = spread * w0
    tc = pick_pos * dt
    return np.exp((-.5) * (((step * dt - tc) * ws) ** 2)) * np.exp(-1j * w0 * (step * dt - tc))


def gaussian_pulse_re(step, dt, center_wv, spread, pick_pos):
    """source.Gaussian.pulse_re (source.py:285-290)."""
    w0 = 2 * np.pi * (c / center_wv)
    ws = spread * w0
    tc = pick_pos * dt
    return np.exp((-.5) * (((step * dt - tc) * ws) ** 2)) * np.cos(w0 * (step * dt - tc))


# ------------------------------------------------------------------ collectors
def local_x_loc(space, gxsrts, gxends):
    """collector._get_local_x_loc (collector.py:30-118)."""
    assert gxsrts >= 0
    assert gxends < space.Nx
    bxsrt, bxend = space.myNx_indice[space.rank]
    gxloc = lxloc = None
    if gxsrts >= bxsrt and gxsrts < bxend and gxends <= bxend:
        gxloc, lxloc = (gxsrts, gxends), (gxsrts - bxsrt, gxends - bxsrt)
    if gxsrts >= bxsrt and gxsrts < bxend and gxends > bxend:
        gxloc, lxloc = (gxsrts, bxend), (gxsrts - bxsrt, bxend - bxsrt)
    if gxsrts < bxsrt and gxends > bxend:
        gxloc, lxloc = (bxsrt, bxend), (0, bxend - bxsrt)
    if gxsrts < bxsrt and gxends > bxsrt and gxends <= bxend:
        gxloc, lxloc = (bxsrt, gxends), (0, gxends - bxsrt)
    return gxloc, lxloc


class OracleSx:
    """collector.Sx (collector.py:266-382): running DFT of Ey,Ez,Hy,Hz on a yz plane."""

    def __init__(self, space_fields, space, xloc, srt, end, freqs):
        self.fields = space_fields        # callable name -> array (allows SF = TF - IF)
        self.space = space
        self.freqs = np.asarray(freqs)
        self.xsrt = _pyround(xloc / space.dx)
        self.ysrt = _pyround(srt[0] / space.dy)
        self.zsrt = _pyround(srt[1] / space.dz)
        self.xend = self.xsrt + 1
        self.yend = _pyround(end[0] / space.dy)
        self.zend = _pyround(end[1] / space.dz)
        self.gxloc, self.lxloc = local_x_loc(space, self.xsrt, self.xend)
        if self.gxloc is not None:
            shp = (len(self.freqs), self.yend - self.ysrt, self.zend - self.zsrt)
            self.DFT = {n: np.zeros(shp, dtype=np.complex128) for n in ('Ey', 'Ez', 'Hy', 'Hz')}

    def do_RFT(self, tstep):
        if self.gxloc is None:
            return
        dt = self.space.dt
        idx = (slice(self.lxloc[0], self.lxloc[1]), slice(self.ysrt, self.yend),
               slice(self.zsrt, self.zend))
        f = (slice(None), None, None)
        for n in ('Ey', 'Hz', 'Ez', 'Hy'):
            self.DFT[n] += self.fields(n)[idx] * np.exp(2.j * np.pi * self.freqs[f] * tstep * dt) * dt

    def get_Sx(self):
        D = self.DFT
        Sx = 0.5 * ((D['Ey'].real * D['Hz'].real) + (D['Ey'].imag * D['Hz'].imag)
                    - (D['Ez'].real * D['Hy'].real) - (D['Ez'].imag * D['Hy'].imag))
        return Sx.sum(axis=(1, 2)) * self.space.dy * self.space.dz


# ------------------------------------------------------------- multi-rank glue
class OracleCluster:
    """R x-slabs stepped in-process with explicit halo hand-over; stands in for
    `mpirun -n R` of the reference (space.py:645-670, 863-887)."""

    def __init__(self, size, *args, **kw):
        self.slabs = [OracleSpace(*args, rank=r, size=size, **kw) for r in range(size)]

    def each(self, fn):
        return [fn(s) for s in self.slabs]

    def update_h(self, t):
        S = self.slabs
        halos = [(S[r + 1].Ey[0].copy(), S[r + 1].Ez[0].copy()) if r + 1 < len(S) else None
                 for r in range(len(S))]
        for s, h in zip(S, halos):
            s.update_h(t, h)

    def update_e(self, t):
        S = self.slabs
        halos = [(S[r - 1].Hy[-1].copy(), S[r - 1].Hz[-1].copy()) if r > 0 else None
                 for r in range(len(S))]
        for s, h in zip(S, halos):
            s.update_e(t, h)

    def gather(self, name):
        return np.concatenate([getattr(s, name) for s in self.slabs], axis=0)
