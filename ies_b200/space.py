"""space.Basic3D / space.Empty3D on the B200 engine (engine='b200').

Mirrors the constructor/method surface of the reference's space.py
(Basic3D: space.py:7-2148, Empty3D: 2151-2179) so the tutorial / example scripts
run with only the engine string switched.  Host-side set-up math (k vectors,
half-cell shift tables, CPML profiles, update coefficients, slab indices) is
NumPy and follows the cited lines; the six fields, the CPML psi arrays and all
per-step work live on the GPU behind the C-ABI of include/ies_b200.h.  There is
no CPU fallback: without libies_b200.so and a CUDA device construction fails.
"""
import ctypes as C
import os

import numpy as np
from scipy.constants import c, mu_0, epsilon_0

try:
    from . import _lib, comm as _comm
except ImportError:  # sys.path drop-in: `import space`
    import _lib
    import comm as _comm

_FIELDS = ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')


def _box_from_index(idx, shape):
    """(lo, hi, squeeze_axes) for an index made of ints and unit-step slices, else None."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    if any(i is Ellipsis for i in idx):
        n = len(shape) - (len(idx) - 1)
        k = idx.index(Ellipsis)
        idx = idx[:k] + (slice(None),) * n + idx[k + 1:]
    if len(idx) > len(shape):
        raise IndexError("too many indices")
    idx = idx + (slice(None),) * (len(shape) - len(idx))
    lo, hi, sq = [], [], []
    for a, (i, n) in enumerate(zip(idx, shape)):
        if isinstance(i, (int, np.integer)):
            i = int(i)
            if i < 0:
                i += n
            if not 0 <= i < n:
                raise IndexError(f"index {i} is out of bounds for axis {a} with size {n}")
            lo.append(i); hi.append(i + 1); sq.append(a)
        elif isinstance(i, slice):
            s, e, st = i.indices(n)
            if st != 1:
                return None
            lo.append(s); hi.append(max(s, e))
        else:
            return None
    return lo, hi, tuple(sq)


class DeviceField:
    """Handle to one device-resident field component.  Indexing copies the addressed
    box to / from the host, so reference-style code (`space.Ey[x,y,z] += pulse`,
    `space.Ey[Fidx]`, `np.asarray(space.Ex)`) keeps working; the engine's own
    Setter / collectors never go through it."""

    def __init__(self, space, name):
        self._space, self._name = space, name
        self._comp = _lib.COMP[name]

    shape = property(lambda self: self._space.loc_grid)
    dtype = property(lambda self: np.dtype(self._space.field_dtype))
    ndim = 3
    size = property(lambda self: int(np.prod(self._space.loc_grid)))
    real = property(lambda self: np.asarray(self).real)
    imag = property(lambda self: np.asarray(self).imag)

    def _get_box(self, lo, hi):
        out = np.empty([h - l for l, h in zip(lo, hi)], dtype=self.dtype)
        if out.size:
            lib = _lib.load()
            _lib.check(lib.ies_get_field(self._space._ctx, self._comp, _lib.I3(*lo), _lib.I3(*hi),
                                         out.ctypes.data_as(C.c_void_p)))
        return out

    def _set_box(self, lo, hi, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        if arr.size:
            lib = _lib.load()
            _lib.check(lib.ies_set_field(self._space._ctx, self._comp, _lib.I3(*lo), _lib.I3(*hi),
                                         arr.ctypes.data_as(C.c_void_p)))

    def __array__(self, dtype=None, copy=None):
        a = self._get_box((0, 0, 0), self.shape)
        return a if dtype is None else a.astype(dtype)

    def get(self):
        return np.asarray(self)

    def copy(self):
        return np.asarray(self)

    def __getitem__(self, idx):
        b = _box_from_index(idx, self.shape)
        if b is None:
            return np.asarray(self)[idx]
        lo, hi, sq = b
        out = self._get_box(lo, hi)
        return out.reshape([n for a, n in enumerate(out.shape) if a not in sq]) if sq else out

    def __setitem__(self, idx, val):
        if isinstance(val, DeviceField):
            val = np.asarray(val)
        b = _box_from_index(idx, self.shape)
        if b is None:
            full = np.asarray(self)
            full[idx] = val
            self._set_box((0, 0, 0), self.shape, full)
            return
        lo, hi, sq = b
        shp = [h - l for l, h in zip(lo, hi)]
        if (isinstance(val, np.ndarray) and val.dtype == self.dtype and list(val.shape) == shp
                and val.flags.c_contiguous):
            self._set_box(lo, hi, val)      # no host staging copy
            return
        tmp = np.empty([n for a, n in enumerate(shp) if a not in sq], dtype=self.dtype)
        tmp[...] = val          # NumPy's own casting / broadcasting rules and errors
        self._set_box(lo, hi, tmp.reshape(shp))

    # array arithmetic on the host copy (Empty3D.get_SF style expressions)
    def __sub__(self, o): return np.asarray(self) - np.asarray(o)
    def __rsub__(self, o): return np.asarray(o) - np.asarray(self)
    def __add__(self, o): return np.asarray(self) + np.asarray(o)
    __radd__ = __add__
    def __mul__(self, o): return np.asarray(self) * np.asarray(o)
    __rmul__ = __mul__
    def __truediv__(self, o): return np.asarray(self) / np.asarray(o)
    def __neg__(self): return -np.asarray(self)
    def __abs__(self): return np.abs(np.asarray(self))
    def __len__(self): return self.shape[0]
    def __repr__(self): return f"DeviceField({self._name}, shape={self.shape}, dtype={self.dtype})"


def _field_property(name):
    def get(self):
        return self._field_handles[name]

    def set_(self, value):
        # reference code rebinds fields (`self.Hx = ...`, space.py:795-797, 2173-2179)
        self._field_handles[name][:, :, :] = np.asarray(value)
    return property(get, set_)


class Basic3D:

    def __init__(self, grid, gridgap, dt, tsteps, field_dtype, mmtdtype, **kwargs):
        """Create the simulation space (space.py:9-141).

        Extra kwargs of this engine: `comm` (an ies_b200.comm communicator; default
        TorchComm under torchrun, else a single slab) and `device` (CUDA ordinal)."""
        self.nm = 1e-9
        self.um = 1e-6

        self.field_dtype = np.dtype(field_dtype).type
        self.mmtdtype = np.dtype(mmtdtype).type
        if np.dtype(self.field_dtype) not in _lib.DTYPE_CODE:
            raise ValueError("Please use field_dtype for numpy dtype!")      # space.py:162

        comm = kwargs.get('comm')
        self.MPIcomm = comm if comm is not None else _comm.default_comm()
        self.MPIrank = self.MPIcomm.Get_rank()
        self.MPIsize = self.MPIcomm.Get_size()
        self.hostname = os.uname().nodename

        assert len(grid) == 3, "Simulation grid should be a tuple with length 3."
        assert len(gridgap) == 3, "Argument 'gridgap' should be a tuple with length 3."

        self.tsteps = tsteps
        self.grid = grid
        self.Nx, self.Ny, self.Nz = self.grid
        self.TOTAL_NUM_GRID = self.Nx * self.Ny * self.Nz
        self.TOTAL_NUM_GRID_SIZE = (self.field_dtype(1).nbytes * self.TOTAL_NUM_GRID) / 1024 / 1024
        self.dimension = 3

        self.Nxc = round(self.Nx / 2)
        self.Nyc = round(self.Ny / 2)
        self.Nzc = round(self.Nz / 2)

        self.gridgap = gridgap
        self.dx, self.dy, self.dz = self.gridgap

        self.Lx = (self.Nx - 1) * self.dx
        self.Ly = (self.Ny - 1) * self.dy
        self.Lz = (self.Nz - 1) * self.dz
        self.VOLUME = self.Lx * self.Ly * self.Lz

        self.method = 'SHPF'
        self.engine = 'b200'
        self.courant = 1. / 4
        self.BBC_called = False
        self.PBC_called = False

        if kwargs.get('engine') is not None: self.engine = kwargs.get('engine')
        if kwargs.get('method') is not None: self.method = kwargs.get('method')
        if kwargs.get('courant') is not None: self.courant = kwargs.get('courant')

        if self.method == 'PSTD':
            assert self.MPIsize == 1, "MPI size must be 1 if you want to use the PSTD method."
        # the reference asserts engine in ('numpy','cupy') (space.py:91); this package IS the
        # third engine and accepts those two strings as aliases so scripts need no other edit.
        assert self.engine in ('b200', 'cupy', 'numpy')
        if self.method not in _lib.METHOD_CODE:
            raise NotImplementedError(
                f"method {self.method!r}: the b200 engine implements FDTD, SHPF and PSTD "
                "(HPF/SPSTD are marked 'in developing' in the reference README)")
        self.xp = np            # host-side array module for set-up math

        self.dt = dt
        self.maxdt = 1. / c / np.sqrt((1. / self.dx) ** 2 + (1. / self.dy) ** 2 + (1. / self.dz) ** 2)
        assert (c * self.dt * np.sqrt((1. / self.dx) ** 2 + (1. / self.dy) ** 2 + (1. / self.dz) ** 2)) < 1.
        assert self.dt < self.maxdt, "Time interval is too big so that causality is broken. Lower the courant number."
        assert float(self.Nx) % self.MPIsize == 0., "Nx must be a multiple of the number of nodes."

        self.myNx = round(self.Nx / self.MPIsize)
        self.loc_grid = (self.myNx, self.Ny, self.Nz)

        self.myNx_slices = []
        self.myNx_indice = []
        for rank in range(self.MPIsize):
            xsrt = (rank) * self.myNx
            xend = (rank + 1) * self.myNx
            self.myNx_slices.append(slice(xsrt, xend))
            self.myNx_indice.append((xsrt, xend))

        # ---- device context: six zero fields (space.py:117-122) + scratch
        self.device = kwargs.get('device')
        if self.device is None:
            self.device = _comm.default_device(self.MPIcomm)
        self._lib = _lib.load()
        self._ctx_obj = None
        if not getattr(self, '_lazy_ctx', False):
            self._create_ctx()
        self._field_handles = {n: DeviceField(self, n) for n in _FIELDS}
        self._dirty = True
        self._malloc_called = False
        self._coeff_ready = False
        self.PMLregion = {}
        self.npml = 0
        self.apply_BBCx = self.apply_BBCy = self.apply_BBCz = False
        self.apply_PBCx = self.apply_PBCy = self.apply_PBCz = False
        self.mmt = None
        if isinstance(self.MPIcomm, _comm.LocalComm):
            self.MPIcomm.register(self, kwargs.get('group_key', 'space'))
        self.MPIcomm.Barrier()

    def _create_ctx(self):
        cfg = _lib.Config(self.myNx, self.Ny, self.Nz, _lib.DTYPE_CODE[np.dtype(self.field_dtype)],
                          _lib.METHOD_CODE[self.method], self.MPIrank, self.MPIsize, int(self.device),
                          self.dx, self.dy, self.dz, self.dt)
        ctx = C.c_void_p()
        _lib.check(self._lib.ies_create(C.byref(cfg), C.byref(ctx)))
        self._ctx_obj = ctx

    @property
    def _ctx(self):
        """The engine context; an Empty3D creates its own only if somebody asks for its fields before
        get_SF made it a (TF, IF) view (a full six-field slab otherwise sits unused: 3 GiB at the
        headline grid)."""
        if self._ctx_obj is None:
            self._create_ctx()
        return self._ctx_obj

    Ex = _field_property('Ex'); Ey = _field_property('Ey'); Ez = _field_property('Ez')
    Hx = _field_property('Hx'); Hy = _field_property('Hy'); Hz = _field_property('Hz')

    def __del__(self):
        try:
            if getattr(self, '_ctx_obj', None) is not None and self._ctx_obj.value:
                self._lib.ies_destroy(self._ctx_obj)
                self._ctx_obj = None
        except Exception:
            pass

    def _use_stream(self, stream):
        """Run on an external CUDA stream handle (int; 0 = the caller's default stream), or on the
        context's own stream again with stream=None."""
        if getattr(self, '_stream', 'own') != stream:
            _lib.check(self._lib.ies_set_stream(self._ctx, C.c_void_p(stream or 0), int(stream is None)))
            self._stream = stream

    def sync(self):
        _lib.check(self._lib.ies_sync(self._ctx))

    # ------------------------------------------------------------------ malloc
    def malloc(self):
        """FFT wavenumber / half-cell shift tables and host material arrays
        (space.py:143-237; the engine=='cupy' branch 168-181)."""
        cplx = np.dtype(self.field_dtype).kind == 'c'
        if cplx:
            self.fftfreq = np.fft.fftfreq
            if self.method in ('PSTD', 'SHPF'): print("Complex FFT kernel assigned.")
        else:
            self.fftfreq = np.fft.rfftfreq
            if self.method in ('PSTD', 'SHPF'): print("Real FFT kernel assigned.")

        self.kx = self.fftfreq(self.Nx, self.dx) * 2 * np.pi
        self.ky = self.fftfreq(self.Ny, self.dy) * 2 * np.pi
        self.kz = self.fftfreq(self.Nz, self.dz) * 2 * np.pi

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

        # one shared eps and one shared mu array, host side until
        # init_update_constants() (space.py:207-218); structures write into them.
        self.eps = np.ones(self.loc_grid, dtype=np.float64) * epsilon_0
        self.eps_Ex = self.eps_Ey = self.eps_Ez = self.eps
        self.mu = np.ones(self.loc_grid, dtype=np.float64) * mu_0
        self.mu_Hx = self.mu_Hy = self.mu_Hz = self.mu
        self._malloc_called = True
        self._dirty = True

    _COND = ('econ_Ex', 'econ_Ey', 'econ_Ez', 'mcon_Hx', 'mcon_Hy', 'mcon_Hz')

    def __getattr__(self, name):
        # the reference's conductivity arrays (space.py:223-234) are identically zero in every shipped
        # script and the engine drops the C1 coefficient on that ground: they exist only when a
        # script touches them, and init_update_constants() refuses non-zero values
        if name in Basic3D._COND and self.__dict__.get('_malloc_called'):
            arr = np.zeros(self.loc_grid, dtype=np.float64)
            self.__dict__[name] = arr
            return arr
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    # --------------------------------------------------------------- apply_PML
    def apply_PML(self, region, npml):
        """CPML grading and the b/a recursion coefficients (space.py:239-361)."""
        self.PMLregion = region
        self.npml = npml
        self.PMLgrading = 2 * self.npml

        self.rc0 = 1.e-16
        self.imp = np.sqrt(mu_0 / epsilon_0)
        self.gO = 3.
        self.sO = 3.
        self.bdw_x = (self.PMLgrading - 1) * (self.dx / 2)
        self.bdw_y = (self.PMLgrading - 1) * (self.dy / 2)
        self.bdw_z = (self.PMLgrading - 1) * (self.dz / 2)

        self.PMLsigmamaxx = -(self.gO + 1) * np.log(self.rc0) / (2 * self.imp * self.bdw_x)
        self.PMLsigmamaxy = -(self.gO + 1) * np.log(self.rc0) / (2 * self.imp * self.bdw_y)
        self.PMLsigmamaxz = -(self.gO + 1) * np.log(self.rc0) / (2 * self.imp * self.bdw_z)

        self.PMLkappamaxx = self.PMLkappamaxy = self.PMLkappamaxz = 7.
        self.PMLalphamaxx = self.PMLalphamaxy = self.PMLalphamaxz = 0.05

        fd = self.field_dtype
        for ax in 'xyz':
            setattr(self, 'PMLsigma' + ax, np.zeros(self.PMLgrading, dtype=fd))
            setattr(self, 'PMLalpha' + ax, np.zeros(self.PMLgrading, dtype=fd))
            setattr(self, 'PMLkappa' + ax, np.ones(self.PMLgrading, dtype=fd))
            setattr(self, 'PMLb' + ax, np.zeros(self.PMLgrading, dtype=fd))
            setattr(self, 'PMLa' + ax, np.zeros(self.PMLgrading, dtype=fd))

        for key, value in self.PMLregion.items():
            if key in 'xyz' and value != '':
                smax = getattr(self, 'PMLsigmamax' + key)
                kmax = getattr(self, 'PMLkappamax' + key)
                amax = getattr(self, 'PMLalphamax' + key)
                loc = np.arange(self.PMLgrading) / (self.PMLgrading - 1)
                sigma = smax * (loc ** self.gO)
                kappa = 1 + ((kmax - 1) * (loc ** self.gO))
                alpha = amax * ((1 - loc) ** self.sO)
                b = np.exp(-(sigma / kappa + alpha) * self.dt / epsilon_0)
                a = sigma / (sigma * kappa + alpha * kappa ** 2) * (b - 1.)
                setattr(self, 'PMLsigma' + key, sigma)
                setattr(self, 'PMLkappa' + key, kappa)
                setattr(self, 'PMLalpha' + key, alpha)
                setattr(self, 'PMLb' + key, b)
                setattr(self, 'PMLa' + key, a)
        self._dirty = True
        return

    def save_pml_parameters(self, path):
        """space.py:363-398.  h5py is optional; without it a .npz with the same
        dataset names is written."""
        if self.MPIrank == 0:
            data = {}
            for key in self.PMLregion.keys():
                for nm in ('PMLsigma', 'PMLkappa', 'PMLalpha', 'PMLb', 'PMLa'):
                    data[nm + key] = getattr(self, nm + key)
            try:
                import h5py
                with h5py.File(path + 'pml_parameters.h5', 'w') as f:
                    for k, v in data.items():
                        f.create_dataset(k, data=v)
            except ImportError:
                np.savez(path + 'pml_parameters.npz', **data)
        self.MPIcomm.Barrier()
        return

    def save_eps_mu(self, path):
        """space.py:400-443: relative eps/mu per rank, dataset names eps_Ex..mu_Hz."""
        save_dir = path + 'eps_mu/'
        if not os.path.exists(save_dir): os.makedirs(save_dir, exist_ok=True)
        data = {'eps_Ex': self.eps_Ex / epsilon_0, 'eps_Ey': self.eps_Ey / epsilon_0,
                'eps_Ez': self.eps_Ez / epsilon_0, 'mu_Hx': self.mu_Hx / mu_0,
                'mu_Hy': self.mu_Hy / mu_0, 'mu_Hz': self.mu_Hz / mu_0}
        base = save_dir + 'eps_r_mu_r_rank{:>02d}'.format(self.MPIrank)
        try:
            import h5py
            with h5py.File(base + '.h5', 'w') as f:
                for k, v in data.items():
                    f.create_dataset(k, data=v)
        except ImportError:
            np.savez(base + '.npz', **data)
        self.MPIcomm.Barrier()
        return

    def load_eps_mu(self, path):
        """Inverse of save_eps_mu (materials interchange): reads eps_r/mu_r of this rank."""
        base = path + 'eps_mu/eps_r_mu_r_rank{:>02d}'.format(self.MPIrank)
        if os.path.exists(base + '.h5'):
            import h5py
            with h5py.File(base + '.h5', 'r') as f:
                er, mr = f['eps_Ex'][...], f['mu_Hx'][...]
        else:
            z = np.load(base + '.npz')
            er, mr = z['eps_Ex'], z['mu_Hx']
        self.eps[...] = er * epsilon_0
        self.mu[...] = mr * mu_0
        self._dirty = True

    # ---------------------------------------------------- init_update_constants
    def init_update_constants(self):
        """Freeze the materials into update coefficients (space.py:445-553).  The
        reference's conductivity arrays are identically zero (space.py:223-234), so
        C1 == 1 exactly and one f64 array per half-step is uploaded:
        CH2 = -2dt/(2mu), CE2 = 2dt/(2eps), evaluated with the reference's expression."""
        assert self._malloc_called, "call malloc() first"
        for nm in Basic3D._COND:
            if nm in self.__dict__ and np.any(self.__dict__[nm]):
                raise NotImplementedError(f"{nm} != 0: the b200 engine implements the lossless update "
                                          "(C1 == 1); conductive media are not supported")
        self.CHx2 = self.CHy2 = self.CHz2 = (-2 * self.dt) / (2. * self.mu)
        self.CEx2 = self.CEy2 = self.CEz2 = (2. * self.dt) / (2. * self.eps)
        n = self.CHx2.size
        _lib.check(self._lib.ies_set_coeff(self._ctx, _lib.HALF_H, _lib.dptr(np.ascontiguousarray(self.CHx2)), n))
        _lib.check(self._lib.ies_set_coeff(self._ctx, _lib.HALF_E, _lib.dptr(np.ascontiguousarray(self.CEx2)), n))
        self._coeff_ready = True
        self._dirty = True

    # ------------------------------------------------------------- BBC / PBC
    def apply_BBC(self, region):
        """space.py:555-615."""
        self.apply_BBCx = region.get('x')
        self.apply_BBCy = region.get('y')
        self.apply_BBCz = region.get('z')
        if True in region.values():
            self.BBC_called = True
            assert self.field_dtype != np.float32
            assert self.field_dtype != np.float64
        if self.apply_BBCx == True:
            assert self.MPIsize == 1
        self._dirty = True
        return

    def apply_PBC(self, region):
        """space.py:617-637."""
        self.PBC_called = True
        self.apply_PBCx = region.get('x')
        self.apply_PBCy = region.get('y')
        self.apply_PBCz = region.get('z')
        if self.apply_PBCx == True: assert self.MPIsize == 1
        self._dirty = True
        return

    # ---------------------------------------------------------- engine set-up
    def _main_boxes(self):
        """Sub-volume of each component's main update (space.py:801-825, 1017-1037)."""
        nx, ny, nz = self.loc_grid
        first, last = self.MPIrank == 0, self.MPIrank == self.MPIsize - 1
        xH = (0, nx - 1) if last else (0, nx)       # Hy, Hz on [:-1] unless a halo plane arrives
        xE = (1, nx) if first else (0, nx)          # Ey, Ez on [1:]  unless a halo plane arrives
        A = lambda n: (0, n)
        if self.method == 'PSTD':
            return {n: (A(nx), A(ny), A(nz)) for n in _FIELDS}
        if self.method == 'SHPF':
            return {'Hx': (A(nx), A(ny), A(nz)), 'Hy': (xH, A(ny), A(nz)), 'Hz': (xH, A(ny), A(nz)),
                    'Ex': (A(nx), A(ny), A(nz)), 'Ey': (xE, A(ny), A(nz)), 'Ez': (xE, A(ny), A(nz))}
        return {'Hx': (A(nx), (0, ny - 1), (0, nz - 1)), 'Hy': (xH, A(ny), (0, nz - 1)),
                'Hz': (xH, (0, ny - 1), A(nz)),
                'Ex': (A(nx), (1, ny), (1, nz)), 'Ey': (xE, A(ny), (1, nz)), 'Ez': (xE, (1, ny), A(nz))}

    def _pml_faces(self):
        """Face order of _updateH_PML/_updateE_PML (space.py:1054-1108)."""
        out = []
        r = self.PMLregion
        for ax in 'yz':
            if ax in r:
                if '+' in r.get(ax): out.append((ax, '+'))
                if '-' in r.get(ax): out.append((ax, '-'))
        if 'x' in r:
            if self.MPIrank == 0:
                if '+' in r.get('x') and self.MPIsize == 1: out.append(('x', '+'))
                if '-' in r.get('x'): out.append(('x', '-'))
            elif self.MPIrank == (self.MPIsize - 1) and self.MPIsize != 1:
                if '+' in r.get('x'): out.append(('x', '+'))
        return out

    def _pml_axis_rule(self, half, ax, side):
        """(profile slice, field range along the axis, psi offset) -- the odd/even
        half-cell sampling table of space.py:1110-1712."""
        P = self.npml
        N = self.loc_grid['xyz'.index(ax)]
        m = self.method
        S = slice
        if half == 'H':
            if side == '+':
                if m == 'PSTD': return S(0, None, 2), (N - P, N), 0
                if m == 'SHPF' and ax != 'x': return S(1, None, 2), (N - P, N), 0
                return S(1, -1, 2), (N - P, N - 1), 0
            if m == 'PSTD' and ax != 'z': return S(-1, None, -2), (0, P), 0
            return S(-2, None, -2), (0, P), 0
        if side == '+':
            return S(0, None, 2), (N - P, N), 0
        if m == 'PSTD':
            return (S(-2, None, -2) if ax == 'z' else S(-1, None, -2)), (0, P), 0
        if m == 'SHPF' and ax != 'x':
            return S(-1, None, -2), (0, P), 0
        return S(-3, None, -2), (1, P), 1

    # component / derivative-slot / sign of the two terms of each (half, axis)
    _PML_ROWS = {
        ('H', 'x'): (('Hy', _lib.D_XFZ, -1.), ('Hz', _lib.D_XFY, +1.)),
        ('H', 'y'): (('Hx', _lib.D_YFZ, +1.), ('Hz', _lib.D_YFX, -1.)),
        ('H', 'z'): (('Hx', _lib.D_ZFY, -1.), ('Hy', _lib.D_ZFX, +1.)),
        ('E', 'x'): (('Ey', _lib.D_XFZ, -1.), ('Ez', _lib.D_XFY, +1.)),
        ('E', 'y'): (('Ex', _lib.D_YFZ, +1.), ('Ez', _lib.D_YFX, -1.)),
        ('E', 'z'): (('Ex', _lib.D_ZFY, -1.), ('Ey', _lib.D_ZFX, +1.)),
    }

    def _pml_terms(self):
        """All CPML correction terms of this slab as boxes + gathered profiles."""
        terms = []
        if not self.PMLregion or self.npml == 0:
            return terms
        boxes = self._main_boxes()
        last = self.MPIrank == self.MPIsize - 1
        for half in ('H', 'E'):
            for ax, side in self._pml_faces():
                a = 'xyz'.index(ax)
                prof, rng, poff = self._pml_axis_rule(half, ax, side)
                b = np.asarray(getattr(self, 'PMLb' + ax), dtype=np.float64)[prof]
                aa = np.asarray(getattr(self, 'PMLa' + ax), dtype=np.float64)[prof]
                kap = np.asarray(getattr(self, 'PMLkappa' + ax), dtype=np.float64)[prof]
                kf = (1. / kap - 1.)
                assert len(b) == rng[1] - rng[0], (half, ax, side, len(b), rng)
                for name, diff, sign in self._PML_ROWS[(half, ax)]:
                    box = [list(r) for r in boxes[name]]
                    box[a] = list(rng)
                    if (self.method == 'FDTD' and half == 'H' and ax == 'z' and side == '-'
                            and name == 'Hy' and last):
                        box[1] = [0, self.Ny - 1]          # space.py:1647-1648
                    terms.append(dict(half=0 if half == 'H' else 1, comp='xyz'.index(name[1]), diff=diff,
                                      axis=a, lo=[r[0] for r in box], hi=[r[1] for r in box],
                                      psi_off=poff, sign=sign, b=np.ascontiguousarray(b),
                                      a=np.ascontiguousarray(aa), kf=np.ascontiguousarray(kf),
                                      name=f'psi_{name.lower()}{ax}_{"p" if side == "+" else "m"}'))
        return terms

    def _full_multiplier(self, ik, shift, kB, n):
        """Full-spectrum multiplier (ik - i kB) * shift; Hermitian extension of the rfft
        table for real fields (irfftn ignores Im of the DC and Nyquist bins)."""
        # the reference forms iky*ypshift in mmtdtype before it meets the spectrum
        # (space.py:709-718): round the product there, then widen
        m = (ik.ravel() * shift.ravel()).astype(self.mmtdtype).astype(np.complex128)
        if kB:
            m = m - 1j * kB * shift.ravel().astype(np.complex128)
        if np.dtype(self.field_dtype).kind == 'c':
            assert m.size == n
            return m
        if n % 2:
            # irfftn without an explicit length returns 2*(bins-1) samples: the reference's real-field
            # spectral derivative only works on even axes (space.py:145-162)
            raise ValueError(f"real field dtype with an odd spectral axis ({n} points): use an even length "
                             "or a complex field dtype")
        full = np.zeros(n, dtype=np.complex128)
        h = n // 2
        full[:h + 1] = m[:h + 1]
        full[0] = m[0].real
        full[h] = m[h].real
        full[h + 1:] = np.conj(m[1:h][::-1])
        return full

    def _finalize(self):
        """Push boxes, multipliers, ghost rules and CPML terms to the engine."""
        lib = self._lib
        if not self._coeff_ready:
            raise RuntimeError("init_update_constants() must be called before updateH/updateE")
        for name, bx in self._main_boxes().items():
            _lib.check(lib.ies_set_update_box(self._ctx, _lib.COMP[name],
                                              _lib.I3(*[r[0] for r in bx]), _lib.I3(*[r[1] for r in bx])))
        _lib.check(lib.ies_set_neighbours(self._ctx, int(self.MPIrank > 0),
                                          int(self.MPIrank < self.MPIsize - 1)))
        bbc = (self.apply_BBCx, self.apply_BBCy, self.apply_BBCz)
        pbc = (self.apply_PBCx, self.apply_PBCy, self.apply_PBCz)
        if self.method in ('SHPF', 'PSTD'):
            one = lambda v: np.ones_like(v)
            tabs = {1: (self.iky, self.ypshift, self.ymshift, self.Ny),
                    2: (self.ikz, self.zpshift, self.zmshift, self.Nz)}
            if self.method == 'PSTD':
                tabs[0] = (self.ikx, self.xpshift, self.xmshift, self.Nx)
            for ax, (ik, ps, ms, n) in tabs.items():
                kB = 0.
                if self.BBC_called and bbc[ax] == True:
                    if self.mmt is None:
                        raise AttributeError("Bloch boundary needs a source.Setter (space.mmt) first")
                    kB = self.mmt[ax]
                    if ('xyz'[ax] in self.PMLregion) and self.PMLregion.get('xyz'[ax]) != '':
                        raise NotImplementedError("Bloch boundary and CPML on the same axis")
                for half, sh in ((_lib.HALF_H, ps), (_lib.HALF_E, ms)):
                    shift = sh if self.method == 'SHPF' else one(sh)
                    full = self._full_multiplier(ik, shift, kB, n)
                    buf = np.ascontiguousarray(full).view(np.float64)
                    _lib.check(lib.ies_set_multiplier(self._ctx, half, ax, _lib.dptr(buf), n))
            # SHPF, x axis: ghost-plane copies of the differentiated field after each update when
            # apply_BBC was called with a True entry (space.py:1898-1912, 2073-2085)
            gx = None
            if self.method == 'SHPF' and self.BBC_called:
                if bbc[0] == True:
                    gx = self.Lx - 2 * self.dx
                elif pbc[0] == True:
                    gx = 0.
            if gx is not None:
                if bbc[1] == True or bbc[2] == True:
                    # the reference evaluates the y/z Bloch terms on the fields AFTER the x ghost
                    # copies (space.py:1898-1930); the fused multiplier sees them before
                    raise NotImplementedError("SHPF: Bloch/periodic x axis together with a Bloch y or z axis")
                if self.mmt is None:
                    raise AttributeError("Bloch boundary needs a source.Setter (space.mmt) first")
                pp, pm = np.exp(+1j * self.mmt[0] * gx), np.exp(-1j * self.mmt[0] * gx)
                _lib.check(lib.ies_set_ghost(self._ctx, 0, 1, pp.real, pp.imag, pm.real, pm.imag))
            else:
                _lib.check(lib.ies_set_ghost(self._ctx, 0, 0, 1., 0., 1., 0.))
        else:
            # FDTD ghost-plane copies (space.py:1798-1858, 1981-2033)
            L = (self.Lx, self.Ly, self.Lz)
            d = (self.dx, self.dy, self.dz)
            for ax in range(3):
                on = False
                if self.BBC_called and bbc[ax] == True:
                    on, newL = True, L[ax] - 2 * d[ax]
                elif self.PBC_called and pbc[ax] == True:
                    on, newL = True, 0
                if on:
                    if self.mmt is None:
                        raise AttributeError("'Basic3D' object has no attribute 'mmt'")   # Q6
                    k = self.mmt[ax]
                    pp, pm = np.exp(+1j * k * newL), np.exp(-1j * k * newL)
                    _lib.check(lib.ies_set_ghost(self._ctx, ax, 1, pp.real, pp.imag, pm.real, pm.imag))
                else:
                    _lib.check(lib.ies_set_ghost(self._ctx, ax, 0, 1., 0., 1., 0.))
        # CPML terms: psi arrays are engine-owned; rebuilding resets them, so only on change
        sig = (self.method, self.MPIrank, self.MPIsize, repr(sorted(self.PMLregion.items())), self.npml)
        if getattr(self, '_pml_sig', None) != sig:
            _lib.check(lib.ies_clear_pml(self._ctx))
            self._keep = []
            for t in self._pml_terms():
                pt = _lib.PmlTerm(t['half'], t['comp'], t['diff'], t['axis'], _lib.I3(*t['lo']),
                                  _lib.I3(*t['hi']), t['psi_off'], self.npml, t['sign'],
                                  _lib.dptr(t['b']), _lib.dptr(t['a']), _lib.dptr(t['kf']))
                _lib.check(lib.ies_add_pml_term(self._ctx, C.byref(pt)))
            self._pml_sig = sig
        self._dirty = False

    # ------------------------------------------------------------- hot path
    def updateH(self, tstep):
        """space.py:639-840: halo (Ey[0],Ez[0] from rank+1), derivatives, update, Bloch, CPML."""
        if self._dirty: self._finalize()
        self._half_step(_lib.HALF_H, tstep)

    def updateE(self, tstep):
        """space.py:842-1052: halo (Hy[-1],Hz[-1] from rank-1), derivatives, update, Bloch, CPML."""
        if self._dirty: self._finalize()
        self._half_step(_lib.HALF_E, tstep)

    def _half_step(self, half, tstep):
        """Halo exchange + update.  IpcComm / LocalComm only ENQUEUE the transfer (copy engine) and the
        stream-ordered wait in front of the update; nothing blocks the host.  A communicator with
        exchange_begin/exchange_end (tests/torch_comm.TorchComm) runs its transfer on its own stream
        while the engine does the part of the half-step that needs no neighbour plane."""
        comm = self.MPIcomm
        if self.MPIsize > 1 and hasattr(comm, 'exchange_begin'):
            token = comm.exchange_begin(self, half)
            _lib.check(self._lib.ies_update_phase(self._ctx, half, 0))
            comm.exchange_end(self, token)
            _lib.check(self._lib.ies_update_phase(self._ctx, half, 1))
            return
        if self.MPIsize > 1: comm.exchange(self, half)
        fn = self._lib.ies_update_h if half == _lib.HALF_H else self._lib.ies_update_e
        _lib.check(fn(self._ctx, int(tstep)))


class Empty3D(Basic3D):
    """Scattered-field view SF = TF - IF (space.py:2151-2179).  get_SF only records
    the pair; the difference is evaluated on the device where a collector samples
    it, or on the host when a field is read."""

    def __init__(self, grid, gridgap, dt, tsteps, field_dtype, mmtdtype, **kwargs):
        self._pair = None
        self._lazy_ctx = True           # no device fields until somebody needs them
        Basic3D.__init__(self, grid, gridgap, dt, tsteps, field_dtype, mmtdtype, **kwargs)

    def get_SF(self, TF, IF):
        self._pair = (TF, IF)


class _LazyDiff:
    def __init__(self, sf, name): self._sf, self._name = sf, name
    shape = property(lambda self: self._sf.loc_grid)
    dtype = property(lambda self: np.dtype(self._sf.field_dtype))
    def __array__(self, dtype=None, copy=None):
        TF, IF = self._sf._pair
        a = np.asarray(getattr(TF, self._name)) - np.asarray(getattr(IF, self._name))
        return a if dtype is None else a.astype(dtype)
    def __getitem__(self, idx):
        TF, IF = self._sf._pair
        return getattr(TF, self._name)[idx] - getattr(IF, self._name)[idx]


def _sf_property(name):
    base = getattr(Basic3D, name)

    def get(self):
        if self._pair is not None:
            return _LazyDiff(self, name)
        return base.fget(self)
    return property(get, base.fset)


for _n in _FIELDS:
    setattr(Empty3D, _n, _sf_property(_n))


def gather_local(spaces, name):
    """Concatenate one field of the slabs of a LocalGroup decomposition along x."""
    return np.concatenate([np.asarray(getattr(s, name)) for s in spaces], axis=0)
