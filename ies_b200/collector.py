"""collector.Sx / Sy / Sz / FieldAtPoint on the b200 engine (reference: collector.py).

The running DFT  DFT_F += F[plane] * exp(2 pi i f t dt) * dt  (collector.py:323-338,
508-527, 716-735) is accumulated on the device by ies_dft_accumulate; an Empty3D
scattered-field space contributes TF - IF on the collector plane only.  Index
logic (Python `round`, per-rank x clipping) and the .npy file names / array
shapes of get_S* follow the cited lines.
"""
import ctypes as C
import os

import numpy as np

try:
    from . import _lib, comm as _comm
except ImportError:
    import _lib
    import comm as _comm


def _pair(space):
    """(ctx_a, ctx_b): a plain space or the (TF, IF) pair behind an Empty3D."""
    p = getattr(space, '_pair', None)
    if p is not None:
        return p[0]._ctx, p[1]._ctx, p[0]
    return space._ctx, None, space


class collector:

    def __init__(self, name, path, space, engine):
        self.engine = engine
        self.xp = np
        self.name = name
        self.space = space
        self.path = path
        if self.space.MPIrank == 0:
            if not os.path.exists(self.path): os.makedirs(self.path, exist_ok=True)
        self.gloc = None
        self.lloc = None

    def _get_local_x_loc(self, gxsrts, gxends):
        """collector.py:30-118: the one shared statement of the rule lives in comm.local_x_loc."""
        return _comm.local_x_loc(self.space, gxsrts, gxends)

    # ---- device DFT plumbing shared by Sx/Sy/Sz
    def _ensure(self):
        if self._h is None:
            a, b, owner = _pair(self.space)
            freqs = np.ascontiguousarray(np.asarray(self.freqs, dtype=np.float64))
            h = C.c_void_p()
            _lib.check(owner._lib.ies_dft_create(a, _lib.I3(*self._lo), _lib.I3(*self._hi),
                                                 _lib.I4(*[_lib.COMP[c] for c in self._comps]),
                                                 _lib.dptr(freqs), len(freqs), C.byref(h)))
            self._h = h
            self._owner = owner

    def _accumulate(self, tstep):
        self._ensure()
        a, b, owner = _pair(self.space)
        _lib.check(owner._lib.ies_dft_accumulate(self._h, a, b, int(tstep)))

    def _read(self, which, shape):
        self._ensure()
        out = np.empty(shape, dtype=np.complex128)
        _lib.check(self._owner._lib.ies_dft_read(self._h, which, out.ctypes.data_as(C.c_void_p)))
        return out

    def __del__(self):
        try:
            if getattr(self, '_h', None) is not None:
                self._owner._lib.ies_dft_destroy(self._h)
                self._h = None
        except Exception:
            pass


class FieldAtPoint(collector):
    """collector.py:123-261: the six components at one cell, every step."""

    def __init__(self, name, path, space, loc, engine):
        collector.__init__(self, name, path, space, engine)
        self.loc = loc
        if len(self.loc) == 3:
            self.xloc = round(loc[0] / space.dx)
            self.yloc = round(loc[1] / space.dy)
            self.zloc = round(loc[2] / space.dz)
        elif len(self.loc) == 2:
            self.xloc = round(loc[0] / space.dx)
            self.yloc = round(loc[1] / space.dy)
        self.gxloc, self.lxloc = collector._get_local_x_loc(self, self.xloc, self.xloc)
        self._p = None
        if self.gxloc != None:
            a, b, owner = _pair(space)
            p = C.c_void_p()
            _lib.check(owner._lib.ies_probe_create(a, self.lxloc[0], self.yloc, self.zloc,
                                                   int(space.tsteps), C.byref(p)))
            self._p, self._owner = p, owner
            self._synced = False

    def get_time_signal(self, tstep):
        if self.gxloc != None:
            a, b, owner = _pair(self.space)
            _lib.check(owner._lib.ies_probe_record(self._p, a, b, int(tstep)))
            self._synced = False

    def _fetch(self):
        if self.gxloc != None and not self._synced:
            self._sig = {}
            for q, n in enumerate(('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')):
                out = np.empty(self.space.tsteps, dtype=self.space.field_dtype)
                _lib.check(self._owner._lib.ies_probe_read(self._p, q, out.ctypes.data_as(C.c_void_p)))
                self._sig[n] = out
            self._synced = True
        return self._sig

    Ex_t = property(lambda self: self._fetch()['Ex'])
    Ey_t = property(lambda self: self._fetch()['Ey'])
    Ez_t = property(lambda self: self._fetch()['Ez'])
    Hx_t = property(lambda self: self._fetch()['Hx'])
    Hy_t = property(lambda self: self._fetch()['Hy'])
    Hz_t = property(lambda self: self._fetch()['Hz'])

    def save_time_signal(self, **kwargs):
        self.space.MPIcomm.barrier()
        self.binary = True
        self.txt = False
        if kwargs.get('binary') != None:
            hey = kwargs.get('binary')
            assert hey == True or hey == False
            self.binary = hey
        if kwargs.get('txt') != None:
            hey = kwargs.get('txt')
            assert hey == True or hey == False
            self.txt = hey
        if self.gxloc != None:
            sig = self._fetch()
            if self.binary == True:
                for n, v in sig.items():
                    np.save("{}/{}_{}_t.npy".format(self.path, self.name, n), v)
            if self.txt == True:
                for n, v in sig.items():
                    np.savetxt("{}/{}_{}_t.txt".format(self.path, self.name, n), v, newline='\n',
                               fmt='%1.15f+%1.15fi')

    def __del__(self):
        try:
            if getattr(self, '_p', None) is not None:
                self._owner._lib.ies_probe_destroy(self._p)
                self._p = None
        except Exception:
            pass


class Sx(collector):
    """collector.py:266-382."""

    def __init__(self, name, path, space, xloc, srt, end, freqs, engine):
        collector.__init__(self, name, path, space, engine)
        self.Nf = len(freqs)
        self.freqs = freqs
        self.xsrt = round(xloc / space.dx)
        self.ysrt = round(srt[0] / space.dy)
        self.zsrt = round(srt[1] / space.dz)
        self.xend = self.xsrt + 1
        self.yend = round(end[0] / space.dy)
        self.zend = round(end[1] / space.dz)
        self.gxloc, self.lxloc = collector._get_local_x_loc(self, self.xsrt, self.xend)
        self._h = None
        if self.gxloc != None:
            self._shape3 = (self.Nf, self.yend - self.ysrt, self.zend - self.zsrt)
            self._lo = (self.lxloc[0], self.ysrt, self.zsrt)
            self._hi = (self.lxloc[1], self.yend, self.zend)
            self._comps = ('Ey', 'Ez', 'Hy', 'Hz')

    def do_RFT(self, tstep):
        if self.gxloc != None:
            self._accumulate(tstep)

    def _pull(self):
        self.DFT_Ey = self._read(0, self._shape3)
        self.DFT_Ez = self._read(1, self._shape3)
        self.DFT_Hy = self._read(2, self._shape3)
        self.DFT_Hz = self._read(3, self._shape3)

    def get_Sx(self, tstep, h5=False):
        self.space.MPIcomm.barrier()
        if self.gxloc != None:
            self._pull()
            self.Sx = 0.5 * ((self.DFT_Ey.real * self.DFT_Hz.real) + (self.DFT_Ey.imag * self.DFT_Hz.imag)
                             - (self.DFT_Ez.real * self.DFT_Hy.real) - (self.DFT_Ez.imag * self.DFT_Hy.imag))
            self.Sx_area = self.Sx.sum(axis=(1, 2)) * self.space.dy * self.space.dz
            r = self.space.MPIrank
            np.save(f"{self.path}{self.name}_DFT_Ey_{tstep:07d}tstep_rank{r:02d}", self.DFT_Ey)
            np.save(f"{self.path}{self.name}_DFT_Ez_{tstep:07d}tstep_rank{r:02d}", self.DFT_Ez)
            np.save(f"{self.path}{self.name}_DFT_Hy_{tstep:07d}tstep_rank{r:02d}", self.DFT_Hy)
            np.save(f"{self.path}{self.name}_DFT_Hz_{tstep:07d}tstep_rank{r:02d}", self.DFT_Hz)
            np.save(f"{self.path}{self.name}_{tstep:07d}tstep_area", self.Sx_area)
            if h5 == True:
                import h5py
                with h5py.File(f'{self.path}{self.name}_DFTs_{tstep:07d}tstep_rank{r:02d}.h5', 'w') as hf:
                    hf.create_dataset('Sx_Ey', data=self.DFT_Ey)
                    hf.create_dataset('Sx_Ez', data=self.DFT_Ez)
                    hf.create_dataset('Sx_Hy', data=self.DFT_Hy)
                    hf.create_dataset('Sx_Hz', data=self.DFT_Hz)
                    hf.create_dataset('Sx_area', data=self.Sx_area)


class _Splane(collector):
    """Common part of Sy (collector.py:387-594) and Sz (599-801): planes that span
    the x-slabs; every rank accumulates its part, rank 0 concatenates along x."""

    def _setup(self, space, xsrt, ysrt, zsrt, xend, yend, zend, freqs, tag, comps):
        self.Nf = len(freqs)
        self.freqs = freqs
        self.xsrt, self.ysrt, self.zsrt = xsrt, ysrt, zsrt
        self.xend, self.yend, self.zend = xend, yend, zend
        who_g, who_l = {}, {}
        for MPIrank in range(self.space.MPIsize):
            node_xsrt, node_xend = self.space.myNx_indice[MPIrank]
            if xsrt < node_xsrt and xend > node_xsrt and xend <= node_xend:
                who_g[MPIrank] = ((node_xsrt, ysrt, zsrt), (xend, yend, zend))
                who_l[MPIrank] = ((0, ysrt, zsrt), (xend - node_xsrt, yend, zend))
            if xsrt < node_xsrt and xend > node_xend:
                who_g[MPIrank] = ((node_xsrt, ysrt, zsrt), (node_xend, yend, zend))
                who_l[MPIrank] = ((0, ysrt, zsrt), (node_xend - node_xsrt, yend, zend))
            if xsrt >= node_xsrt and xsrt < node_xend and xend <= node_xend:
                who_g[MPIrank] = ((xsrt, ysrt, zsrt), (xend, yend, zend))
                who_l[MPIrank] = ((xsrt - node_xsrt, ysrt, zsrt), (xend - node_xsrt, yend, zend))
            if xsrt >= node_xsrt and xsrt < node_xend and xend > node_xend:
                who_g[MPIrank] = ((xsrt, ysrt, zsrt), (node_xend, yend, zend))
                who_l[MPIrank] = ((xsrt - node_xsrt, ysrt, zsrt), (node_xend - node_xsrt, yend, zend))
        setattr(self, f'who_get_{tag}_gxloc', who_g)
        setattr(self, f'who_get_{tag}_lxloc', who_l)
        self._who = who_l
        self.space.MPIcomm.barrier()
        self._h = None
        self._comps = comps
        if self.space.MPIrank in who_l:
            self.gloc = who_g[self.space.MPIrank]
            self.lloc = who_l[self.space.MPIrank]
            self._lo, self._hi = tuple(self.lloc[0]), tuple(self.lloc[1])

    def do_RFT(self, tstep):
        if self.space.MPIrank in self._who:
            self._accumulate(tstep)

    def _save_parts(self, tstep, shape):
        parts = {}
        if self.space.MPIrank in self._who:
            for q, n in enumerate(self._comps):
                arr = self._read(q, shape)
                setattr(self, 'DFT_' + n, arr)
                np.save(f"{self.path}{self.name}_DFT_{n}_{tstep:07d}tstep_rank{self.space.MPIrank:02d}", arr)
        self.space.MPIcomm.barrier()
        if self.space.MPIrank == 0:
            for n in self._comps:
                parts[n] = np.concatenate(
                    [np.load(f"{self.path}{self.name}_DFT_{n}_{tstep:07d}tstep_rank{rank:02d}.npy")
                     for rank in self._who], axis=1)
        return parts


class Sy(_Splane):

    def __init__(self, name, path, space, yloc, srt, end, freqs, engine):
        collector.__init__(self, name, path, space, engine)
        xsrt = round(srt[0] / space.dx)
        ysrt = round(yloc / space.dy)
        zsrt = round(srt[1] / space.dz)
        xend = round(end[0] / space.dx)
        yend = ysrt + 1
        zend = round(end[1] / space.dz)
        self._setup(space, xsrt, ysrt, zsrt, xend, yend, zend, freqs, 'Sy', ('Ex', 'Ez', 'Hx', 'Hz'))

    def get_Sy(self, tstep, h5=False):
        self.space.MPIcomm.barrier()
        shape = None
        if self.space.MPIrank in self._who:
            shape = (self.Nf, self.lloc[1][0] - self.lloc[0][0], self.zend - self.zsrt)
        p = self._save_parts(tstep, shape)
        if self.space.MPIrank == 0:
            DFT_Ex, DFT_Ez, DFT_Hx, DFT_Hz = p['Ex'], p['Ez'], p['Hx'], p['Hz']
            self.Sy = 0.5 * (-(DFT_Ex.real * DFT_Hz.real) - (DFT_Ex.imag * DFT_Hz.imag)
                             + (DFT_Ez.real * DFT_Hx.real) + (DFT_Ez.imag * DFT_Hx.imag))
            self.Sy_area = self.Sy.sum(axis=(1, 2)) * self.space.dx * self.space.dz
            np.save(f"{self.path}{self.name}_{tstep:07d}tstep_area", self.Sy_area)


class Sz(_Splane):

    def __init__(self, name, path, space, zloc, srt, end, freqs, engine):
        collector.__init__(self, name, path, space, engine)
        xsrt = round(srt[0] / space.dx)
        ysrt = round(srt[1] / space.dy)
        zsrt = round(zloc / space.dz)
        xend = round(end[0] / space.dx)
        yend = round(end[1] / space.dz)      # sic: the reference divides by dz (collector.py:641)
        zend = zsrt + 1
        self._setup(space, xsrt, ysrt, zsrt, xend, yend, zend, freqs, 'Sz', ('Ex', 'Ey', 'Hx', 'Hy'))

    def get_Sz(self, tstep, h5=False):
        self.space.MPIcomm.barrier()
        shape = None
        if self.space.MPIrank in self._who:
            shape = (self.Nf, self.lloc[1][0] - self.lloc[0][0], self.yend - self.ysrt)
        p = self._save_parts(tstep, shape)
        if self.space.MPIrank == 0:
            DFT_Ex, DFT_Ey, DFT_Hx, DFT_Hy = p['Ex'], p['Ey'], p['Hx'], p['Hy']
            self.Sz = 0.5 * (-(DFT_Ey.real * DFT_Hx.real) - (DFT_Ey.imag * DFT_Hx.imag)
                             + (DFT_Ex.real * DFT_Hy.real) + (DFT_Ex.imag * DFT_Hy.imag))
            self.Sz_area = self.Sz.sum(axis=(1, 2)) * self.space.dx * self.space.dy
            np.save(f"{self.path}{self.name}_{tstep:07d}tstep_area", self.Sz_area)
