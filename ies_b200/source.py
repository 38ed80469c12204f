"""source.Setter and the pulse shapes on the b200 engine (reference: source.py).

Index logic (Python banker's `round`, the zero-length-x decrement, the owning
rank) is a line-by-line restatement of source.py:57-165 -- the host API mirror, whose
integers must equal the reference's -- and the pulse classes keep the reference's formulas;
the injection itself is one small kernel on the device (ies_put_src) instead of a sliced
`+=` on the array.
"""
import numpy as np
from scipy.constants import c, mu_0, epsilon_0

try:
    from . import _lib
except ImportError:
    import _lib


class Setter:

    def __init__(self, space, src_srt, src_end, mmt):
        """Set the position of the source (source.py:8-165)."""
        self.space = space
        self.xp = np
        self.who_put_src = None

        self.src_xsrt = round(src_srt[0] / self.space.dx)
        self.src_xend = round(src_end[0] / self.space.dx)
        self.src_ysrt = round(src_srt[1] / self.space.dy)
        self.src_yend = round(src_end[1] / self.space.dy)
        if space.dimension == 3:
            self.src_zsrt = round(src_srt[2] / self.space.dz)
            self.src_zend = round(src_end[2] / self.space.dz)

        self.space.MPIcomm.Barrier()

        for rank in range(self.space.MPIsize):
            my_xsrt = self.space.myNx_indice[rank][0]
            my_xend = self.space.myNx_indice[rank][1]

            # x position of the source has zero length (source.py:88)
            if self.src_xsrt == self.src_xend: self.src_xsrt = self.src_xend - 1

            if self.src_xsrt == (self.src_xend - 1):
                if self.src_xsrt >= my_xsrt and self.src_xend <= my_xend:
                    self.who_put_src = rank
                    if self.space.MPIrank == self.who_put_src:
                        self.my_src_xsrt = self.src_xsrt - my_xsrt
                        self.my_src_xend = self.src_xend - my_xsrt
                        self.src = np.zeros(self.space.tsteps, dtype=self.space.field_dtype)
                else:
                    continue
            elif self.src_xsrt < self.src_xend:
                assert self.space.MPIsize == 1
                self.who_put_src = 0
                self.my_src_xsrt = self.src_xsrt
                self.my_src_xend = self.src_xend
                self.src = np.zeros(self.space.tsteps, dtype=self.space.field_dtype)
            elif self.src_xsrt > self.src_xend:
                raise ValueError("src_end[0] should be bigger than src_srt[0]")
            else:
                raise ValueError('x location of the source is not defined!')

        # momentum of the source: also the Bloch vector of the space (source.py:139)
        self.space.mmt = mmt
        self.space._dirty = True

        if self.space.MPIrank == self.who_put_src:
            kx, ky = mmt[0], mmt[1]
            self.px = np.exp(+1j * kx * np.arange(self.my_src_xsrt, self.my_src_xend) * self.space.dx)
            self.py = np.exp(+1j * ky * np.arange(self.src_ysrt, self.src_yend) * self.space.dy)
            xdist = self.my_src_xend - self.my_src_xsrt
            ydist = self.src_yend - self.src_ysrt
            if xdist == 1: self.px = np.exp(1j * kx * np.arange(1) * self.space.dx)
            if ydist == 1: self.py = np.exp(1j * ky * np.arange(1) * self.space.dy)
            if space.dimension == 3:
                kz = mmt[2]
                self.pz = np.exp(+1j * kz * np.arange(self.src_zsrt, self.src_zend) * self.space.dz)
                zdist = self.src_zend - self.src_zsrt
                if zdist == 1: self.pz = np.exp(1j * kz * np.arange(1) * self.space.dz)

    def put_src(self, where, pulse, put_type):
        """Put the source into the designated field (source.py:167-253)."""
        self.put_type = put_type
        self.where = where
        self.pulse = pulse

        if self.space.MPIrank != self.who_put_src:
            return
        if put_type not in ('soft', 'hard'):
            raise ValueError("Please insert 'soft' or 'hard'")
        # the reference compares with exactly two spellings per component, 'Ex' / 'ex'
        # (source.py:236-251), and silently ignores anything else
        if where not in _lib.COMP and not (len(where) == 2 and where[0] in 'eh' and where[1] in 'xyz'):
            return
        name = where[0].upper() + where[1]
        sp = self.space
        # NumPy slice semantics of `F[x0:x1, y0:y1, z0:z1]`: extents are clamped to the array and an
        # inverted range is empty
        rng = [slice(a, b).indices(n) for a, b, n in ((self.my_src_xsrt, self.my_src_xend, sp.loc_grid[0]),
                                                       (self.src_ysrt, self.src_yend, sp.loc_grid[1]),
                                                       (self.src_zsrt, self.src_zend, sp.loc_grid[2]))]
        lo = tuple(r[0] for r in rng)
        hi = tuple(max(r[0], r[1]) for r in rng)
        if any(h <= l for l, h in zip(lo, hi)):
            return
        real_field = np.dtype(sp.field_dtype).kind != 'c'
        if real_field and (isinstance(pulse, complex) or np.iscomplexobj(pulse)):
            # what `real_array[...] += complex` / `real_array[...] = complex` do in NumPy
            if put_type == 'soft':
                raise TypeError("Cannot cast ufunc 'add' output from dtype('complex128') to "
                                f"dtype('{np.dtype(sp.field_dtype).name}') with casting rule 'same_kind'")
            raise TypeError("float() argument must be a string or a real number, not 'complex'")
        pv = complex(pulse)
        px = py = pz = None
        if sp.BBC_called == True:
            # phase tables cut to the clamped box (they are indexed from the unclamped start)
            cut = lambda t, l, h, s0: (t if t.size == 1 else t[l - s0:h - s0])
            px = np.ascontiguousarray(cut(self.px, lo[0], hi[0], self.my_src_xsrt), dtype=np.complex128)
            py = np.ascontiguousarray(cut(self.py, lo[1], hi[1], self.src_ysrt), dtype=np.complex128)
            pz = np.ascontiguousarray(cut(self.pz, lo[2], hi[2], self.src_zsrt), dtype=np.complex128)
            self._keep = (px, py, pz)
        vp = lambda a: None if a is None else a.ctypes.data
        _lib.check(sp._lib.ies_put_src(sp._ctx, _lib.COMP[name], _lib.I3(*lo), _lib.I3(*hi),
                                       pv.real, pv.imag, int(put_type == 'hard'), vp(px), vp(py), vp(pz)))


class Gaussian:
    """source.py:256-354."""

    def __init__(self, dt, center_wv, spread, pick_pos, dtype):
        self.dt = dt
        self.dtype = dtype
        self.wvlenc = center_wv
        self.spread = spread
        self.pick_pos = pick_pos
        self.freqc = c / self.wvlenc
        self.w0 = 2 * np.pi * self.freqc
        self.ws = self.spread * self.w0
        self.ts = 1. / self.ws
        self.tc = self.pick_pos * self.dt

    def pulse_c(self, step):
        return np.exp((-.5) * (((step * self.dt - self.tc) * self.ws) ** 2)) * \
            np.exp(-1j * self.w0 * (step * self.dt - self.tc))

    def pulse_re(self, step):
        return np.exp((-.5) * (((step * self.dt - self.tc) * self.ws) ** 2)) * \
            np.cos(self.w0 * (step * self.dt - self.tc))

    def pulse_im(self, step):
        return np.exp((-.5) * (((step * self.dt - self.tc) * self.ws) ** 2)) * \
            -np.sin(self.w0 * (step * self.dt - self.tc))

    def plot_pulse(self, tsteps, freqs, savedir):
        """source.py:299-354 draws the pulse and its spectrum with matplotlib; here the
        same series are computed and saved as .npy when matplotlib is unavailable."""
        import os
        time_domain = np.arange(tsteps, dtype=np.float64) * self.dt
        t = time_domain
        pulse_re = np.exp((-.5) * (((t - self.tc) * self.ws) ** 2)) * np.cos(self.w0 * (t - self.tc))
        pulse_im = np.exp((-.5) * (((t - self.tc) * self.ws) ** 2)) * -np.sin(self.w0 * (t - self.tc))
        os.makedirs(savedir, exist_ok=True)
        try:
            import matplotlib
            matplotlib.use('Agg')
            import matplotlib.pyplot as plt
            fig, ax = plt.subplots(1, 1, figsize=(8, 4))
            ax.plot(time_domain, pulse_re, label='real')
            ax.plot(time_domain, pulse_im, label='imag')
            ax.legend(); ax.grid(True)
            fig.savefig(savedir + "src_theoretical.png")
            plt.close(fig)
        except Exception:
            np.save(savedir + "src_theoretical.npy", np.stack([time_domain, pulse_re, pulse_im]))


class _Mono:
    def set_freq(self, freq):
        self.freq = freq
        self.wvlen = c / self.freq
        self.omega = 2 * np.pi * self.freq
        self.wvector = 2 * np.pi / self.wvlen

    def set_wvlen(self, wvlen):
        self.wvlen = wvlen
        self.freq = c / self.wvlen
        self.omega = 2 * np.pi * self.freq
        self.wvector = 2 * np.pi / self.wvlen


class Sine(_Mono):
    """source.py:357-383."""
    def __init__(self, dt, dtype):
        self.dt, self.dtype = dt, dtype

    def signal(self, tstep):
        return np.sin(self.omega * tstep * self.dt)


class Cosine(_Mono):
    """source.py:386-411."""
    def __init__(self, dt, dtype):
        self.dt, self.dtype = dt, dtype

    def signal(self, tstep):
        return np.cos(self.omega * tstep * self.dt)


class Harmonic(_Mono):
    """source.py:414-438."""
    def __init__(self, dt):
        self.dt = dt

    def apply(self, tstep):
        return np.exp(-1j * self.omega * tstep * self.dt)


class Smoothing:
    """source.py:441-455."""
    def __init__(self, dt, threshold):
        self.dt, self.threshold = dt, threshold

    def apply(self, tstep):
        if tstep < self.threshold: return tstep / self.threshold
        return 1.


class SmoothInOut:
    """source.py:458-476."""
    def __init__(self, dt, inc, dec):
        self.dt, self.inc, self.dec = dt, inc, dec

    def apply(self, tstep):
        if tstep < self.inc: return tstep / self.inc
        elif tstep >= self.inc and tstep <= self.dec: return 1
        elif tstep > self.dec and tstep < (self.inc + self.dec):
            return (self.inc - tstep) / (self.inc - self.dec)
        return 0


class Delta:
    """source.py:479-488."""
    def __init__(self, pick):
        self.pick = pick

    def apply(self, tstep):
        if tstep == self.pick: return 1.
        else: return 0.
