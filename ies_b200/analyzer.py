"""analyzer.SpectrumAnalyzer for the b200 engine: band-structure post-processing of
`collector.FieldAtPoint` time signals (reference: analyzer.py:8-246).

The reference hands the signal to the third-party `harminv` package (Mandelshtam & Taylor's
filter-diagonalisation method, `hv.Harminv(signal, fmin, fmax, dt, nf)`, analyzer.py:189) and
prints its columns freq / decay / Q / amplitude / phase / error (analyzer.py:229-244).  harminv
is not vendored in the reference and not installed here, so `harminv_fdm` below is an
independent implementation of the published algorithm (V. A. Mandelshtam and H. S. Taylor,
J. Chem. Phys. 107, 6756 (1997), single-window FDM on a Fourier basis) returning the same
columns with harminv's sign conventions.  PARITY UNPINNED: no harminv output exists to compare
with; tests/test_analyzer.py checks it against signals whose modes are known exactly.

Host-side NumPy / SciPy only -- this is offline analysis, not the hot path.
"""
import os

import numpy as np
from scipy.constants import c

_FIELDS = ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz')


class HarminvResult:
    """Columns of harminv's output (the attributes analyzer.py:196-244 reads), sorted by frequency:
    signal ~ sum_k amplitude_k exp(-i (2 pi freq_k t - phase_k)) exp(-decay_k t)."""

    def __init__(self, freq, decay, amplitude, phase, error):
        self.freq = np.asarray(freq, dtype=np.float64)
        self.decay = np.asarray(decay, dtype=np.float64)
        with np.errstate(divide='ignore', invalid='ignore'):
            self.Q = np.pi * np.abs(self.freq) / self.decay
        self.amplitude = np.asarray(amplitude, dtype=np.float64)
        self.phase = np.asarray(phase, dtype=np.float64)
        self.error = np.asarray(error, dtype=np.float64)
        self.omega = 2 * np.pi * self.freq - 1j * self.decay

    def __len__(self):
        return self.freq.size

    def table(self):
        """Rows (freq, decay, Q, amplitude, phase, error) -- harminv's CLI column order."""
        return np.stack([self.freq, self.decay, self.Q, self.amplitude, self.phase, self.error], axis=1)


def _u_matrix(sig, z, p, M):
    """U^(p)_{jj'} = sum_{n,n'=0..M} c_{n+n'+p} z_j^{-n} z_j'^{-n'} (eq. 29-30 of the FDM paper),
    assembled from two O(J N) power sums per basis function instead of the O(J^2 N) double sum."""
    a = 1. / z                                        # (J,)
    m0 = np.arange(0, M + 1)
    m1 = np.arange(M + 1, 2 * M + 1)
    P0 = a[:, None] ** m0[None, :]                    # a^m, m = 0..M
    G0 = P0 @ sig[p:p + M + 1]                        # sum_{m<=M} c_{m+p} a^m
    P1 = a[:, None] ** (m1 - M - 1)[None, :]
    G1 = P1 @ sig[p + M + 1:p + 2 * M + 1]            # sum_{m>M} c_{m+p} a^{m-M-1}
    aM1 = a ** (M + 1)
    J = z.size
    A, B = a[:, None], a[None, :]
    with np.errstate(divide='ignore', invalid='ignore'):
        U = (B * G0[None, :] - A * G0[:, None] + (aM1[None, :] * A) * G1[:, None] - (aM1[:, None] * B) * G1[None, :]) / (B - A)
    w0 = (m0 + 1.)
    w1 = (2 * M - m1 + 1.)
    diag = (P0 * w0[None, :]) @ sig[p:p + M + 1] + aM1 * ((P1 * w1[None, :]) @ sig[p + M + 1:p + 2 * M + 1])
    U[np.arange(J), np.arange(J)] = diag
    return U


def harminv_fdm(signal, fmin, fmax, dt, nf=10, error_cut=0.1, amp_rel_cut=1e-4):
    """Harmonic inversion of `signal` (sampled every `dt`) in the window [fmin, fmax] with `nf` Fourier
    basis functions.  Returns a HarminvResult.  Modes are kept when fmin <= freq <= fmax, the
    decay is not negative beyond round-off, error <= error_cut and the amplitude is at least
    amp_rel_cut of the strongest mode (harminv's default screening has the same criteria)."""
    sig = np.asarray(signal, dtype=np.complex128).ravel()
    N = sig.size
    if N < 8:
        raise ValueError("signal too short")
    J = int(max(2, nf))
    M = (N - 3) // 2                                   # U^(2) needs c up to 2M + 2
    f = np.linspace(fmin, fmax, J)
    z = np.exp(-1j * 2 * np.pi * f * dt)               # basis z_j on the unit circle
    from scipy.linalg import eig
    U0 = _u_matrix(sig, z, 0, M)
    U1 = _u_matrix(sig, z, 1, M)
    U2 = _u_matrix(sig, z, 2, M)
    # regularise: drop the numerical null space of U0 (its singular vectors carry no signal)
    s_u, s_v, s_w = np.linalg.svd(U0)
    keep = s_v > s_v[0] * 1e-10
    Pl, Pr = s_u[:, keep].conj().T, s_w[keep].conj().T
    r0, r1, r2 = Pl @ U0 @ Pr, Pl @ U1 @ Pr, Pl @ U2 @ Pr
    u, Bk = eig(r1, r0)
    B = Pr @ Bk                                        # back to the Fourier basis: U1 B = u U0 B
    # bilinear (not Hermitian) normalisation B_k^T U0 B_k = 1
    nrm = np.einsum('jk,jl,lk->k', B, U0, B)
    B = B / np.sqrt(nrm)[None, :]
    Cj = (1. / z)[:, None] ** np.arange(0, M + 1)[None, :] @ sig[:M + 1]
    d = (B.T @ Cj) ** 2                                # complex amplitudes d_k
    u2 = np.einsum('jk,jl,lk->k', B, U2, B)            # B_k^T U2 B_k = u_k^2 for an exact mode
    with np.errstate(divide='ignore', invalid='ignore'):
        err = np.abs(np.log(u2 / (u * u))) / np.maximum(np.abs(np.log(u)), 1e-300)
    freq = -np.angle(u) / (2 * np.pi * dt)
    decay = -np.log(np.abs(u)) / dt
    amp = np.abs(d)
    phase = np.angle(d)
    ok = np.isfinite(freq) & np.isfinite(decay) & np.isfinite(err) & (freq >= fmin) & (freq <= fmax)
    ok &= decay > -1e-6 * np.abs(2 * np.pi * freq) - 1e-12 / dt
    ok &= err <= error_cut
    if ok.any():
        ok &= amp >= amp_rel_cut * amp[ok].max()
    idx = np.where(ok)[0]
    idx = idx[np.argsort(freq[idx])]
    return HarminvResult(freq[idx], np.maximum(decay[idx], 0.), amp[idx], phase[idx], err[idx])


def fft_peaks(signal, dt, fmin=None, fmax=None, npeaks=10):
    """Peak table of the plain FFT spectrum (the analysis of analyzer.use_fft, analyzer.py:77-166):
    local maxima of |FFT| refined by a parabola through the three bins around each maximum.
    Returns an array of rows (freq, |amplitude|)."""
    sig = np.asarray(signal, dtype=np.complex128).ravel()
    # harminv's convention: a mode is exp(-i 2 pi f t), i.e. it peaks at +f in the e^{+i} transform
    spec = np.abs(np.fft.ifft(sig)) * sig.size
    fr = np.fft.fftfreq(sig.size, dt)
    order = np.argsort(fr)
    fr, spec = fr[order], spec[order]
    lo = fr[0] if fmin is None else fmin
    hi = fr[-1] if fmax is None else fmax
    rows = []
    for i in range(1, fr.size - 1):
        if lo <= fr[i] <= hi and spec[i] > spec[i - 1] and spec[i] >= spec[i + 1]:
            y0, y1, y2 = spec[i - 1], spec[i], spec[i + 1]
            den = y0 - 2 * y1 + y2
            off = 0.5 * (y0 - y2) / den if den != 0 else 0.
            rows.append((fr[i] + off * (fr[1] - fr[0]), y1 - 0.25 * (y0 - y2) * off))
    rows.sort(key=lambda r: -r[1])
    return np.array(rows[:npeaks]).reshape(-1, 2)


class SpectrumAnalyzer:
    """analyzer.py:8-246: load the six `<name>_<F>_t.npy` signals a FieldAtPoint saved
    (collector.py:203-261) and analyse them by FFT or harmonic inversion."""

    def __init__(self, loaddir, savedir, name, **kwargs):
        self.cname = name
        self.savedir = savedir
        self.loaddir = loaddir
        binary = kwargs.get('binary', True)
        for f in _FIELDS:
            base = os.path.join(self.loaddir, "{}_{}_t".format(name, f))
            if binary:
                sig = np.load(base + ".npy")
            else:
                raw = np.loadtxt(base + ".txt", dtype=str)
                sig = np.array([complex(s.replace('+-', '-').replace('i', 'j')) for s in np.atleast_1d(raw)])
            setattr(self, f + '_t', sig)

    def normalized_freq(self, freqs, lattice_constant):
        self.lc = lattice_constant
        return freqs * self.lc / c

    def use_fft(self, dt, lc, **kwargs):
        """analyzer.py:77-166: FFT of the six signals; optional .npy / .txt / .csv dumps with the
        reference's file names and the csv columns Nfreqs, freqs, Ex_w .. Hz_w."""
        self.dt = dt
        os.makedirs(self.savedir, exist_ok=True)
        for f in _FIELDS:
            setattr(self, f + '_w', np.fft.fft(getattr(self, f + '_t')))
            if kwargs.get('binary'):
                np.save("{}/{}_{}_w_fft.npy".format(self.savedir, self.cname, f), getattr(self, f + '_w'))
            if kwargs.get('txt'):
                w = getattr(self, f + '_w')
                np.savetxt("{}/{}_{}_w_fft.txt".format(self.savedir, self.cname, f),
                           np.column_stack([w.real, w.imag]), newline='\n', fmt='%1.15f+%1.15fi')
        if kwargs.get('csv'):
            import pandas as pd
            fftfreq = np.fft.fftfreq(len(self.Ex_t), self.dt)
            df = pd.DataFrame()
            df['Nfreqs'] = self.normalized_freq(fftfreq, lc)
            df['freqs'] = fftfreq
            for f in _FIELDS:
                df[f + '_w'] = abs(getattr(self, f + '_w'))
            df.to_csv("{}/{}_fft_results.csv".format(self.savedir, self.cname))

    def use_pharminv(self, name, dt, fmin, fmax, spacing, **kwargs):
        """analyzer.py:170-246: harmonic inversion of one component; prints the reference's table when
        printing=True and returns the result object (freq, decay, Q, amplitude, phase, error)."""
        signal = getattr(self, name + '_t')
        nf = kwargs.get('nf', 10)
        harm = harminv_fdm(signal, fmin, fmax, dt, nf=nf)
        if kwargs.get('printing') and len(harm):
            scale, funit, wunit, wscale = 1., 'Hz', 'm', 1.
            for lim, fu, wu, ws in ((1e3, 'KHz', 'km', 1e3), (1e6, 'MHz', 'm', 1e0), (1e9, 'GHz', 'mm', 1e-3),
                                    (1e12, 'THz', 'um', 1e-6), (1e15, 'PHz', 'nm', 1e-9)):
                if harm.freq[-1] > lim:
                    scale, funit, wunit, wscale = lim, fu, wu, ws
            nfreqs = self.normalized_freq(harm.freq, spacing)
            print(name, ':')
            for i in range(len(harm)):
                print("NFreq: {:+7.4f}, Freq: {:+5.3e}{:>4s}, WL: {:+5.3e}{:>3s}, Q: {:+5.3e}, Amp: {:+5.3e}, "
                      "Decay: {:+5.3e}, Phase: {:+5.3e}, Err: {:+5.3e}".format(
                          nfreqs[i], harm.freq[i] / scale, funit, c / harm.freq[i] / wscale, wunit, harm.Q[i],
                          harm.amplitude[i], harm.decay[i], harm.phase[i], harm.error[i]))
        return harm

    def band_table(self, name, dt, fmin, fmax, spacing, nf=10, csv=None):
        """One row per resonance: normalised frequency a/lambda, frequency, decay, Q, amplitude, phase,
        error -- what the reference's band-structure scripts collect per k-point.  csv: file name to write."""
        harm = self.use_pharminv(name, dt, fmin, fmax, spacing, nf=nf)
        tab = np.column_stack([self.normalized_freq(harm.freq, spacing), harm.table()]) if len(harm) else np.zeros((0, 7))
        if csv:
            os.makedirs(self.savedir, exist_ok=True)
            np.savetxt(os.path.join(self.savedir, csv), tab, delimiter=',',
                       header='Nfreq,freq,decay,Q,amplitude,phase,error', comments='')
        return tab
