"""CPU: band-structure post-processing (ies_b200/analyzer.py).  The reference uses
the un-vendored third-party harminv (parity unpinned); the filter-diagonalisation routine is
checked against signals whose modes are known exactly."""
import os

import numpy as np
import pytest

from ies_b200 import analyzer as A


def _signal(modes, n, dt, noise=0., seed=0):
    t = np.arange(n) * dt
    s = np.zeros(n, dtype=np.complex128)
    for f, g, a, ph in modes:
        s += a * np.exp(-1j * (2 * np.pi * f * t - ph)) * np.exp(-g * t)
    if noise:
        rng = np.random.default_rng(seed)
        s += noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return s


MODES = [(1.10e12, 2.0e9, 1.0, 0.3), (1.23e12, 6.0e9, 0.5, -1.1), (1.26e12, 1.0e9, 0.2, 2.0)]


@pytest.mark.parametrize('noise', [0., 1e-6])
def test_fdm_recovers_known_modes(noise):
    dt = 2.0e-14
    s = _signal(MODES, 4000, dt, noise)
    h = A.harminv_fdm(s, 1.0e12, 1.4e12, dt, nf=40)
    assert len(h) >= 3
    for f, g, a, ph in MODES:
        k = int(np.argmin(np.abs(h.freq - f)))
        tol = 1e-9 if noise == 0 else 1e-6
        assert abs(h.freq[k] - f) / f < tol
        assert abs(h.decay[k] - g) / g < (1e-6 if noise == 0 else 5e-3)
        assert abs(h.amplitude[k] - a) / a < (1e-6 if noise == 0 else 5e-3)
        assert abs(np.angle(np.exp(1j * (h.phase[k] - ph)))) < (1e-6 if noise == 0 else 5e-3)
        assert abs(h.Q[k] - np.pi * f / g) / (np.pi * f / g) < (1e-6 if noise == 0 else 5e-3)
        assert h.error[k] < 1e-3


def test_fdm_real_signal_and_close_doublet():
    """A real-valued probe signal (real field dtype) and two modes 0.2 % apart -- closer than the FFT
    bin spacing of the record, the case filter diagonalisation exists for."""
    dt = 2.0e-14
    n = 3000
    modes = [(1.200e12, 3.0e9, 1.0, 0.), (1.2024e12, 3.0e9, 0.7, 0.5)]
    s = _signal(modes, n, dt).real
    assert 1. / (n * dt) > 2.4e9 * 2              # the doublet is not resolved by the FFT grid
    h = A.harminv_fdm(s, 1.1e12, 1.3e12, dt, nf=30)
    got = sorted(h.freq[np.argsort(-h.amplitude)[:2]])
    assert abs(got[0] - 1.200e12) / 1.2e12 < 1e-6 and abs(got[1] - 1.2024e12) / 1.2e12 < 1e-6


def test_fft_peaks_and_analyzer_files(tmp_path):
    dt = 2.0e-14
    s = _signal(MODES[:2], 8192, dt)
    pk = A.fft_peaks(s, dt, 0.9e12, 1.5e12, npeaks=2)
    assert abs(pk[0, 0] - 1.10e12) / 1.1e12 < 2e-3 and abs(pk[1, 0] - 1.23e12) / 1.23e12 < 2e-3
    # FieldAtPoint.save_time_signal layout -> SpectrumAnalyzer
    for f in ('Ex', 'Ey', 'Ez', 'Hx', 'Hy', 'Hz'):
        np.save(os.path.join(tmp_path, f'fap1_{f}_t.npy'), s if f == 'Ey' else 0.1 * s)
    an = A.SpectrumAnalyzer(str(tmp_path) + '/', str(tmp_path) + '/out', 'fap1')
    an.use_fft(dt, 512e-6, csv=True, binary=True)
    assert os.path.isfile(os.path.join(tmp_path, 'out', 'fap1_fft_results.csv'))
    assert os.path.isfile(os.path.join(tmp_path, 'out', 'fap1_Ey_w_fft.npy'))
    harm = an.use_pharminv('Ey', dt, 1.0e12, 1.4e12, 512e-6, nf=30)
    assert len(harm) >= 2 and abs(harm.freq[0] - 1.10e12) / 1.1e12 < 1e-8
    tab = an.band_table('Ey', dt, 1.0e12, 1.4e12, 512e-6, nf=30, csv='bands.csv')
    assert tab.shape[1] == 7 and os.path.isfile(os.path.join(tmp_path, 'out', 'bands.csv'))
    assert abs(tab[0, 0] - 1.10e12 * 512e-6 / 299792458.0) < 1e-6
