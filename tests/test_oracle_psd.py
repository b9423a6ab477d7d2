"""The output-tail oracle (oracle/psd.py: restated mne psd_array_multitaper / psd_array_welch, parity unpinned -- mne is not
installable here) against known answers, and the host-side pieces of the product's tail (DPSS, .npy writer, frequency grids)
against scipy / numpy.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest
from scipy import signal
from scipy.signal.windows import dpss

from oracle import psd as op


def test_multitaper_white_noise_level_and_parseval():
    """Unit-norm tapers: E|X_k(f)|^2 = sigma^2, so the 'length'-normalised one-sided PSD of white noise is 2 sigma^2 at every
    interior bin (sigma^2 at DC / Nyquist); 'full' divides by sfreq."""
    rng = np.random.default_rng(0)
    x = rng.normal(0.0, 3.0, size=(400, 1000))
    psd, freqs = op.psd_array_multitaper(x, 100.0, remove_dc=False)
    assert freqs.shape == (501,) and freqs[0] == 0 and freqs[-1] == 50.0
    m = psd.mean(axis=0)
    np.testing.assert_allclose(m[5:-5].mean(), 2 * 9.0, rtol=0.01)
    np.testing.assert_allclose(m[0], 9.0, rtol=0.15)
    np.testing.assert_allclose(m[-1], 9.0, rtol=0.15)
    full, _ = op.psd_array_multitaper(x, 100.0, remove_dc=False, normalization="full")
    np.testing.assert_allclose(full, psd / 100.0)


def test_multitaper_tapers_and_tone():
    tapers, eig = op.dpss_windows(3000, 4.0, True)
    assert tapers.shape == (7, 3000)                    # 2*NW - 1 tapers pass low_bias at NW = 4
    assert (eig > 0.9).all()
    np.testing.assert_allclose((tapers ** 2).sum(axis=1), 1.0, atol=1e-3)   # unit L2 norm (sym=False drops one tiny sample)
    # a bin-centred tone of amplitude A: the energy A^2 N / 4 ... spread over the 2W band; total one-sided power is conserved
    n, fs, f0, A = 3000, 100.0, 10.0, 2.0
    t = np.arange(n) / fs
    x = A * np.sin(2 * np.pi * f0 * t)
    psd, freqs = op.psd_array_multitaper(x[None], fs, fmax=18.0)
    assert freqs[-1] == 18.0 and len(freqs) == 541      # sample_trials.py:174 grid: 0 .. 18 Hz in steps of 1/30 Hz
    band = (freqs > f0 - 0.2) & (freqs < f0 + 0.2)
    np.testing.assert_allclose(psd[0].sum(), A * A * n / 2.0, rtol=1e-3)         # sum_f psd = 2 * sum_f |X|^2 = A^2 N / 2 (Parseval)
    np.testing.assert_allclose(psd[0, band].sum(), A * A * n / 2.0, rtol=5e-3)   # ... concentrated in the 2W = 0.27 Hz band
    assert psd[0, ~band].max() < 1e-2 * psd[0, band].max()


def test_welch_matches_scipy_welch():
    """psd_array_welch with mne's defaults is scipy.signal.welch(window='hamming', nperseg=256, noverlap=0,
    detrend='constant') -- same estimator through spectrogram()."""
    rng = np.random.default_rng(1)
    x = rng.normal(size=(3, 3000))
    psd, freqs = op.psd_array_welch(x, 100.0, fmax=18.0)
    f2, p2 = signal.welch(x, fs=100.0, window="hamming", nperseg=256, noverlap=0, nfft=256, detrend="constant")
    mask = f2 <= 18.0
    np.testing.assert_allclose(freqs, f2[mask])
    np.testing.assert_allclose(psd, p2[:, mask], rtol=1e-10)
    with pytest.raises(ValueError):
        op.psd_array_welch(x[:, :100], 100.0)


def test_sample_tail_shapes():
    rng = np.random.default_rng(2)
    s = rng.normal(size=(4, 1, 3072))
    cropped, db, freqs, mean = op.sample_tail(s)
    assert cropped.shape == (4, 1, 3000) and db.shape == (4, 1, 541) and mean.shape == (4, 541)
    np.testing.assert_array_equal(cropped, s[:, :, 36:-36])
    np.testing.assert_allclose(mean, db[:, 0])


# ---- host-side pieces of the product's tail (C ABI, no GPU) --------------------------------------------------------------
@pytest.mark.parametrize("N,NW,K,sym", [(3000, 4.0, 8, 0), (3000, 4.0, 8, 1), (256, 2.5, 4, 1), (1000, 3.0, 6, 0), (65, 2.0, 3, 0)])
def test_native_dpss_matches_scipy(built_lib, N, NW, K, sym):
    w = np.empty((K, N))
    r = np.empty(K)
    assert built_lib.eegldm_dpss(N, NW, K, sym, w.ctypes.data_as(C.POINTER(C.c_double)), r.ctypes.data_as(C.POINTER(C.c_double))) == 0
    ws, rs = dpss(N, NW, K, sym=bool(sym), norm=2, return_ratios=True)
    np.testing.assert_allclose(w, ws, atol=1e-10)
    np.testing.assert_allclose(r, rs, atol=1e-12)


def test_native_dpss_rejects_bad_arguments(built_lib):
    w = np.empty((2, 16))
    r = np.empty(2)
    wp, rp = w.ctypes.data_as(C.POINTER(C.c_double)), r.ctypes.data_as(C.POINTER(C.c_double))
    assert built_lib.eegldm_dpss(16, 8.0, 2, 1, wp, rp) != 0      # NW >= N/2
    assert built_lib.eegldm_dpss(16, 2.0, 0, 1, wp, rp) != 0      # Kmax < 1
    assert built_lib.eegldm_dpss(16, 2.0, 2, 1, None, rp) != 0


def test_native_npy_writer_round_trips(built_lib, tmp_path):
    import eegldm
    rng = np.random.default_rng(3)
    for shape in [(1, 1, 3000), (5,), (2, 3), (0, 4), ()]:
        a = rng.normal(size=shape).astype(np.float32)
        p = tmp_path / f"a{len(shape)}.npy"
        eegldm.save_npy(p, a)
        b = np.load(p)
        assert b.dtype == np.float32 and b.shape == a.shape
        np.testing.assert_array_equal(a, b)
    w = rng.normal(size=(3, 1, 3000)).astype(np.float32)
    eegldm.save_windows(tmp_path / "out", w, first_index=7)
    for i in range(3):
        b = np.load(tmp_path / "out" / f"sample_{7 + i}.npy")
        assert b.shape == (1, 1, 3000)                             # what sample_trials.py:170 writes per seed
        np.testing.assert_array_equal(b[0], w[i])
    assert not os.path.exists(tmp_path / "out" / "sample_10.npy")


def test_native_frequency_grids(built_lib):
    import eegldm
    f = eegldm.psd_freqs(3000, 100.0, 0.0, 18.0)
    ref = np.fft.rfftfreq(3000, 0.01)
    np.testing.assert_allclose(f, ref[ref <= 18.0], rtol=1e-6)
    assert len(f) == 541
    f = eegldm.psd_freqs(3000, 100.0, 0.5, 12.0, method="welch")
    ref = np.fft.rfftfreq(256, 0.01)
    np.testing.assert_allclose(f, ref[(ref >= 0.5) & (ref <= 12.0)], rtol=1e-6)
    assert len(eegldm.psd_freqs(3001, 100.0)) == 1501
