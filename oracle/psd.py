"""CPU restatement of the sampling scripts' output tail (TEST INFRASTRUCTURE).

Reference call sites: ``src/sample_trials.py:169-197`` (also ``sample_trials_ddpm.py:105-128``, ``util.py:66-112``)::

    cropped = sample.cpu().numpy()[:, :, 36:-36]
    np.save(output_dir / f"sample_{i}.npy", cropped)
    epochs = mne.EpochsArray(cropped, mne.create_info(1, ch_types=['eeg'], sfreq=100))     # util.py:66-89
    spectrum = epochs.compute_psd(fmax=18)
    psds, freqs = spectrum.average().get_data(return_freqs=True)
    psds = 10 * np.log10(psds);  psds_mean = psds.mean(axis=0)

The PSD estimator lives in a third-party dependency that is absent from ``/root/reference`` and not installable here:
``mne`` (``requirements.txt``, version unpinned).  **Parity unpinned** against mne itself; the functions below restate the
published algorithms of ``mne.time_frequency.psd_array_multitaper`` (the default ``method`` of ``Epochs.compute_psd`` since
mne 1.2) and ``mne.time_frequency.psd_array_welch`` (the default of ``Raw.compute_psd``; a thin wrapper over
``scipy.signal.spectrogram``) on top of the same scipy primitives mne calls (``scipy.signal.windows.dpss``,
``scipy.signal.spectrogram``), and are anchored by known answers in ``tests/test_oracle_psd.py``: white-noise level,
Parseval, a bin-centred tone, the DC / Nyquist halving, frequency grids.
"""
from __future__ import annotations

import numpy as np
from scipy import signal
from scipy.signal.windows import dpss as _sp_dpss


def dpss_windows(n_times, half_nbw=4.0, low_bias=True):
    """mne.time_frequency.multitaper.dpss_windows(N, half_nbw, Kmax=int(2*half_nbw), sym=False, low_bias=...) ->
    (tapers [K, N], eigvals [K]): scipy's DPSS with unit L2 norm, keeping the tapers with concentration > 0.9."""
    kmax = int(2 * half_nbw)
    tapers, eigvals = _sp_dpss(n_times, half_nbw, kmax, sym=False, norm=2, return_ratios=True)
    if low_bias:
        idx = eigvals > 0.9
        if not idx.any():
            idx = np.zeros_like(idx)
            idx[np.argmax(eigvals)] = True
        tapers, eigvals = tapers[idx], eigvals[idx]
    return tapers, eigvals


def psd_array_multitaper(x, sfreq, fmin=0.0, fmax=np.inf, bandwidth=None, low_bias=True, normalization="length",
                         remove_dc=True):
    """mne.time_frequency.psd_array_multitaper(x, sfreq, fmin, fmax, bandwidth=None, adaptive=False, low_bias=True,
    normalization='length', remove_dc=True) -> (psd [..., n_freqs], freqs)."""
    x = np.asarray(x, dtype=np.float64)
    n_times = x.shape[-1]
    half_nbw = float(bandwidth) * n_times / (2.0 * sfreq) if bandwidth is not None else 4.0   # _compute_mt_params
    tapers, eigvals = dpss_windows(n_times, half_nbw, low_bias)
    freqs = np.fft.rfftfreq(n_times, 1.0 / sfreq)
    mask = (freqs >= fmin) & (freqs <= fmax)
    if remove_dc:
        x = x - x.mean(axis=-1, keepdims=True)                    # _mt_spectra
    x_mt = np.fft.rfft(x[..., np.newaxis, :] * tapers, n=n_times)
    x_mt[..., 0] /= np.sqrt(2.0)                                    # "Adjust DC and maybe Nyquist" (one-sided transform)
    if n_times % 2 == 0:
        x_mt[..., -1] /= np.sqrt(2.0)
    weights = np.sqrt(eigvals)[:, np.newaxis]                       # adaptive=False
    psd = weights * x_mt[..., mask]                                 # _psd_from_mt
    psd = (psd * psd.conj()).real.sum(axis=-2)
    psd *= 2.0 / (weights * weights.conj()).real.sum(axis=-2)
    if normalization == "full":
        psd /= sfreq
    return psd, freqs[mask]


def psd_array_welch(x, sfreq, fmin=0.0, fmax=np.inf, n_fft=256, n_overlap=0, remove_dc=True):
    """mne.time_frequency.psd_array_welch(x, sfreq, fmin, fmax, n_fft=256, n_overlap=0, n_per_seg=None, average='mean',
    window='hamming', remove_dc=True): scipy.signal.spectrogram(mode='psd') per segment, mean over segments."""
    x = np.asarray(x, dtype=np.float64)
    if n_fft > x.shape[-1]:
        raise ValueError("n_fft is larger than the signal")
    freqs, _, spec = signal.spectrogram(x, fs=sfreq, window="hamming", nperseg=n_fft, noverlap=n_overlap, nfft=n_fft,
                                        detrend="constant" if remove_dc else False, scaling="density", mode="psd")
    psd = spec.mean(axis=-1)
    mask = (freqs >= fmin) & (freqs <= fmax)
    return psd[..., mask], freqs[mask]


def sample_tail(sample, crop=36, sfreq=100.0, fmax=18.0, method="multitaper"):
    """sample_trials.py:169-188 for one batch ``sample`` [B, C, L] (the reference runs it with B = 1 per seed, so its
    'average over epochs' is the identity): -> (cropped [B, C, L-2*crop], psds_db [B, C, F], freqs [F], psds_mean [B, F])."""
    cropped = np.asarray(sample)[:, :, crop:-crop] if crop else np.asarray(sample)
    fn = psd_array_multitaper if method == "multitaper" else psd_array_welch
    psds, freqs = fn(cropped, sfreq, fmin=0.0, fmax=fmax)
    psds_db = 10.0 * np.log10(psds)
    return cropped, psds_db, freqs, psds_db.mean(axis=1)
