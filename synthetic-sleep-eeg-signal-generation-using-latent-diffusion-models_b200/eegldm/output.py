"""Output tail of the sampling scripts (``src/sample_trials.py:169-197``), batched: crop ``[36:-36]``, ``.npy`` files and the
power spectral density that the reference computes per window on the host through MNE
(``mne.EpochsArray(...).compute_psd(fmax=18)``, ``util.py:66-89``) -- here for ALL windows in one pass on the device
(``eegldm_psd``: one window kernel, one batched cuFFT call, one reduction kernel; ``eegldm_crop_to_host``;
``eegldm_save_windows_npy``)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._module import check_cuda_f32

_METHODS = {"multitaper": 0, "welch": 1}


def _psd_cfg(sfreq, fmin, fmax, method, bandwidth, low_bias, normalization, n_fft, n_overlap, remove_dc, db):
    if method not in _METHODS:
        raise ValueError("method must be 'multitaper' (mne Epochs.compute_psd default) or 'welch'")
    if normalization not in ("length", "full"):
        raise ValueError("normalization must be 'length' or 'full'")
    cfg = _lib.PsdCfg()
    cfg.method, cfg.sfreq, cfg.fmin = _METHODS[method], float(sfreq), float(fmin)
    cfg.fmax = float(min(fmax, 3.0e38))
    cfg.bandwidth = float(bandwidth) if bandwidth is not None else 0.0
    cfg.low_bias, cfg.normalization = int(bool(low_bias)), int(normalization == "full")
    cfg.n_fft, cfg.n_overlap = int(n_fft), int(n_overlap)
    cfg.remove_dc, cfg.db = int(bool(remove_dc)), int(bool(db))
    return cfg


def psd_freqs(n_times, sfreq=100.0, fmin=0.0, fmax=float("inf"), method="multitaper", n_fft=256, n_overlap=0):
    cfg = _psd_cfg(sfreq, fmin, fmax, method, None, True, "length", n_fft, n_overlap, True, False)
    n = C.c_int(0)
    _lib.check(_lib.lib().eegldm_psd_freqs(C.byref(cfg), int(n_times), C.byref(n), None))
    f = np.empty(n.value, dtype=np.float32)
    _lib.check(_lib.lib().eegldm_psd_freqs(C.byref(cfg), int(n_times), C.byref(n), f.ctypes.data_as(C.POINTER(C.c_float))))
    return f


@torch.no_grad()
def compute_psd(x, sfreq=100.0, fmin=0.0, fmax=float("inf"), method="multitaper", bandwidth=None, low_bias=True,
                normalization="length", n_fft=256, n_overlap=0, remove_dc=True, db=False, crop=0):
    """PSD of every (window, channel) row of ``x`` [B, C, L] on the device -> (psds [B, C, F] CUDA tensor, freqs [F] numpy).
    ``crop`` drops that many samples at both ends first (an offset pointer, no copy).  Defaults follow
    ``mne.Epochs.compute_psd()``; the reference calls it with ``fmax=18`` (sample_trials.py:174) on 100 Hz data."""
    x = check_cuda_f32(x, "x")
    if x.dim() != 3:
        raise ValueError("x must be [B, C, L]")
    B, Cc, Lx = x.shape
    N = Lx - 2 * crop
    if N < 2:
        raise ValueError("crop leaves no signal")
    cfg = _psd_cfg(sfreq, fmin, fmax, method, bandwidth, low_bias, normalization, n_fft, n_overlap, remove_dc, db)
    freqs = psd_freqs(N, sfreq, fmin, fmax, method, n_fft, n_overlap)
    out = torch.empty((B, Cc, len(freqs)), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().eegldm_psd(C.byref(cfg), C.c_void_p(x.data_ptr() + 4 * crop), int(B * Cc), int(N), int(Lx),
                                         C.c_void_p(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr(x.device))))
    return out, freqs


@torch.no_grad()
def crop_to_host(x, crop=36, out_host=None):
    """``x.cpu().numpy()[:, :, crop:-crop]`` (sample_trials.py:169) as ONE strided device->host copy into (pinned) host memory."""
    x = check_cuda_f32(x, "x")
    B, Cc, Lx = x.shape
    N = Lx - 2 * crop
    if out_host is None:
        out_host = torch.empty((B, Cc, N), dtype=torch.float32, pin_memory=True)
    elif out_host.is_cuda or out_host.dtype != torch.float32 or not out_host.is_contiguous() or tuple(out_host.shape) != (B, Cc, N):
        raise ValueError(f"out_host must be a contiguous fp32 CPU tensor of shape {(B, Cc, N)}")
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().eegldm_crop_to_host(C.c_void_p(x.data_ptr()), int(B * Cc), int(Lx), int(crop), int(crop),
                                                  C.c_void_p(out_host.data_ptr()), C.c_void_p(_lib.current_stream_ptr(x.device))))
    return out_host


def save_npy(path, array_host):
    """``np.save(path, array)`` for a contiguous fp32 host tensor / array through the native writer."""
    a = array_host.numpy() if isinstance(array_host, torch.Tensor) else np.asarray(array_host)
    a = np.require(a, dtype=np.float32, requirements="C")   # (ascontiguousarray would promote a 0-d array to 1-d)
    shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
    _lib.check(_lib.lib().eegldm_write_npy_f32(os.fspath(path).encode(), C.c_void_p(a.ctypes.data), shape, a.ndim))


def save_windows(output_dir, windows_host, first_index=0, prefix="sample_"):
    """One ``sample_{i}.npy`` of shape [1, C, L] per window, i = first_index ... (sample_trials.py:170)."""
    w = windows_host if isinstance(windows_host, torch.Tensor) else torch.from_numpy(np.asarray(windows_host))
    if w.is_cuda or w.dtype != torch.float32 or not w.is_contiguous() or w.dim() != 3:
        raise ValueError("windows_host must be a contiguous fp32 CPU tensor [B, C, L]")
    B, Cc, Lw = w.shape
    os.makedirs(output_dir, exist_ok=True)
    _lib.check(_lib.lib().eegldm_save_windows_npy(os.fspath(output_dir).encode(), prefix.encode(), int(first_index),
                                                  C.c_void_p(w.data_ptr()), int(B), int(Cc), int(Lw)))


@torch.no_grad()
def sample_tail(sample, output_dir=None, first_index=0, crop=36, sfreq=100.0, fmax=18.0, method="multitaper",
                legacy_psd_files=False):
    """sample_trials.py:169-197 for a whole batch of decoded windows ``sample`` [B, C, L] (CUDA):
    crop, ``sample_{i}.npy`` per window, PSD in dB (``10 * log10``) and its channel mean.
    Returns ``(cropped_host [B, C, L-2*crop], psds_db [B, C, F] (host), freqs [F], psds_mean [B, F])``; with ``output_dir`` it
    also writes ``psd.npy`` / ``freqs.npy`` / ``psd_mean.npy`` for the batch and, with ``legacy_psd_files``, the reference's
    pickled per-window ``psd_list_{i}.npy`` ([psds, freqs, psds_mean]) and ``psd_list.npy``."""
    psds, freqs = compute_psd(sample, sfreq=sfreq, fmin=0.0, fmax=fmax, method=method, db=True, crop=crop)
    cropped = crop_to_host(sample, crop)
    psds_h = psds.cpu()
    mean = psds_h.mean(dim=1)
    if output_dir is not None:
        save_windows(output_dir, cropped, first_index)
        save_npy(os.path.join(output_dir, "psd.npy"), psds_h)
        save_npy(os.path.join(output_dir, "freqs.npy"), freqs)
        save_npy(os.path.join(output_dir, "psd_mean.npy"), mean)
        if legacy_psd_files:
            allinfo = []
            for i in range(psds_h.shape[0]):
                info = np.empty(3, dtype=object)
                info[0], info[1], info[2] = psds_h[i].numpy(), freqs, mean[i].numpy()
                allinfo.append(info)
                np.save(os.path.join(output_dir, f"psd_list_{first_index + i}.npy"), info, allow_pickle=True)
            np.save(os.path.join(output_dir, "psd_list.npy"), np.array(allinfo, dtype=object), allow_pickle=True)
    return cropped, psds_h, freqs, mean
