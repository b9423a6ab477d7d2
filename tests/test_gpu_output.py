"""the sampling scripts' output tail (crop, .npy, PSD) on the device vs the CPU oracle (oracle/psd.py)."""
import numpy as np
import pytest
import torch

from oracle import psd as op

pytestmark = pytest.mark.gpu


def _signals(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L) / 100.0
    x = torch.randn(B, 1, L, generator=g) * 0.3 + 0.7            # DC offset: exercises remove_dc
    x += 2.0 * torch.sin(2 * torch.pi * 10.0 * t) + 0.5 * torch.sin(2 * torch.pi * 3.3 * t + 1.0)
    return x


@pytest.mark.parametrize("method", ["multitaper", "welch"])
@pytest.mark.parametrize("B,L,crop", [(5, 3072, 36), (3, 3000, 0), (2, 1001, 0), (130, 3072, 36)])
def test_psd_matches_oracle(built_lib, cuda_device, method, B, L, crop):
    import eegldm
    x = _signals(B, L, L + B)
    xs = x[:, :, crop:L - crop] if crop else x
    fn = op.psd_array_multitaper if method == "multitaper" else op.psd_array_welch
    ref, fref = fn(xs.numpy(), 100.0, fmin=0.0, fmax=18.0)
    psd, freqs = eegldm.compute_psd(x.to(cuda_device), sfreq=100.0, fmax=18.0, method=method, crop=crop)
    np.testing.assert_allclose(freqs, fref, rtol=1e-6)
    assert psd.shape == (B, 1, len(fref))
    # fp32 FFT of N = 3000 against fp64: relative to each window's spectral peak
    scale = ref.max(axis=-1, keepdims=True)
    np.testing.assert_allclose(psd.cpu().numpy() / scale, ref / scale, atol=2e-6, rtol=2e-4)
    db, _ = eegldm.compute_psd(x.to(cuda_device), sfreq=100.0, fmax=18.0, method=method, crop=crop, db=True)
    np.testing.assert_allclose(db.cpu().numpy(), 10 * np.log10(psd.cpu().numpy()), atol=1e-4)


def test_psd_options(built_lib, cuda_device):
    import eegldm
    x = _signals(3, 2000, 0)
    xd = x.to(cuda_device)
    for kw in (dict(normalization="full"), dict(low_bias=False), dict(bandwidth=1.0), dict(remove_dc=False), dict(fmin=2.0, fmax=30.0)):
        ref, fref = op.psd_array_multitaper(x.numpy(), 100.0, **{**dict(fmin=0.0, fmax=np.inf), **kw})
        psd, freqs = eegldm.compute_psd(xd, sfreq=100.0, **kw)
        np.testing.assert_allclose(freqs, fref, rtol=1e-6)
        scale = ref.max(axis=-1, keepdims=True)
        np.testing.assert_allclose(psd.cpu().numpy() / scale, ref / scale, atol=2e-6, rtol=2e-4)
    for kw in (dict(n_fft=512, n_overlap=128), dict(n_fft=2000), dict(n_fft=100, n_overlap=50, remove_dc=False)):
        ref, fref = op.psd_array_welch(x.numpy(), 100.0, **kw)
        psd, freqs = eegldm.compute_psd(xd, sfreq=100.0, method="welch", **kw)
        np.testing.assert_allclose(freqs, fref, rtol=1e-6)
        scale = ref.max(axis=-1, keepdims=True)
        np.testing.assert_allclose(psd.cpu().numpy() / scale, ref / scale, atol=2e-6, rtol=2e-4)
    with pytest.raises(eegldm.EegldmError):
        eegldm.compute_psd(xd, method="welch", n_fft=4096)        # mne raises too: n_fft > n_times
    # multi-channel windows [B, C, L]: every (window, channel) row is its own signal
    x2 = torch.cat([x, 2 * x], dim=1)
    psd2, _ = eegldm.compute_psd(x2.to(cuda_device), sfreq=100.0, fmax=18.0)
    ref, _ = op.psd_array_multitaper(x2.numpy(), 100.0, fmax=18.0)
    np.testing.assert_allclose(psd2.cpu().numpy() / ref.max(), ref / ref.max(), atol=2e-6, rtol=2e-4)


def test_sample_tail_end_to_end(built_lib, cuda_device, tmp_path):
    """sample_trials.py:169-197 on a batch: cropped windows bit-exact, per-window .npy files, PSD in dB vs the oracle."""
    import eegldm
    s = _signals(6, 3072, 1)
    cropped, db, freqs, mean = eegldm.sample_tail(s.to(cuda_device), tmp_path, first_index=3, legacy_psd_files=True)
    rc, rdb, rf, rmean = op.sample_tail(s.numpy())
    assert torch.equal(cropped, s[:, :, 36:-36])
    np.testing.assert_allclose(freqs, rf, rtol=1e-6)
    np.testing.assert_allclose(db.numpy(), rdb, atol=2e-3)        # dB of fp32 vs fp64 spectra
    np.testing.assert_allclose(mean.numpy(), rmean, atol=2e-3)
    for i in range(6):
        np.testing.assert_array_equal(np.load(tmp_path / f"sample_{3 + i}.npy"), rc[i:i + 1].astype(np.float32))
        info = np.load(tmp_path / f"psd_list_{3 + i}.npy", allow_pickle=True)   # [psds, freqs, psds_mean], sample_trials.py:188-190
        np.testing.assert_allclose(info[0], rdb[i], atol=2e-3)
        np.testing.assert_allclose(info[2], rmean[i], atol=2e-3)
    np.testing.assert_allclose(np.load(tmp_path / "psd.npy"), rdb, atol=2e-3)
    assert np.load(tmp_path / "psd_list.npy", allow_pickle=True).shape[0] == 6
    # empty batch
    p, f = eegldm.compute_psd(torch.zeros(0, 1, 3072, device=cuda_device), fmax=18.0, crop=36)
    assert p.shape == (0, 1, 541)
