"""eegldm -- B200-native drop-in for the reference's latent-diffusion sampling path.

Re-exposes, on top of the C ABI in ``include/eegldm.h`` (hand-written sm_100a CUDA kernels):

* ``UNetModel``            <- ``src/models/unet.py:330``  (``forward(x, timesteps)``)
* ``AutoencoderKL``        <- ``generative.networks.nets.AutoencoderKL`` as built at ``src/sample_trials.py:95-100``
* ``DDIMScheduler`` / ``DDPMScheduler`` <- ``generative.networks.schedulers`` as used at ``src/sample_trials.py:136-163``
* ``JukeboxLoss``          <- ``generative.losses.JukeboxLoss`` as used at ``src/train_autoencoderkl.py:158,208``
* ``AutoencoderKL.train_step`` <- the generator half of the training step ``src/train_autoencoderkl.py:204-220``
* ``ddim_sample``          <- the sampling loop ``src/sample_trials.py:153-169`` as one fused call
* ``sample_tail`` / ``compute_psd`` <- the output tail ``src/sample_trials.py:169-197`` (crop, ``.npy``, MNE ``compute_psd``), batched

PyTorch is used for device memory, streams and ``nn.Module`` plumbing only.
"""
from ._lib import EegldmError, lib, LIB_PATH  # noqa: F401
from .unet import UNetModel  # noqa: F401
from .aekl import AutoencoderKL  # noqa: F401
from .schedulers import DDIMScheduler, DDPMScheduler  # noqa: F401
from .losses import JukeboxLoss  # noqa: F401
from .discriminator import PatchDiscriminator, PatchAdversarialLoss  # noqa: F401
from .sampler import ddim_sample, ddim_sample_host, shard_range, sample_sharded  # noqa: F401
from .output import compute_psd, crop_to_host, save_npy, save_windows, sample_tail, psd_freqs  # noqa: F401
from . import synthetic  # noqa: F401


def launch_count() -> int:
    """CUDA kernels launched by libeegldm since process start."""
    return int(lib().eegldm_launch_count())


def version() -> str:
    return lib().eegldm_version().decode()
