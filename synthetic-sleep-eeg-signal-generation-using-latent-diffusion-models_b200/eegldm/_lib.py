"""ctypes binding of the eegldm C ABI (include/eegldm.h).

There is no CPU fallback: if ``libeegldm.so`` has not been built (``python <package>/build.py`` or
``__graft_entry__.build()``) importing this module raises, and every compute entry point returns
``EEGLDM_ERR_CUDA`` without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeegldm.so")


class EegldmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"eegldm error {code}: {msg}")
        self.code = code


class UNetCfg(C.Structure):
    _fields_ = [
        ("image_size", C.c_int32), ("in_channels", C.c_int32), ("model_channels", C.c_int32),
        ("out_channels", C.c_int32), ("num_res_blocks", C.c_int32),
        ("n_attention_resolutions", C.c_int32), ("attention_resolutions", C.c_int32 * 8),
        ("n_channel_mult", C.c_int32), ("channel_mult", C.c_int32 * 8),
        ("num_heads", C.c_int32), ("num_head_channels", C.c_int32), ("num_heads_upsample", C.c_int32),
        ("resblock_updown", C.c_int32), ("conv_resample", C.c_int32), ("use_scale_shift_norm", C.c_int32),
    ]


class AeklCfg(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("n_levels", C.c_int32),
        ("num_channels", C.c_int32 * 8), ("num_res_blocks", C.c_int32 * 8),
        ("latent_channels", C.c_int32), ("norm_num_groups", C.c_int32),
    ]


class AeklTrainCfg(C.Structure):
    _fields_ = [("kl_weight", C.c_float), ("spectral_weight", C.c_float), ("lr", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("adam_eps", C.c_float)]


class DiscCfg(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("num_channels", C.c_int32), ("num_layers_d", C.c_int32),
                ("kernel_size", C.c_int32), ("padding", C.c_int32)]


class AeklAdvTrainCfg(C.Structure):
    _fields_ = [("kl_weight", C.c_float), ("spectral_weight", C.c_float), ("adv_weight", C.c_float), ("lr_g", C.c_float),
                ("lr_d", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float),
                ("no_activation_leastsq", C.c_int32)]


class PsdCfg(C.Structure):
    _fields_ = [("method", C.c_int32), ("sfreq", C.c_float), ("fmin", C.c_float), ("fmax", C.c_float), ("bandwidth", C.c_float),
                ("low_bias", C.c_int32), ("normalization", C.c_int32), ("n_fft", C.c_int32), ("n_overlap", C.c_int32),
                ("remove_dc", C.c_int32), ("db", C.c_int32)]


class LdmTrainCfg(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float)]


class SchedCfg(C.Structure):
    _fields_ = [
        ("num_train_timesteps", C.c_int32), ("beta_start", C.c_float), ("beta_end", C.c_float),
        ("schedule", C.c_int32), ("prediction_type", C.c_int32), ("set_alpha_to_one", C.c_int32),
        ("steps_offset", C.c_int32),
    ]


MATH_FP32_SIMT, MATH_F16X3_TC, MATH_BF16_TC = 0, 1, 2
MATH_MODES = {"fp32": MATH_FP32_SIMT, "fp32_simt": MATH_FP32_SIMT, "f16x3": MATH_F16X3_TC, "f16x3_tc": MATH_F16X3_TC,
              "bf16": MATH_BF16_TC, "bf16_tc": MATH_BF16_TC}

_P = C.c_void_p
_FP = C.POINTER(C.c_float)
_I64P = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol include/eegldm.h declares
SIGNATURES = {
    "eegldm_last_error": (C.c_char_p, []),
    "eegldm_version": (C.c_char_p, []),
    "eegldm_launch_count": (C.c_int64, []),
    "eegldm_set_graphs": (C.c_int, [C.c_int]),
    "eegldm_set_conv_cluster": (C.c_int, [C.c_int]),
    "eegldm_set_sample_lanes": (C.c_int, [C.c_int]),
    "eegldm_set_conv_tuning": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "eegldm_profile_enable": (C.c_int, [C.c_int]),
    "eegldm_profile_record": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "eegldm_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_int64)]),
    "eegldm_unet_create": (C.c_int, [C.POINTER(UNetCfg), C.POINTER(_P)]),
    "eegldm_unet_destroy": (None, [_P]),
    "eegldm_unet_num_params": (C.c_int, [_P]),
    "eegldm_unet_param_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), _I64P, C.POINTER(C.c_int)]),
    "eegldm_unet_load": (C.c_int, [_P, C.c_char_p, _P, _I64P, C.c_int]),
    "eegldm_unet_finalize": (C.c_int, [_P]),
    "eegldm_unet_set_math": (C.c_int, [_P, C.c_int]),
    "eegldm_unet_forward": (C.c_int, [_P, _P, _FP, C.c_int, _P, C.c_int, C.c_int, _P]),
    "eegldm_unet_forward_devt": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "eegldm_unet_range_status": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "eegldm_bench_conv_timeline": (C.c_int, [C.c_int] * 9 + [_P, C.POINTER(C.c_double), _P]),
    "eegldm_bench_attention": (C.c_int, [C.c_int] * 5 + [_P, C.POINTER(C.c_double), _P]),
    "eegldm_aekl_create": (C.c_int, [C.POINTER(AeklCfg), C.POINTER(_P)]),
    "eegldm_aekl_destroy": (None, [_P]),
    "eegldm_aekl_num_params": (C.c_int, [_P]),
    "eegldm_aekl_param_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), _I64P, C.POINTER(C.c_int)]),
    "eegldm_aekl_load": (C.c_int, [_P, C.c_char_p, _P, _I64P, C.c_int]),
    "eegldm_aekl_finalize": (C.c_int, [_P]),
    "eegldm_aekl_encode": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "eegldm_aekl_decode": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P]),
    "eegldm_aekl_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "eegldm_jukebox_loss": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "eegldm_aekl_train_step": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.POINTER(AeklTrainCfg), _FP, _P]),
    "eegldm_aekl_train_export": (C.c_int, [_P, C.c_int, C.c_char_p, _FP]),
    "eegldm_aekl_train_sync": (C.c_int, [_P]),
    "eegldm_aekl_forward_train": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "eegldm_aekl_backward": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "eegldm_disc_forward_train": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "eegldm_disc_backward": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P]),
    "eegldm_unet_train_step": (C.c_int, [_P, C.POINTER(SchedCfg), _P, _P, _P, C.c_int, C.c_int, C.POINTER(LdmTrainCfg), _FP, _P]),
    "eegldm_unet_train_export": (C.c_int, [_P, C.c_int, C.c_char_p, _FP]),
    "eegldm_unet_train_sync": (C.c_int, [_P]),
    "eegldm_unet_forward_train": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "eegldm_unet_backward": (C.c_int, [_P, _P, _P]),
    "eegldm_sched_alphas_cumprod": (C.c_int, [C.POINTER(SchedCfg), _FP]),
    "eegldm_sched_ddim_tables": (C.c_int, [C.POINTER(SchedCfg), C.c_int, _I64P, _FP]),
    "eegldm_timestep_embedding": (C.c_int, [_FP, C.c_int, C.c_int, _FP]),
    "eegldm_ddim_sample": (C.c_int, [_P, _P, C.POINTER(SchedCfg), _P, C.c_float, C.c_int, _P, C.c_int, C.c_int, _P]),
    "eegldm_disc_create": (C.c_int, [C.POINTER(DiscCfg), C.POINTER(_P)]),
    "eegldm_disc_destroy": (None, [_P]),
    "eegldm_disc_num_params": (C.c_int, [_P]),
    "eegldm_disc_param_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), _I64P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "eegldm_disc_load": (C.c_int, [_P, C.c_char_p, _P, _I64P, C.c_int]),
    "eegldm_disc_finalize": (C.c_int, [_P]),
    "eegldm_disc_forward": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "eegldm_disc_out_len": (C.c_int, [_P, C.c_int]),
    "eegldm_disc_set_math": (C.c_int, [_P, C.c_int]),
    "eegldm_disc_export": (C.c_int, [_P, C.c_int, C.c_char_p, _FP]),
    "eegldm_aekl_train_step_adv": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(AeklAdvTrainCfg), _FP, _P]),
    "eegldm_psd_freqs": (C.c_int, [C.POINTER(PsdCfg), C.c_int, C.POINTER(C.c_int), _FP]),
    "eegldm_psd": (C.c_int, [C.POINTER(PsdCfg), _P, C.c_int, C.c_int, C.c_int64, _P, _P]),
    "eegldm_dpss": (C.c_int, [C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "eegldm_crop_to_host": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, _P]),
    "eegldm_write_npy_f32": (C.c_int, [C.c_char_p, _P, _I64P, C.c_int]),
    "eegldm_save_windows_npy": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int64, _P, C.c_int, C.c_int, C.c_int]),
    "eegldm_test_conv_gn": (C.c_int, [_P, _P, _P] + [C.c_int] * 6 + [_P, _P, _P, _P]),
    "eegldm_test_qkv_attention": (C.c_int, [_P, _P, _P] + [C.c_int] * 4 + [_P, _P]),
    "eegldm_bench_conv": (C.c_int, [C.c_int] * 9 + [_P, _P]),
    "eegldm_test_conv": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "eegldm_test_attention": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "eegldm_ddim_sample_host": (C.c_int, [_P, _P, C.POINTER(SchedCfg), _P, C.c_float, C.c_int, _P, C.c_int, C.c_int, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libeegldm.so (once) and bind every symbol; raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA library has not been built. Run "
                f"`python {os.path.join(os.path.dirname(_HERE), 'build.py')}` (needs nvcc). "
                "eegldm has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise EegldmError(code, lib().eegldm_last_error().decode("utf-8", "replace"))


def current_stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def load_state_dict_into(handle, load_fn, state_dict) -> None:
    """One C-ABI load call per entry (host fp32, contiguous, reference [Cout,Cin,k] layout)."""
    import torch
    for name, t in state_dict.items():
        ht = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
        shape = (C.c_int64 * max(ht.dim(), 1))(*ht.shape)
        check(load_fn(handle, name.encode(), C.c_void_p(ht.data_ptr()), shape, ht.dim()))


def param_infos(handle, num_fn, info_fn):
    out = []
    for i in range(num_fn(handle)):
        name = C.c_char_p()
        shape = (C.c_int64 * 4)()
        nd = C.c_int()
        check(info_fn(handle, i, C.byref(name), shape, C.byref(nd)))
        out.append((name.value.decode(), tuple(int(shape[k]) for k in range(nd.value))))
    return out
