"""Drop-in for the reference denoiser ``UNetModel`` (``/root/reference/src/models/unet.py:330-563``).

Same constructor keywords (``unet.py:331-351``), same ``state_dict`` keys, same
``forward(x, timesteps=None, context=None, y=None, **kwargs)`` contract (``unet.py:512``); the
computation is ``eegldm_unet_forward`` in ``libeegldm.so``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._module import EngineModule, check_cuda_f32, default_init, register_tree


def _zero_init(name: str) -> bool:
    """Tensors the reference wraps in ``zero_module`` (unet.py:161,290-292,504)."""
    return (".out_layers.3." in name) or (".proj_out." in name) or name.startswith("out.2.")


class _UNetTrainFn(torch.autograd.Function):
    """``UNetModel.forward`` across the autograd boundary (training.py:430): forward = ``eegldm_unet_forward_train`` (the pass stays
    in the engine handle), backward = ``eegldm_unet_backward`` + one gradient per parameter (they are Function inputs so autograd
    routes them to ``p.grad``).  The input latent gets no gradient (the reference computes it under ``no_grad``)."""

    @staticmethod
    def forward(ctx, mod, x, ts, *params):
        B, _, T = x.shape
        out = torch.empty((B, mod.out_channels, T), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            mod._sync_weights()
            _lib.check(_lib.lib().eegldm_unet_forward_train(mod._h, C.c_void_p(x.data_ptr()), C.c_void_p(ts.data_ptr()), C.c_void_p(out.data_ptr()),
                                                            int(B), int(T), C.c_void_p(_lib.current_stream_ptr(x.device))))
        mod._pass_token = getattr(mod, "_pass_token", 0) + 1
        ctx.mod, ctx.token, ctx.dev = mod, mod._pass_token, x.device
        return out

    @staticmethod
    def backward(ctx, d_out):
        mod = ctx.mod
        if ctx.token != mod._pass_token:
            raise RuntimeError("eegldm.UNetModel: backward through a forward pass that a later forward() has replaced "
                               "(the engine keeps one recorded pass per model)")
        d = d_out.contiguous().float()
        with torch.cuda.device(ctx.dev):
            _lib.check(_lib.lib().eegldm_unet_backward(mod._h, C.c_void_p(d.data_ptr()), C.c_void_p(_lib.current_stream_ptr(ctx.dev))))
        grads = mod._export(1)
        return (None, None, None) + tuple(grads[n].to(ctx.dev) for n, _ in mod.named_parameters())


class UNetModel(EngineModule):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, num_classes=None, num_heads=1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 n_embed=None, math="f16x3"):
        super().__init__()
        if num_classes is not None or n_embed is not None:
            raise NotImplementedError("class-conditional / codebook heads are not on the reference's live path")
        if use_scale_shift_norm:
            raise NotImplementedError("use_scale_shift_norm=True is not used by any reference config")
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = list(attention_resolutions)
        self.dropout = dropout  # Dropout is the identity in eval mode; this engine is inference-only
        self.channel_mult = tuple(channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = None
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.predict_codebook_ids = False

        cfg = _lib.UNetCfg()
        cfg.image_size = int(image_size)
        cfg.in_channels, cfg.model_channels, cfg.out_channels = int(in_channels), int(model_channels), int(out_channels)
        cfg.num_res_blocks = int(num_res_blocks)
        if len(self.attention_resolutions) > 8 or len(self.channel_mult) > 8:
            raise ValueError("at most 8 attention_resolutions / channel_mult entries")
        cfg.n_attention_resolutions = len(self.attention_resolutions)
        for i, v in enumerate(self.attention_resolutions):
            cfg.attention_resolutions[i] = int(v)
        cfg.n_channel_mult = len(self.channel_mult)
        for i, v in enumerate(self.channel_mult):
            cfg.channel_mult[i] = int(v)
        cfg.num_heads, cfg.num_head_channels = int(num_heads), int(num_head_channels)
        cfg.num_heads_upsample = int(num_heads_upsample)
        cfg.resblock_updown, cfg.conv_resample = int(bool(resblock_updown)), int(bool(conv_resample))
        cfg.use_scale_shift_norm = 0
        self._cfg = cfg
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.eegldm_unet_create(C.byref(cfg), C.byref(h)))
        self._h = h
        infos = _lib.param_infos(h, L.eegldm_unet_num_params, L.eegldm_unet_param_info)
        shapes = dict(infos)

        def init(name, shape):
            wname = name.rsplit(".", 1)[0] + ".weight"
            return default_init(name, shape, shapes[wname], zero=_zero_init(name))

        register_tree(self, infos, init)
        self._math = "fp32"
        self.set_math(math)   # default f16x3: the tensor-pipe parity mode (layers whose shapes are not eligible run fp32 SIMT)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().eegldm_unet_destroy(h)
            except Exception:
                pass
            object.__setattr__(self, "_h", None)   # nn.Module.__setattr__ may already be torn down at interpreter exit

    def set_math(self, mode: str) -> "UNetModel":
        """``"fp32"`` (SIMT, exact fp32), ``"f16x3"`` (tcgen05, fp16 hi + scaled fp16 lo, 3 products, ~fp32-accurate),
        ``"bf16"`` (tcgen05, fast; does NOT meet the fp32 parity tolerance)."""
        _lib.check(_lib.lib().eegldm_unet_set_math(self._h, _lib.MATH_MODES[mode]))
        self._math = mode
        return self

    def _upload(self, state_dict) -> None:
        L = _lib.lib()
        _lib.load_state_dict_into(self._h, L.eegldm_unet_load, state_dict)
        _lib.check(L.eegldm_unet_finalize(self._h))

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """``UNetModel.forward`` (unet.py:512-563).  In training mode under autograd (grad mode on, a parameter requires grad) the call
        is differentiable with respect to the parameters, so the reference's own loop runs unchanged (training.py:420-443);
        ``train_step`` is the fused form of the same step."""
        assert y is None, "must specify y if and only if the model is class-conditional"   # unet.py:521-523
        assert timesteps is not None, "need to implement no-timestep usage"               # unet.py:524
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            x = check_cuda_f32(x, "x").detach()
            if x.dim() != 3 or x.shape[1] != self.in_channels:
                raise ValueError(f"x must be [B, {self.in_channels}, T], got {tuple(x.shape)}")
            ts = torch.as_tensor(timesteps).reshape(-1).to(device=x.device, dtype=torch.float32)
            if ts.numel() == 1:
                ts = ts.expand(x.shape[0])
            if ts.numel() != x.shape[0]:
                raise ValueError("timesteps must have 1 or B entries")
            return _UNetTrainFn.apply(self, x, ts.contiguous(), *self.parameters())
        return self._forward_nograd(x, timesteps)

    @torch.no_grad()
    def _forward_nograd(self, x, timesteps):
        x = check_cuda_f32(x, "x")
        if x.dim() != 3 or x.shape[1] != self.in_channels:
            raise ValueError(f"x must be [B, {self.in_channels}, T], got {tuple(x.shape)}")
        B, _, T = x.shape
        ts = torch.as_tensor(timesteps).reshape(-1)
        if ts.numel() not in (1, B):
            raise ValueError("timesteps must have 1 or B entries")
        out = torch.empty((B, self.out_channels, T), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            self._sync_weights()
            L = _lib.lib()
            stream = C.c_void_p(_lib.current_stream_ptr(x.device))
            if ts.is_cuda:   # the reference's loops pass a CUDA tensor (training.py:430, sample_trials.py:157): no host sync
                ts = ts.to(device=x.device, dtype=torch.float32).contiguous()   # .float(): unet.py:28
                _lib.check(L.eegldm_unet_forward_devt(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(ts.data_ptr()), int(ts.numel()),
                                                      C.c_void_p(out.data_ptr()), int(B), int(T), stream))
            else:
                ts = ts.to(dtype=torch.float32).contiguous()
                _lib.check(L.eegldm_unet_forward(self._h, C.c_void_p(x.data_ptr()), C.cast(C.c_void_p(ts.data_ptr()), C.POINTER(C.c_float)),
                                                 int(ts.numel()), C.c_void_p(out.data_ptr()), int(B), int(T), stream))
        return out

    def train_step(self, z0, noise, timesteps, scheduler, lr=1e-4, betas=(0.9, 0.999), adam_eps=1e-8, return_loss=True):
        """One latent-diffusion training step on the device (the batch body of train_epoch_ldm, training.py:420-443):
        ``noisy = scheduler.add_noise(z0, noise, timesteps)``; ``pred = self(noisy, timesteps)``; target = ``noise`` or
        ``scheduler.get_velocity(z0, noise, timesteps)`` by ``scheduler.prediction_type``; ``F.mse_loss``; backward through the
        whole UNet; Adam.  ``z0`` is the scaled latent ``stage1(images) * scale_factor`` [B, C, T]; ``timesteps`` int64 [B]
        (CUDA, as the reference draws them).  ``lr <= 0`` computes the loss and gradients only.  Parameters are updated inside
        the engine: ``sync_trained()`` copies them back into this module / the inference weights; ``grad_dict()`` returns the
        last gradients.  fp32 (SIMT) or f16x3 (tensor pipe) by ``set_math``; returns the loss (one sync) if ``return_loss``."""
        z0 = check_cuda_f32(z0, "z0")
        noise = check_cuda_f32(noise, "noise")
        if z0.dim() != 3 or z0.shape[1] != self.in_channels or noise.shape != z0.shape:
            raise ValueError(f"z0 / noise must be [B, {self.in_channels}, T] and equal in shape")
        B, _, T = z0.shape
        ts = torch.as_tensor(timesteps).reshape(-1).to(device=z0.device, dtype=torch.int64).contiguous()
        if ts.numel() != B:
            raise ValueError("timesteps must have B entries")
        cfg = _lib.LdmTrainCfg(float(lr), float(betas[0]), float(betas[1]), float(adam_eps))
        out = (C.c_float * 1)()
        with torch.cuda.device(z0.device):
            self._sync_weights()
            _lib.check(_lib.lib().eegldm_unet_train_step(
                self._h, C.byref(scheduler._cfg), C.c_void_p(z0.data_ptr()), C.c_void_p(noise.data_ptr()), C.c_void_p(ts.data_ptr()),
                int(B), int(T), C.byref(cfg), out if return_loss else None, C.c_void_p(_lib.current_stream_ptr(z0.device))))
        self._trained = True
        return float(out[0]) if return_loss else None

    def _export(self, what: int):
        res = {}
        L = _lib.lib()
        for name, p in self.named_parameters():
            buf = torch.empty(tuple(p.shape), dtype=torch.float32)
            _lib.check(L.eegldm_unet_train_export(self._h, what, name.encode(), C.cast(C.c_void_p(buf.data_ptr()), C.POINTER(C.c_float))))
            res[name] = buf
        return res

    def grad_dict(self):
        """Gradients of the last train_step, keyed and laid out like ``state_dict()``."""
        return self._export(1)

    @torch.no_grad()
    def sync_trained(self):
        """Copy the engine's trained parameters into this module and into the inference weights."""
        if not getattr(self, "_trained", False):
            return self
        new = self._export(0)
        for name, p in self.named_parameters():
            p.copy_(new[name].to(p.device))
        _lib.check(_lib.lib().eegldm_unet_train_sync(self._h))
        self._uploaded_key = self._weights_key()   # the engine already holds exactly these values
        return self

    def range_overflow(self) -> bool:
        """f16x3 operand-range guard: True when, since the last call, some activation handed to the tensor pipe had
        |x| >= 65504 or was NaN (that forward's output is invalid: switch to ``set_math("fp32")``).  Synchronises."""
        v = C.c_int(0)
        _lib.check(_lib.lib().eegldm_unet_range_status(self._h, C.byref(v)))
        return bool(v.value)
