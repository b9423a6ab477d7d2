"""Drop-ins for ``generative.networks.nets.PatchDiscriminator`` and ``generative.losses.PatchAdversarialLoss`` (monai-generative)
as the reference uses them (``src/train_autoencoderkl.py:135-137,156,213-234``; ``config/config_aekl_eeg.yaml:30-40``):
1-D, BatchNorm, LeakyReLU(0.2), no conv bias.  Same constructor keywords and ``state_dict`` keys (MONAI ``Convolution``
naming: ``initial_conv.conv.*``, ``{l}.conv.weight``, ``{l}.adn.N.*``, ``final_conv.conv.*``).

The discriminator's training lives inside the engine: ``AutoencoderKL.train_step(x, discriminator=disc)`` runs the whole
step of ``train_autoencoderkl.py:204-234`` (generator AND discriminator halves, both Adam updates) on the device.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib
from ._module import EngineModule, check_cuda_f32, _Node


class _DiscTrainFn(torch.autograd.Function):
    """``discriminator(x)[-1]`` across the autograd boundary (train_autoencoderkl.py:213-234): forward records the pass in one of
    the engine's two slots, backward returns the gradient with respect to the input signal (the generator's adversarial term)
    and one gradient per parameter (the discriminator part)."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        B, _, Lx = x.shape
        L = _lib.lib()
        out = torch.empty((B, 1, L.eegldm_disc_out_len(mod._h, int(Lx))), device=x.device, dtype=torch.float32)
        slot = mod._next_slot
        mod._next_slot ^= 1
        with torch.cuda.device(x.device):
            mod._sync_weights()
            _lib.check(L.eegldm_disc_forward_train(mod._h, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), int(B), int(Lx), int(slot),
                                                   C.c_void_p(_lib.current_stream_ptr(x.device))))
        mod._slot_token[slot] += 1
        ctx.mod, ctx.slot, ctx.token, ctx.dev, ctx.xshape = mod, slot, mod._slot_token[slot], x.device, tuple(x.shape)
        mod._trained = True
        mod._pull_buffers()   # BatchNorm running statistics moved inside the engine: mirror them into the module's buffers
        return out

    @staticmethod
    def backward(ctx, dlogits):
        mod = ctx.mod
        if ctx.token != mod._slot_token[ctx.slot]:
            raise RuntimeError("eegldm.PatchDiscriminator: backward through a forward pass that later forward() calls have replaced "
                               "(the engine keeps two recorded passes per discriminator)")
        want_dx = ctx.needs_input_grad[1]
        want_p = any(ctx.needs_input_grad[2:])
        dl = dlogits.contiguous().float()
        dx = torch.empty(ctx.xshape, device=ctx.dev, dtype=torch.float32) if want_dx else None
        with torch.cuda.device(ctx.dev):
            _lib.check(_lib.lib().eegldm_disc_backward(mod._h, int(ctx.slot), C.c_void_p(dl.data_ptr()),
                                                       C.c_void_p(dx.data_ptr()) if want_dx else None, int(want_p),
                                                       C.c_void_p(_lib.current_stream_ptr(ctx.dev))))
        if want_p:
            grads = mod._export(1)
            pg = tuple(grads[n].to(ctx.dev) for n, _ in mod.named_parameters())
        else:
            pg = tuple(None for _ in mod.parameters())
        return (None, dx) + pg


class PatchDiscriminator(EngineModule):
    def __init__(self, spatial_dims=1, num_channels=64, in_channels=1, out_channels=1, num_layers_d=3, kernel_size=4,
                 activation=("LEAKYRELU", {"negative_slope": 0.2}), norm="BATCH", bias=False, padding=1, dropout=0.0,
                 last_conv_kernel_size=None, math="f16x3"):
        super().__init__()
        if spatial_dims != 1:
            raise NotImplementedError("the reference's discriminator is 1-D (config_aekl_eeg.yaml:32)")
        if str(norm).upper() != "BATCH" or bias or dropout:
            raise NotImplementedError("the reference uses norm='BATCH', bias=False, no dropout (config_aekl_eeg.yaml:38-39)")
        slope = activation[1].get("negative_slope", 0.2) if isinstance(activation, (tuple, list)) and len(activation) > 1 else 0.2
        if abs(slope - 0.2) > 1e-12 or (isinstance(activation, (tuple, list)) and str(activation[0]).upper() != "LEAKYRELU"):
            raise NotImplementedError("activation is fixed at LeakyReLU(0.2) (the upstream default the reference keeps)")
        if last_conv_kernel_size not in (None, kernel_size):
            raise NotImplementedError("last_conv_kernel_size must equal kernel_size")
        self.num_layers_d, self.num_channels = num_layers_d, num_channels
        cfg = _lib.DiscCfg(int(in_channels), int(out_channels), int(num_channels), int(num_layers_d), int(kernel_size), int(padding))
        self._cfg = cfg
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.eegldm_disc_create(C.byref(cfg), C.byref(h)))   # rejects kernel_size != 3 / padding != 1
        self._h = h
        self._trained = False
        self._next_slot, self._slot_token = 0, [0, 0]
        self.set_math(math)
        g = torch.Generator().manual_seed(torch.initial_seed() & 0x7FFFFFFF)
        for i in range(L.eegldm_disc_num_params(h)):
            name, shape, nd, isb = C.c_char_p(), (C.c_int64 * 4)(), C.c_int(), C.c_int()
            _lib.check(L.eegldm_disc_param_info(h, i, C.byref(name), shape, C.byref(nd), C.byref(isb)))
            key = name.value.decode()
            shp = tuple(int(shape[k]) for k in range(nd.value))
            parts = key.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Node())
                mod = mod._modules[p]
            if isb.value:
                if key.endswith("num_batches_tracked"):
                    mod.register_buffer(parts[-1], torch.zeros((), dtype=torch.long))
                else:
                    mod.register_buffer(parts[-1], torch.ones(shp) if key.endswith("running_var") else torch.zeros(shp))
            else:   # upstream initialise_weights: conv N(0, 0.02); BatchNorm weight N(1, 0.02), bias 0; conv bias: PyTorch default
                if key.endswith("conv.weight"):
                    t = 0.02 * torch.randn(shp, generator=g)
                elif key.endswith("N.weight"):
                    t = 1.0 + 0.02 * torch.randn(shp, generator=g)
                elif key.endswith("N.bias"):
                    t = torch.zeros(shp)
                else:
                    fan_in = int(in_channels) * int(kernel_size) if key.startswith("initial_conv") else shp[0]
                    t = (torch.rand(shp, generator=g) * 2 - 1) / max(fan_in, 1) ** 0.5
                mod.register_parameter(parts[-1], nn.Parameter(t))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().eegldm_disc_destroy(h)
            except Exception:
                pass
            object.__setattr__(self, "_h", None)

    def set_math(self, mode: str) -> "PatchDiscriminator":
        """``"f16x3"`` (default; tcgen05 for the wide convs and their gradients, fp32-accurate) or ``"fp32"`` (SIMT)."""
        if mode not in ("fp32", "f16x3"):
            raise ValueError("math must be 'fp32' or 'f16x3'")
        _lib.check(_lib.lib().eegldm_disc_set_math(self._h, _lib.MATH_MODES[mode]))
        self._math = mode
        return self

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _upload(self, state_dict) -> None:
        L = _lib.lib()
        _lib.load_state_dict_into(self._h, L.eegldm_disc_load, state_dict)
        _lib.check(L.eegldm_disc_finalize(self._h))

    @torch.no_grad()
    def _pull_buffers(self):
        L = _lib.lib()
        sd = self.state_dict(keep_vars=True)
        for name, _ in self.named_buffers():
            buf = torch.empty(tuple(sd[name].shape), dtype=torch.float32)
            _lib.check(L.eegldm_disc_export(self._h, 0, name.encode(), C.cast(C.c_void_p(buf.data_ptr()), C.POINTER(C.c_float))))
            sd[name].copy_(buf.to(sd[name].device).to(sd[name].dtype))
        self._uploaded_key = self._weights_key()

    def forward(self, x):
        """-> list whose LAST element is the patch logits ``[B, 1, L_out]`` (the reference takes ``discriminator(x)[-1]``,
        train_autoencoderkl.py:213,226,228).  Upstream also returns every intermediate feature map; they are not
        materialised here (the list has one element).  In training mode under autograd (the input or a parameter requires grad)
        the call is differentiable -- the reference's own loop (train_autoencoderkl.py:213-234) runs unchanged; the fused
        ``AutoencoderKL.train_step(discriminator=...)`` is the fast form of the same step."""
        if self.training and torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            xin = check_cuda_f32(x, "x")
            return [_DiscTrainFn.apply(self, xin, *self.parameters())]
        return self._forward_nograd(x)

    @torch.no_grad()
    def _forward_nograd(self, x):
        x = check_cuda_f32(x, "x")
        B, Cin, Lx = x.shape
        L = _lib.lib()
        out = torch.empty((B, 1, L.eegldm_disc_out_len(self._h, int(Lx))), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            self._sync_weights()
            _lib.check(L.eegldm_disc_forward(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), int(B), int(Lx),
                                             int(self.training), C.c_void_p(_lib.current_stream_ptr(x.device))))
        if self.training:
            self._trained = True
        return [out]

    def _export(self, what: int):
        res = {}
        L = _lib.lib()
        names = [n for n, _ in self.named_parameters()] + ([n for n, _ in self.named_buffers()] if what == 0 else [])
        sd = self.state_dict()
        for name in names:
            buf = torch.empty(tuple(sd[name].shape), dtype=torch.float32)
            _lib.check(L.eegldm_disc_export(self._h, what, name.encode(), C.cast(C.c_void_p(buf.data_ptr()), C.POINTER(C.c_float))))
            res[name] = buf
        return res

    def grad_dict(self):
        """Gradients of the last discriminator step, keyed and laid out like ``state_dict()``."""
        return self._export(1)

    @torch.no_grad()
    def sync_trained(self):
        """Copy the engine's trained parameters and BatchNorm running statistics back into this module."""
        if not self._trained:
            return self
        new = self._export(0)
        sd = self.state_dict(keep_vars=True)
        for name, v in new.items():
            sd[name].copy_(v.to(sd[name].device).to(sd[name].dtype))
        self._uploaded_key = self._weights_key()
        return self


class PatchAdversarialLoss(nn.Module):
    """``PatchAdversarialLoss(criterion="least_squares")`` (train_autoencoderkl.py:156): MSE against 1 (real) / 0 (fake) of
    the logits passed through LeakyReLU(0.05) (upstream's default for least squares unless ``no_activation_leastsq``).
    Elementwise torch ops on the caller's tensors (autograd-capable); the fused training step has its own kernel."""

    def __init__(self, reduction="mean", criterion="least_squares", no_activation_leastsq=False):
        super().__init__()
        if criterion != "least_squares" or reduction != "mean":
            raise NotImplementedError("the reference uses criterion='least_squares' with the default mean reduction")
        self.real_label, self.fake_label = 1.0, 0.0
        self.no_activation_leastsq = no_activation_leastsq

    def forward(self, input, target_is_real, for_discriminator):
        if not for_discriminator and not target_is_real:
            target_is_real = True
        outs = input if isinstance(input, (list, tuple)) else [input]
        losses = []
        for o in outs:
            y = o if self.no_activation_leastsq else torch.nn.functional.leaky_relu(o, 0.05)
            losses.append(torch.mean((y - (self.real_label if target_is_real else self.fake_label)) ** 2))
        return torch.mean(torch.stack(losses))
