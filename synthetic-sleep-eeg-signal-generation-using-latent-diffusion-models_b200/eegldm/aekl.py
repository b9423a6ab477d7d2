"""Drop-in for ``generative.networks.nets.AutoencoderKL`` (monai-generative) as the reference builds
it (``src/train_autoencoderkl.py:129-133``, ``src/sample_trials.py:95-100``, ``config/config_aekl_eeg*.yaml``):
1-D, GroupNorm(norm_num_groups) + SiLU ResBlocks, no attention.  Same ``state_dict`` keys (MONAI's
``Convolution(conv_only=True)`` naming) and the same method surface.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._module import EngineModule, check_cuda_f32, default_init, register_tree


class _AeklTrainFn(torch.autograd.Function):
    """``AutoencoderKL.forward`` across the autograd boundary: forward = ``eegldm_aekl_forward_train`` (the tensors the backward pass
    needs stay in the engine handle), backward = ``eegldm_aekl_backward`` + one gradient per parameter.  The parameters are passed
    as inputs so that autograd routes their gradients to ``p.grad`` (what the reference's ``optimizer_g.step()`` consumes)."""

    @staticmethod
    def forward(ctx, mod, x, eps, *params):
        B, _, Lx = x.shape
        T = Lx // mod._factor
        recon = torch.empty((B, mod.out_channels, Lx), device=x.device, dtype=torch.float32)
        z_mu = torch.empty((B, mod.latent_channels, T), device=x.device, dtype=torch.float32)
        z_sigma = torch.empty_like(z_mu)
        with torch.cuda.device(x.device):
            mod._sync_weights()
            _lib.check(_lib.lib().eegldm_aekl_forward_train(
                mod._h, C.c_void_p(x.data_ptr()), C.c_void_p(eps.data_ptr()), C.c_void_p(recon.data_ptr()), C.c_void_p(z_mu.data_ptr()),
                C.c_void_p(z_sigma.data_ptr()), int(B), int(Lx), C.c_void_p(_lib.current_stream_ptr(x.device))))
        mod._pass_token = getattr(mod, "_pass_token", 0) + 1
        ctx.mod, ctx.token, ctx.dev = mod, mod._pass_token, x.device
        return recon, z_mu, z_sigma

    @staticmethod
    def backward(ctx, d_recon, d_mu, d_sigma):
        mod = ctx.mod
        if ctx.token != mod._pass_token:
            raise RuntimeError("eegldm.AutoencoderKL: backward through a forward pass that a later forward() has replaced "
                               "(the engine keeps one recorded pass per model)")

        def ptr(t):
            return None if t is None else C.c_void_p(t.contiguous().float().data_ptr())
        keep = [None if t is None else t.contiguous().float() for t in (d_recon, d_mu, d_sigma)]
        with torch.cuda.device(ctx.dev):
            _lib.check(_lib.lib().eegldm_aekl_backward(
                mod._h, *[None if t is None else C.c_void_p(t.data_ptr()) for t in keep], None,
                C.c_void_p(_lib.current_stream_ptr(ctx.dev))))
        grads = mod._export(1)
        return (None, None, None) + tuple(grads[n].to(ctx.dev) for n, _ in mod.named_parameters())


class AutoencoderKL(EngineModule):
    def __init__(self, spatial_dims=1, in_channels=1, out_channels=1, num_res_blocks=(2, 2, 2, 2),
                 num_channels=(32, 64, 64, 64), attention_levels=(False, False, True, True), latent_channels=3,
                 norm_num_groups=32, norm_eps=1e-6, with_encoder_nonlocal_attn=True, with_decoder_nonlocal_attn=True,
                 use_flash_attention=False):
        super().__init__()
        if spatial_dims != 1:
            raise NotImplementedError("the reference's EEG autoencoder is 1-D (config_aekl_eeg.yaml:20)")
        num_channels = list(num_channels)
        if isinstance(num_res_blocks, int):
            num_res_blocks = [num_res_blocks] * len(num_channels)
        num_res_blocks = list(num_res_blocks)
        attention_levels = list(attention_levels)[:len(num_channels)]
        if any(attention_levels) or with_encoder_nonlocal_attn or with_decoder_nonlocal_attn:
            raise NotImplementedError("attention is off in every reference autoencoder config "
                                      "(config_aekl_eeg.yaml:26-28); pass attention_levels=[False,...], "
                                      "with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False")
        if abs(norm_eps - 1e-6) > 0:
            raise NotImplementedError("norm_eps is fixed at 1e-6")
        if len(num_res_blocks) != len(num_channels) or len(num_channels) > 8:
            raise ValueError("num_res_blocks / num_channels length mismatch")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_channels, self.latent_channels = num_channels, latent_channels
        cfg = _lib.AeklCfg()
        cfg.in_channels, cfg.out_channels, cfg.n_levels = int(in_channels), int(out_channels), len(num_channels)
        for i, (c, r) in enumerate(zip(num_channels, num_res_blocks)):
            cfg.num_channels[i], cfg.num_res_blocks[i] = int(c), int(r)
        cfg.latent_channels, cfg.norm_num_groups = int(latent_channels), int(norm_num_groups)
        self._cfg = cfg
        self._factor = 1 << (len(num_channels) - 1)
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.eegldm_aekl_create(C.byref(cfg), C.byref(h)))
        self._h = h
        infos = _lib.param_infos(h, L.eegldm_aekl_num_params, L.eegldm_aekl_param_info)
        shapes = dict(infos)

        def init(name, shape):
            return default_init(name, shape, shapes[name.rsplit(".", 1)[0] + ".weight"])

        register_tree(self, infos, init)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().eegldm_aekl_destroy(h)
            except Exception:
                pass
            object.__setattr__(self, "_h", None)   # nn.Module.__setattr__ may already be torn down at interpreter exit

    def _upload(self, state_dict) -> None:
        L = _lib.lib()
        _lib.load_state_dict_into(self._h, L.eegldm_aekl_load, state_dict)
        _lib.check(L.eegldm_aekl_finalize(self._h))

    # --- reference method surface ---------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x):
        """-> (z_mu, z_sigma)"""
        x = check_cuda_f32(x, "x")
        B, Cin, Lx = x.shape
        if Cin != self.in_channels:
            raise ValueError("channel mismatch")
        T = Lx // self._factor
        z_mu = torch.empty((B, self.latent_channels, T), device=x.device, dtype=torch.float32)
        z_sigma = torch.empty_like(z_mu)
        with torch.cuda.device(x.device):
            self._sync_weights()
            _lib.check(_lib.lib().eegldm_aekl_encode(
                self._h, C.c_void_p(x.data_ptr()), C.c_void_p(z_mu.data_ptr()), C.c_void_p(z_sigma.data_ptr()),
                int(B), int(Lx), C.c_void_p(_lib.current_stream_ptr(x.device))))
        return z_mu, z_sigma

    def sampling(self, z_mu, z_sigma):
        eps = torch.randn_like(z_sigma)
        return z_mu + eps * z_sigma

    @torch.no_grad()
    def decode(self, z):
        z = check_cuda_f32(z, "z")
        B, Cz, T = z.shape
        if Cz != self.latent_channels:
            raise ValueError("channel mismatch")
        out = torch.empty((B, self.out_channels, T * self._factor), device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device):
            self._sync_weights()
            _lib.check(_lib.lib().eegldm_aekl_decode(
                self._h, C.c_void_p(z.data_ptr()), C.c_void_p(out.data_ptr()), int(B), int(T),
                C.c_void_p(_lib.current_stream_ptr(z.device))))
        return out

    # --- training step (generator half of src/train_autoencoderkl.py:204-220) ------------------------
    def attach_discriminator(self, discriminator=None, seed=None, adv_weight=0.01, lr_d=5e-4, **disc_kwargs):
        """Make ``train_step`` the FULL step of train_autoencoderkl.py:204-234 by default: keeps ``discriminator`` (or builds
        the reference's ``PatchDiscriminator(num_layers_d=3, num_channels=64, kernel_size=3)``, config_aekl_eeg.yaml:30-40)."""
        from .discriminator import PatchDiscriminator
        if discriminator is None:
            if seed is not None:
                torch.manual_seed(seed)
            kw = dict(spatial_dims=1, num_layers_d=3, num_channels=64, in_channels=1, out_channels=1, kernel_size=3, norm="BATCH",
                      bias=False, padding=1)
            kw.update(disc_kwargs)
            discriminator = PatchDiscriminator(**kw)
        dev = next(self.parameters()).device
        object.__setattr__(self, "_disc", discriminator.to(dev))   # not a sub-module: its parameters belong to optimizer_d
        self._adv_weight, self._lr_d = adv_weight, lr_d
        return discriminator

    def train_step(self, x, eps=None, kl_weight=1e-9, spectral_weight=1e4, lr=5e-3, betas=(0.9, 0.999), adam_eps=1e-8,
                   return_losses=True, discriminator=None, adv_weight=None, lr_d=None, no_activation_leastsq=False):
        """One fused step on the device: forward, L1 + kl_weight*KL + spectral_weight*JukeboxLoss, backward, Adam.
        ``eps`` is the reparameterisation noise (drawn with torch.randn if omitted); ``lr <= 0`` computes losses and
        gradients only.  Parameters are updated inside the engine; call ``sync_trained()`` to copy them back into this
        module's ``nn.Parameter``s / the inference weights.  Returns {"l1", "kl", "spectral", "total"} (one sync).

        With ``discriminator`` (an ``eegldm.PatchDiscriminator``, or one attached by ``attach_discriminator``) this is the
        whole step of train_autoencoderkl.py:204-234: the generator loss gains ``adv_weight * MSE(lrelu(D(recon)), 1)`` and the
        discriminator takes its own step (``lr_d``) on ``0.5 * (MSE(lrelu(D(recon.detach())), 0) + MSE(lrelu(D(x)), 1))``;
        the returned dict then also has ``"generator"`` and ``"discriminator"``."""
        disc = discriminator if discriminator is not None else getattr(self, "_disc", None)
        if disc is not None:
            return self._train_step_adv(x, eps, disc, kl_weight, spectral_weight, lr, betas, adam_eps, return_losses,
                                        getattr(self, "_adv_weight", 0.01) if adv_weight is None else adv_weight,
                                        getattr(self, "_lr_d", 5e-4) if lr_d is None else lr_d, no_activation_leastsq)
        x = check_cuda_f32(x, "x")
        B, Cin, Lx = x.shape
        T = Lx // self._factor
        if eps is None:
            eps = torch.randn((B, self.latent_channels, T), device=x.device, dtype=torch.float32)
        eps = check_cuda_f32(eps, "eps")
        cfg = _lib.AeklTrainCfg(float(kl_weight), float(spectral_weight), float(lr), float(betas[0]), float(betas[1]), float(adam_eps))
        out = (C.c_float * 4)()
        with torch.cuda.device(x.device):
            self._sync_weights()
            _lib.check(_lib.lib().eegldm_aekl_train_step(
                self._h, C.c_void_p(x.data_ptr()), C.c_void_p(eps.data_ptr()), int(B), int(Lx), C.byref(cfg),
                out if return_losses else None, C.c_void_p(_lib.current_stream_ptr(x.device))))
        self._trained = True
        if return_losses:
            return {"l1": out[0], "kl": out[1], "spectral": out[2], "total": out[3]}
        return None

    def _train_step_adv(self, x, eps, disc, kl_weight, spectral_weight, lr, betas, adam_eps, return_losses, adv_weight, lr_d, no_act):
        x = check_cuda_f32(x, "x")
        B, Cin, Lx = x.shape
        T = Lx // self._factor
        if eps is None:
            eps = torch.randn((B, self.latent_channels, T), device=x.device, dtype=torch.float32)
        eps = check_cuda_f32(eps, "eps")
        cfg = _lib.AeklAdvTrainCfg(float(kl_weight), float(spectral_weight), float(adv_weight), float(lr), float(lr_d), float(betas[0]),
                                   float(betas[1]), float(adam_eps), int(bool(no_act)))
        out = (C.c_float * 6)()
        with torch.cuda.device(x.device):
            self._sync_weights()
            disc._sync_weights()
            _lib.check(_lib.lib().eegldm_aekl_train_step_adv(
                self._h, disc._h, C.c_void_p(x.data_ptr()), C.c_void_p(eps.data_ptr()), int(B), int(Lx), C.byref(cfg),
                out if return_losses else None, C.c_void_p(_lib.current_stream_ptr(x.device))))
        self._trained = True
        disc._trained = True
        if return_losses:
            return {"l1": out[0], "kl": out[1], "spectral": out[2], "total": out[3], "generator": out[4], "discriminator": out[5]}
        return None

    def _export(self, what: int):
        res = {}
        L = _lib.lib()
        for name, p in self.named_parameters():
            buf = torch.empty(tuple(p.shape), dtype=torch.float32)
            _lib.check(L.eegldm_aekl_train_export(self._h, what, name.encode(), C.cast(C.c_void_p(buf.data_ptr()), C.POINTER(C.c_float))))
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
        _lib.check(_lib.lib().eegldm_aekl_train_sync(self._h))
        self._uploaded_key = self._weights_key()   # the engine already holds exactly these values
        return self

    def reconstruct(self, x):
        z_mu, _ = self.encode(x)
        return self.decode(z_mu)

    def forward(self, x):
        """-> (reconstruction, z_mu, z_sigma), ``generative``'s AutoencoderKL.forward.  Under autograd (grad mode on and a parameter
        that requires grad) the call is differentiable with respect to the parameters, so the reference's own training loop runs
        unchanged (train_autoencoderkl.py:204-220: ``model(x)`` -> losses -> ``loss_g.backward()`` -> ``optimizer_g.step()``);
        ``train_step`` is the fused, much faster form of the same step."""
        # (the engine's training kernels cover in / out / latent channels = 1, every reference config; other widths run the
        # inference path, whose outputs do not require grad -- a .backward() on them fails loudly in PyTorch)
        trainable = self.in_channels == 1 and self.out_channels == 1 and self.latent_channels == 1
        if trainable and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            x = check_cuda_f32(x, "x")
            B, Cin, Lx = x.shape
            if Cin != self.in_channels or Lx % self._factor:
                raise ValueError("bad input shape")
            eps = torch.randn((B, self.latent_channels, Lx // self._factor), device=x.device, dtype=torch.float32)   # sampling()
            return _AeklTrainFn.apply(self, x.detach(), eps, *self.parameters())
        z_mu, z_sigma = self.encode(x)
        z = self.sampling(z_mu, z_sigma)
        return self.decode(z), z_mu, z_sigma

    def encode_stage_2_inputs(self, x):
        z_mu, z_sigma = self.encode(x)
        return self.sampling(z_mu, z_sigma)

    def decode_stage_2_outputs(self, z):
        return self.decode(z)
