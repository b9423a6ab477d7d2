/* eegldm -- C ABI of the B200-native 1-D latent-diffusion engine for sleep-EEG windows.
 *
 * The reference (bruAristimunha/Synthetic-Sleep-EEG-Signal-Generation-using-Latent-Diffusion-Models)
 * has no FFI: its boundary for this path is the Python nn.Module call surface.  Each entry point below
 * names the reference interface (file:line under /root/reference) it replaces; the ctypes binding that
 * re-exposes the reference's Python signatures on top of this ABI lives in
 * <package>/eegldm/ and is described in INTEGRATION.md.
 *
 * Conventions
 *  - All tensors are fp32, contiguous, in the REFERENCE layout: [B, C, T] (PyTorch NCL).
 *  - "dev" pointers are device pointers on the current CUDA device (e.g. torch.Tensor.data_ptr());
 *    "host" pointers are ordinary host memory.  The caller owns every I/O buffer; the engine owns
 *    weights and workspace.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls taking dev
 *    pointers are asynchronous on that stream and never synchronise; the *_host variants copy in/out
 *    and return after the result is in host memory.
 *  - Return value: 0 on success, a negative eegldm_status otherwise; eegldm_last_error() returns a
 *    thread-local message.  Nothing throws or aborts across the ABI.
 *  - A handle is not thread-safe; use one handle per (device, stream).
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    EEGLDM_ERR_CUDA.
 */
#ifndef EEGLDM_H_
#define EEGLDM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    EEGLDM_OK = 0,
    EEGLDM_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
    EEGLDM_ERR_SHAPE = -2,        /* tensor shape does not match the configuration */
    EEGLDM_ERR_MISSING = -3,      /* a state_dict entry was never loaded / unknown name */
    EEGLDM_ERR_CUDA = -4,         /* CUDA runtime error (message has the cudaError string) */
    EEGLDM_ERR_NOMEM = -5
} eegldm_status;

/* Arithmetic used by the convolution GEMMs.
 *  FP32_SIMT : fp32 FMA on the CUDA cores, the parity baseline.
 *  F16X3_TC  : tcgen05 tensor cores; each fp32 operand is split x = hi + lo*2^-11 (hi, lo fp16) and 3 products
 *              are issued (hi*hi | hi*lo + lo*hi in a second TMEM accumulator): ~2^-22 operand error, fp32
 *              accumulate.  Meets the fp32 parity tolerance; operands must satisfy |x| < 65504.
 *  BF16_TC   : single bf16 product (fast mode; does NOT meet the fp32 parity tolerance). */
typedef enum { EEGLDM_MATH_FP32_SIMT = 0, EEGLDM_MATH_F16X3_TC = 1, EEGLDM_MATH_BF16_TC = 2 } eegldm_math;

const char* eegldm_last_error(void);
const char* eegldm_version(void);
/* number of CUDA kernels launched by this library since process start (bench.py's gpu_launches) */
int64_t eegldm_launch_count(void);

/* Denoise-step graph of eegldm_ddim_sample: lanes = 2 plans the batch as two independent halves captured on two streams
 * inside the graph, so the HBM-bound passes of one half overlap the tensor-bound convolutions of the other; lanes = 1
 * (default; measured faster under the board power cap) is a single chain.  Rows are independent: results are bit-identical.  Takes effect for graphs captured afterwards
 * (new (B, T) shapes or new models). */
int eegldm_set_sample_lanes(int lanes);

/* Tuning knob of the tcgen05 conv kernel: CTAs per thread-block cluster that share each weight stage through a
 * multicast bulk copy (1, 2 or 4; default 2).  Changing it invalidates nothing but must not race with launches. */
int eegldm_set_conv_cluster(int ctas);
/* Shape selection of the tcgen05 conv kernel.  pair (bit mask, default 1): bit 0 -- the launches with 256-wide tiles run as CTA
 * pairs: the two CTAs of a cluster issue one M=256 cta_group::2 MMA over both, each stages half of the weight columns, loads are
 * cp.async.bulk.tensor.cta_group::2 completing on the leader's mbarrier; bit 1 -- the same for the 128-wide launches (measured
 * slower: off); 0: single-CTA MMAs with the multicast cluster of eegldm_set_conv_cluster.
 * bn256_min_stages (default 1): tiles are 256 output channels wide when Cout % 256 == 0 and a tile's mainloop has at least this
 * many weight stages (else 128).  fuse_epilogues (bit mask, default 15):
 * bit 0 -- a conv whose output feeds a GroupNorm writes that GroupNorm's statistics from its epilogue (no separate pass);
 * bit 1 -- an AttentionBlock's qkv conv writes the attention kernel's fp16 hi/lo operand images instead of fp32 (f16x3): no
 *          qkv_split pass;
 * bit 2 -- the conv kernel's producer warps read the fp32 input and build the fp16 hi/lo operand tiles in shared memory
 *          themselves (GroupNorm apply + SiLU + nearest-x2 + split), replacing the act_split pre-pass and its U tensors;
 * bit 3 -- the tcgen05 attention kernel writes its result as proj_out's operand image (no fp32 attention output, no pre-pass);
 * bit 4 -- the tcgen05 attention kernel reads the fp32 qkv tensor and splits q, k, v to fp16 hi/lo in its own producer warps
 *          (only without bit 1; T <= 208; measured no faster: off);
 * bit 5 -- (default clear) set to switch OFF the N = 128 tiles' concatenated MMA a_hi x [w_hi | w_lo] (one N = 256 instruction
 *          into both f16x3 accumulators instead of two N = 128 ones; same arithmetic, fewer shared-memory operand reads);
 * bit 6 -- (default clear) set to switch OFF the two-warpgroup conv epilogue (eight warps draining the accumulators in 16-column
 *          chunks instead of four in 32-column chunks; single-CTA launches only, CTA pairs always use it);
 * bit 7 -- fused producer also for 1x1 convs with more than two N tiles (the qkv conv; measured slower: off);
 * bit 8 -- (default clear) set for epilogue GroupNorm records at the consumer's group width instead of 4 / 8 channels (then the
 *          concat norms with 12- / 24-channel groups need their own pass over the tensor again).
 * bit 10 -- (default clear) set for one GroupNorm record per 16-position segment everywhere (default: one per 128-row tile where the
 *          length is a multiple of 128, so that a tile never straddles samples).
 * bit 9 -- (default clear) set to switch OFF the polyphase form of the up-sampling ResBlocks' first conv (nearest x2 -> 3-tap conv
 *          computed on the low-resolution input as an even and an odd output phase with two taps each: a third fewer MMAs).
 * Call before creating models: plans cache the choices. */
int eegldm_set_conv_tuning(int pair, int bn256_min_stages, int fuse_epilogues);

/* Live per-kernel profile (bench.py's roofline leg).  While enabled, every launch made outside CUDA-graph
 * capture is bracketed by CUDA events on the launching stream.  eegldm_profile_read sums, for one kernel
 * family (0 = conv implicit-GEMM [in a tensor-pipe math mode: the tcgen05 launches only], 1 = GroupNorm statistics, 2 = attention,
 * 3 = other, 4 = activation split pre-pass, 5 = the narrow fp32 SIMT convs of a tensor-pipe model: 1-channel in / out convs and the autoencoder), the measured
 * milliseconds, the ALGORITHMIC flops and HBM bytes and the launch count since enable. */
int eegldm_profile_enable(int on);
int eegldm_profile_read(int kind, double* ms, double* flops, double* bytes, int64_t* launches);
/* the i-th recorded launch (in launch order) since enable; returns EEGLDM_ERR_INVALID past the end */
int eegldm_profile_record(int i, int* kind, double* ms, double* flops, double* bytes);

/* ------------------------------------------------------------------------------------------------
 * Denoiser: replaces UNetModel (src/models/unet.py:330-563).  Fields mirror the constructor kwargs
 * (unet.py:331-351) that the reference's configs set (config/config_ldm.yaml:30-43). */
typedef struct {
    int32_t image_size;                 /* informational (unet.py:357); any T divisible by 2^(levels-1) runs */
    int32_t in_channels;
    int32_t model_channels;
    int32_t out_channels;
    int32_t num_res_blocks;
    int32_t n_attention_resolutions;
    int32_t attention_resolutions[8];
    int32_t n_channel_mult;
    int32_t channel_mult[8];
    int32_t num_heads;
    int32_t num_head_channels;          /* -1 = use num_heads */
    int32_t num_heads_upsample;         /* -1 = num_heads; heads of the output-block attention (unet.py:353-354,476) */
    int32_t resblock_updown;            /* 1 = ResBlock up/down (config_ldm.yaml:43) */
    int32_t conv_resample;              /* used when resblock_updown = 0 */
    int32_t use_scale_shift_norm;       /* must be 0 (no reference config enables it) */
} eegldm_unet_cfg;

typedef struct eegldm_unet eegldm_unet;

int eegldm_unet_create(const eegldm_unet_cfg* cfg, eegldm_unet** out);
void eegldm_unet_destroy(eegldm_unet* h);
/* number of state_dict entries the configuration expects and the name/shape of entry i
 * (order = the reference's registration order; key grammar: SURVEY.md section 8c) */
int eegldm_unet_num_params(const eegldm_unet* h);
int eegldm_unet_param_info(const eegldm_unet* h, int i, const char** name, int64_t shape[4], int* ndim);
/* nn.Module.load_state_dict, one entry per call (host pointer, reference [Cout,Cin,k] layout);
 * the engine copies and repacks. */
int eegldm_unet_load(eegldm_unet* h, const char* name, const float* host, const int64_t* shape, int ndim);
/* verifies every entry was loaded, uploads and packs the weights (strict=True semantics) */
int eegldm_unet_finalize(eegldm_unet* h);
int eegldm_unet_set_math(eegldm_unet* h, eegldm_math mode);
/* UNetModel.forward(x, timesteps)  (unet.py:512-563)
 *   x_dev   [B, in_channels, T] ; out_dev [B, out_channels, T]
 *   timesteps_host: nt values, nt == 1 (broadcast, sample_trials.py:158) or nt == B (training.py:430);
 *   float, as the reference converts with .float() (unet.py:28). */
int eegldm_unet_forward(eegldm_unet* h, const float* x_dev, const float* timesteps_host, int nt, float* out_dev,
                        int B, int T, void* stream);
/* Same with DEVICE-resident timesteps (fp32, nt values): no host round trip.  The reference's training loop draws the
 * timesteps on the GPU (src/training/training.py:420-430) and its sampler passes a CUDA tensor (sample_trials.py:157-159). */
int eegldm_unet_forward_devt(eegldm_unet* h, const float* x_dev, const float* timesteps_dev, int nt, float* out_dev,
                             int B, int T, void* stream);
/* F16X3_TC operand-range guard.  Every kernel that splits fp32 values into fp16 hi/lo raises a device flag when a value has
 * |x| >= 65504 or is NaN (the split would yield inf/NaN: the output is then invalid).  This call synchronises with the
 * device, returns the flag in *overflow_out and clears it.  eegldm_ddim_sample_host checks it itself and fails with
 * EEGLDM_ERR_INVALID; weights outside the range keep their layer on the fp32 SIMT kernel at finalize time. */
int eegldm_unet_range_status(eegldm_unet* h, int* overflow_out);

/* ------------------------------------------------------------------------------------------------
 * Autoencoder: replaces generative.networks.nets.AutoencoderKL as constructed at
 * src/train_autoencoderkl.py:129-133 / src/sample_trials.py:95-100 (config/config_aekl_eeg*.yaml). */
typedef struct {
    int32_t in_channels;
    int32_t out_channels;
    int32_t n_levels;
    int32_t num_channels[8];
    int32_t num_res_blocks[8];
    int32_t latent_channels;
    int32_t norm_num_groups;
} eegldm_aekl_cfg;

typedef struct eegldm_aekl eegldm_aekl;

int eegldm_aekl_create(const eegldm_aekl_cfg* cfg, eegldm_aekl** out);
void eegldm_aekl_destroy(eegldm_aekl* h);
int eegldm_aekl_num_params(const eegldm_aekl* h);
int eegldm_aekl_param_info(const eegldm_aekl* h, int i, const char** name, int64_t shape[4], int* ndim);
int eegldm_aekl_load(eegldm_aekl* h, const char* name, const float* host, const int64_t* shape, int ndim);
int eegldm_aekl_finalize(eegldm_aekl* h);
/* AutoencoderKL.encode(x) -> (z_mu, z_sigma)   x [B,in,L] -> [B,z,L/2^(levels-1)] each */
int eegldm_aekl_encode(eegldm_aekl* h, const float* x_dev, float* z_mu_dev, float* z_sigma_dev, int B, int L,
                       void* stream);
/* AutoencoderKL.decode(z) / decode_stage_2_outputs   z [B,z,T] -> [B,out,T*2^(levels-1)] */
int eegldm_aekl_decode(eegldm_aekl* h, const float* z_dev, float* out_dev, int B, int T, void* stream);
/* AutoencoderKL.forward(x) with the sampling noise supplied by the caller:
 * z = mu + eps*sigma; recon = decode(z).  eps_dev [B,z,T]. */
int eegldm_aekl_forward(eegldm_aekl* h, const float* x_dev, const float* eps_dev, float* recon_dev,
                        float* z_mu_dev, float* z_sigma_dev, int B, int L, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Spectral loss: replaces generative.losses.JukeboxLoss(spatial_dims=1, reduction=...) as called at
 * src/train_autoencoderkl.py:158,208.  input/target [B, C=1, N] fp32 device; loss_dev receives
 * reduce((|fft_ortho(target)| - |fft_ortho(input)|)^2) (reduction 0 = "sum", 1 = "mean"); grad_input_dev
 * (nullable) receives d loss / d input.  Batched cuFFT R2C (+ C2R for the gradient) and one custom kernel. */
int eegldm_jukebox_loss(const float* input_dev, const float* target_dev, int B, int C, int N, int reduction, float* loss_dev,
                        float* grad_input_dev, void* stream);

/* Generator half of the autoencoder training step, src/train_autoencoderkl.py:204-220 (adversarial term excluded):
 *   recon, mu, sigma = model(x)  with  z = mu + eps * sigma  (eps supplied by the caller, [B, z, L/2^(levels-1)])
 *   loss = L1Loss(recon, x) + kl_weight * KL(mu, sigma) + spectral_weight * JukeboxLoss(recon, x);  backward;  Adam step.
 * lr <= 0 computes losses and gradients only.  losses_host (nullable; synchronises) receives
 * {l1, kl, spectral, total}.  Parameters, gradients and Adam moments live on the device in the engine's packed
 * layouts; eegldm_aekl_train_export returns one state_dict entry (what = 0) or its gradient (what = 1) in the
 * reference layout; eegldm_aekl_train_sync copies the trained parameters back into the inference weights
 * (encode / decode / sampling use them from then on). */
typedef struct {
    float kl_weight;        /* 1e-9  (config_aekl_eeg.yaml:15) */
    float spectral_weight;  /* 1e4   (config_aekl_eeg.yaml:17) */
    float lr;               /* 5e-3  optimizer_g_lr (config_aekl_eeg.yaml:12) */
    float beta1, beta2, adam_eps;   /* torch.optim.Adam defaults 0.9, 0.999, 1e-8 */
} eegldm_aekl_train_cfg;
int eegldm_aekl_train_step(eegldm_aekl* h, const float* x_dev, const float* eps_dev, int B, int L, const eegldm_aekl_train_cfg* cfg,
                           float* losses_host, void* stream);
int eegldm_aekl_train_export(eegldm_aekl* h, int what, const char* name, float* host_out);
int eegldm_aekl_train_sync(eegldm_aekl* h);
/* The same autoencoder across an AUTOGRAD boundary, for the reference's unchanged loop (src/train_autoencoderkl.py:204-220:
 * `reconstruction, z_mu, z_sigma = model(x)`; losses in PyTorch; `loss_g.backward()`; `optimizer_g.step()`):
 * eegldm_aekl_forward_train = AutoencoderKL.forward(x) in training mode with the sampling noise eps supplied by the caller; the
 * tensors the backward pass needs stay inside the handle.  eegldm_aekl_backward consumes that recorded pass: d_recon_dev /
 * d_mu_dev / d_sigma_dev are dL/d(reconstruction), dL/d(z_mu), dL/d(z_sigma) (each nullable = zero); the parameter gradients are
 * then read with eegldm_aekl_train_export(h, 1, name, ...).  dx_dev must be NULL (the input signal's gradient is not computed). */
int eegldm_aekl_forward_train(eegldm_aekl* h, const float* x_dev, const float* eps_dev, float* recon_dev, float* z_mu_dev,
                              float* z_sigma_dev, int B, int L, void* stream);
int eegldm_aekl_backward(eegldm_aekl* h, const float* d_recon_dev, const float* d_mu_dev, const float* d_sigma_dev, float* dx_dev,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Discriminator: replaces generative.networks.nets.PatchDiscriminator as built at src/train_autoencoderkl.py:135-137 from
 * config/config_aekl_eeg.yaml:30-40 (spatial_dims 1, norm "BATCH", bias false, LeakyReLU 0.2; kernel_size 3, padding 1).
 * state_dict keys: initial_conv.conv.{weight,bias}, {l}.conv.weight, {l}.adn.N.{weight,bias,running_mean,running_var,
 * num_batches_tracked}, final_conv.conv.{weight,bias}. */
typedef struct {
    int32_t in_channels;     /* 1 */
    int32_t out_channels;    /* 1 */
    int32_t num_channels;    /* 64 */
    int32_t num_layers_d;    /* 3 */
    int32_t kernel_size;     /* 3 */
    int32_t padding;         /* 1 */
} eegldm_disc_cfg;
typedef struct eegldm_disc eegldm_disc;
int eegldm_disc_create(const eegldm_disc_cfg* cfg, eegldm_disc** out);
void eegldm_disc_destroy(eegldm_disc* h);
int eegldm_disc_num_params(const eegldm_disc* h);
/* entries include the BatchNorm buffers (is_buffer = 1; num_batches_tracked has ndim 0 and is carried as a float) */
int eegldm_disc_param_info(const eegldm_disc* h, int i, const char** name, int64_t shape[4], int* ndim, int* is_buffer);
int eegldm_disc_load(eegldm_disc* h, const char* name, const float* host, const int64_t* shape, int ndim);
int eegldm_disc_finalize(eegldm_disc* h);
/* PatchDiscriminator.forward(x)[-1]: x_dev [B, 1, L] -> logits_dev [B, 1, eegldm_disc_out_len(L)].  training != 0: BatchNorm uses
 * batch statistics and updates the running ones (the reference never calls discriminator.eval()); else the running statistics. */
int eegldm_disc_forward(eegldm_disc* h, const float* x_dev, float* logits_dev, int B, int L, int training, void* stream);
int eegldm_disc_out_len(const eegldm_disc* h, int L);
/* The discriminator across an AUTOGRAD boundary (the reference's unchanged loop, train_autoencoderkl.py:213-234):
 * eegldm_disc_forward_train = discriminator(x)[-1] in training mode (batch statistics, running statistics updated once), the pass
 * recorded in `slot` (0 or 1: two passes may be alive, e.g. the fake and the real batch of the discriminator part);
 * eegldm_disc_backward consumes it: dlogits_dev [B,1,L_out] -> dx_dev [B,1,L] (nullable), and with want_param_grads the parameter
 * gradients of this pass (eegldm_disc_export(h, 1, name, ...)). */
int eegldm_disc_forward_train(eegldm_disc* h, const float* x_dev, float* logits_dev, int B, int L, int slot, void* stream);
int eegldm_disc_backward(eegldm_disc* h, int slot, const float* dlogits_dev, float* dx_dev, int want_param_grads, void* stream);
/* EEGLDM_MATH_F16X3_TC (default): the 64->128->256->512 convs, their data gradients (the forward kernel on transformed weights) and
 * their weight gradients (split-K GEMM over the positions) run on tcgen05 in the f16x3 arithmetic; EEGLDM_MATH_FP32_SIMT: fp32 FMA. */
int eegldm_disc_set_math(eegldm_disc* h, int mode);
/* one state_dict entry in the reference layout: what = 0 value (parameters and buffers), 1 = gradient of the last step */
int eegldm_disc_export(eegldm_disc* h, int what, const char* name, float* host_out);

/* The FULL autoencoder training step, src/train_autoencoderkl.py:204-234:
 *   loss_g = L1 + kl_weight * KL + spectral_weight * Jukebox + adv_weight * MSE(lrelu_0.05(D(recon)), 1);  backward;  Adam(lr_g)
 *   loss_d = adv_weight * 0.5 * (MSE(lrelu_0.05(D(recon.detach())), 0) + MSE(lrelu_0.05(D(x)), 1));      backward;  Adam(lr_d)
 * (PatchAdversarialLoss(criterion="least_squares"): the logits pass through LeakyReLU(0.05) unless no_activation_leastsq.)
 * losses_host (nullable; synchronises) receives {l1, kl, spectral, total_g, generator_loss, discriminator_loss}. */
typedef struct {
    float kl_weight, spectral_weight, adv_weight;   /* 1e-9, 1e4, 0.01  (config_aekl_eeg.yaml:14-17) */
    float lr_g, lr_d;                               /* 5e-3, 5e-4       (config_aekl_eeg.yaml:12-13) */
    float beta1, beta2, adam_eps;
    int32_t no_activation_leastsq;                  /* 0 (upstream default) */
} eegldm_aekl_adv_train_cfg;
int eegldm_aekl_train_step_adv(eegldm_aekl* h, eegldm_disc* disc, const float* x_dev, const float* eps_dev, int B, int L,
                               const eegldm_aekl_adv_train_cfg* cfg, float* losses_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Scheduler + sampling loop: replaces generative.networks.schedulers.DDIMScheduler as used at
 * src/sample_trials.py:136-145 and the loop at src/sample_trials.py:153-166. */
typedef struct {
    int32_t num_train_timesteps;   /* 1000 */
    float beta_start;              /* 0.0015 */
    float beta_end;                /* 0.0205 */
    int32_t schedule;              /* 0 = linear_beta, 1 = scaled_linear_beta */
    int32_t prediction_type;       /* 0 = epsilon, 1 = v_prediction */
    int32_t set_alpha_to_one;      /* 1 (upstream default) */
    int32_t steps_offset;          /* 0 */
} eegldm_sched_cfg;

/* host-only helpers (no GPU needed): alphas_cumprod[num_train_timesteps]; DDIM timesteps[n_steps]
 * (descending) and the per-step update coefficients coef[2*i+0..1] with
 * x_prev = coef0 * x + coef1 * model_output (eta = 0, clip_sample = False). */
int eegldm_sched_alphas_cumprod(const eegldm_sched_cfg* cfg, float* out);
int eegldm_sched_ddim_tables(const eegldm_sched_cfg* cfg, int n_steps, int64_t* timesteps, float* coef);
/* timestep_embedding (unet.py:12-36) on the host: out[nt][dim] */
int eegldm_timestep_embedding(const float* timesteps, int nt, int dim, float* out);

/* Full DDIM sampling of B windows on one GPU.
 *   noise_dev [B, z, T] (the reference draws torch.randn on the host, sample_trials.py:151)
 *   aekl may be NULL: then out_dev receives the final latent [B, z, T];
 *   otherwise out_dev receives decode(latent / scale_factor) [B, out, T*2^(levels-1)]  (sample_trials.py:166).
 * One CUDA graph per denoise step is captured on first use for (B, T) and replayed n_steps times. */
int eegldm_ddim_sample(eegldm_unet* unet, eegldm_aekl* aekl, const eegldm_sched_cfg* sched, const float* noise_dev,
                       float scale_factor, int n_steps, float* out_dev, int B, int T, void* stream);
/* Same, with HOST buffers: H2D of the noise, sampling, D2H of the result, stream-synchronised on return.
 * The buffers should be pinned for full copy bandwidth. */
int eegldm_ddim_sample_host(eegldm_unet* unet, eegldm_aekl* aekl, const eegldm_sched_cfg* sched,
                            const float* noise_host, float scale_factor, int n_steps, float* out_host, int B, int T,
                            void* stream);
/* disable (0) / enable (1) CUDA-graph replay inside eegldm_ddim_sample (default 1) */
int eegldm_set_graphs(int enabled);

/* ------------------------------------------------------------------------------------------------
 * Latent-diffusion training step: replaces the body of train_epoch_ldm, src/training/training.py:420-443, for one batch
 * (with the DDPMScheduler of src/train_ldm.py:199-200):
 *   noisy = scheduler.add_noise(z0, noise, timesteps)                          (training.py:429)
 *   pred  = model(noisy, timesteps)                                            (training.py:430; UNetModel.forward in training mode)
 *   target = noise (epsilon) | scheduler.get_velocity(z0, noise, timesteps)    (training.py:432-436, sched->prediction_type)
 *   loss = F.mse_loss(pred, target);  backward through the whole UNet;  Adam step (train_ldm.py:208)
 * z0_dev / noise_dev [B, in_channels, T] fp32 (z0 = stage1(images) * scale_factor, the caller's AutoencoderKL.encode + sampling);
 * timesteps_dev [B] int64 on the device (training.py:420).  in_channels == out_channels is required (the target has the
 * latent's shape).  Arithmetic: fp32 throughout -- tensor-pipe convolutions (forward, data gradient, weight gradient) in the
 * f16x3 scheme when the handle's math mode is EEGLDM_MATH_F16X3_TC, fp32 SIMT otherwise; the reference's fp16 autocast +
 * GradScaler (training.py:423,441-443) is a memory / speed device of its PyTorch path whose loss scaling cancels exactly, so the
 * fp32 result is the value it approximates.  lr <= 0 computes the loss and the gradients only.  loss_host (nullable;
 * synchronises) receives the loss.  Parameters, gradients and Adam moments live on the device (created from the loaded
 * state_dict at the first step); eegldm_unet_train_export returns one state_dict entry (what = 0) or its gradient (what = 1) in
 * the reference layout; eegldm_unet_train_sync copies the trained parameters back into the inference weights. */
typedef struct {
    float lr;                       /* 1e-4  base_lr (config_ldm.yaml:11) */
    float beta1, beta2, adam_eps;   /* torch.optim.Adam defaults 0.9, 0.999, 1e-8 */
} eegldm_ldm_train_cfg;
int eegldm_unet_train_step(eegldm_unet* h, const eegldm_sched_cfg* sched, const float* z0_dev, const float* noise_dev,
                           const int64_t* timesteps_dev, int B, int T, const eegldm_ldm_train_cfg* cfg, float* loss_host, void* stream);
int eegldm_unet_train_export(eegldm_unet* h, int what, const char* name, float* host_out);
int eegldm_unet_train_sync(eegldm_unet* h);
/* The same denoiser across an AUTOGRAD boundary, for the reference's unchanged loop (training.py:420-443: `noise_pred =
 * model(x=noisy_e, timesteps=timesteps)`; the loss in PyTorch; `scaler.scale(loss).backward()`; `scaler.step(optimizer)`):
 * eegldm_unet_forward_train = UNetModel.forward in training mode (x_dev [B, in, T], timesteps_dev [B] fp32 on the device, out_dev
 * [B, out, T]) with the pass recorded inside the handle; eegldm_unet_backward consumes it with d_out_dev = dL/d(output) and leaves the
 * parameter gradients for eegldm_unet_train_export(h, 1, name, ...). */
int eegldm_unet_forward_train(eegldm_unet* h, const float* x_dev, const float* timesteps_dev, float* out_dev, int B, int T, void* stream);
int eegldm_unet_backward(eegldm_unet* h, const float* d_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Output tail of the sampling scripts: replaces, batched over all windows, what src/sample_trials.py:169-197 (and
 * sample_trials_ddpm.py:105-128, util.py:66-112) do per window on the host --
 *   cropped = sample.cpu().numpy()[:, :, 36:-36];  np.save(sample_i.npy, cropped);
 *   spectrum = mne.EpochsArray(cropped, sfreq=100).compute_psd(fmax=18);  psds = 10 * log10(spectrum.average().get_data()). */
typedef struct {
    int32_t method;         /* 0 = multitaper (mne Epochs.compute_psd default), 1 = welch (mne Raw.compute_psd default) */
    float sfreq;            /* 100 (util.py:83) */
    float fmin, fmax;       /* 0, 18 (sample_trials.py:174); frequencies with fmin <= f <= fmax are returned */
    float bandwidth;        /* multitaper: full bandwidth in Hz; <= 0 = mne's default (half-bandwidth product NW = 4) */
    int32_t low_bias;       /* multitaper: keep tapers with concentration > 0.9 (mne default 1) */
    int32_t normalization;  /* multitaper: 0 = "length" (mne default), 1 = "full" (divide by sfreq) */
    int32_t n_fft;          /* welch: segment / FFT length (<= 0 = mne's default 256) */
    int32_t n_overlap;      /* welch: overlap between segments (mne default 0) */
    int32_t remove_dc;      /* subtract the mean of each segment (mne default 1) */
    int32_t db;             /* 1: return 10 * log10(psd) (sample_trials.py:181) */
} eegldm_psd_cfg;
/* number of returned frequencies and (freqs_host nullable) their values for signals of N samples */
int eegldm_psd_freqs(const eegldm_psd_cfg* cfg, int N, int* n_freqs_out, float* freqs_host);
/* PSD of B signals of N samples; signal b starts at x_dev + b * row_stride (row_stride >= N: a crop is an offset pointer, no
 * copy); psd_dev [B][n_freqs].  One window kernel, ONE batched cuFFT R2C over all windows x tapers (or segments), one reduction
 * kernel; asynchronous on `stream`. */
int eegldm_psd(const eegldm_psd_cfg* cfg, const float* x_dev, int B, int N, int64_t row_stride, float* psd_dev, void* stream);
/* host helper: scipy.signal.windows.dpss(N, half_nbw, Kmax, sym, norm=2, return_ratios=True) -- windows_out [Kmax][N] (unit
 * L2 norm before the sym = 0 truncation), ratios_out [Kmax] (spectral concentrations).  No GPU needed. */
int eegldm_dpss(int N, double half_nbw, int Kmax, int sym, double* windows_out, double* ratios_out);
/* sample[:, :, crop_left : L - crop_right] of `rows` device rows of L floats, copied straight into host memory (one strided
 * copy; synchronises the stream) */
int eegldm_crop_to_host(const float* x_dev, int64_t rows, int L, int crop_left, int crop_right, float* out_host, void* stream);
/* numpy .npy writer (format 1.0, '<f4', C order) and the per-window files of sample_trials.py:170:
 * <dir>/<prefix><first_index + i>.npy, each of shape [1, C, L], from a host batch [B][C][L] */
int eegldm_write_npy_f32(const char* path, const float* data_host, const int64_t* shape, int ndim);
int eegldm_save_windows_npy(const char* dir, const char* prefix, int64_t first_index, const float* data_host, int B, int C, int L);

/* ------------------------------------------------------------------------------------------------
 * Test hook (not part of the drop-in surface): ONE fused convolution launch
 *   out[B][Tc][Cout] = bias + res + Conv1d_k( resample( silu?( scale*x + shift ) ) ),  x [B][Tin][Cin] channels-last,
 * w_host in the reference's [Cout][Cin][k] layout, in the requested math mode.  Lets the tests compare the
 * tcgen05 kernel with the fp32 SIMT kernel and torch's conv1d one layer at a time.  Synchronises the stream. */
int eegldm_test_conv(const float* x_dev, const float* scale_dev, const float* shift_dev, int silu, int resample,
                     const float* w_host, const float* bias_host, const float* res_dev, int B, int Tin, int Cin, int Cout, int k,
                     int math, float* out_dev, void* stream);
/* tcgen05 convolution (f16x3, bias only) whose epilogue also writes the GroupNorm(G, eps 1e-6) statistics of its output:
 * out_dev [B][T][Cout], mean_dev / rstd_dev [B][G].  Shapes must satisfy the tensor-pipe constraints and Cout/G in {4,8,16,32}. */
int eegldm_test_conv_gn(const float* x_dev, const float* w_host, const float* bias_host, int B, int T, int Cin, int Cout, int k,
                        int G, float* out_dev, float* mean_dev, float* rstd_dev, void* stream);
/* AttentionBlock core on the fused path: qkv = Conv1d(C, 3C, 1)(x) written as attention operand images by the conv epilogue,
 * then QKVAttentionLegacy (unet.py:107-125).  x_dev, out_dev [B][T][C] channels-last, C = H*ch; w_host [3C][C][1]. */
int eegldm_test_qkv_attention(const float* x_dev, const float* w_host, const float* bias_host, int B, int T, int H, int ch,
                              float* out_dev, void* stream);
/* Timing hook (tools/conv_bench.py): average milliseconds of `reps` launches of the tcgen05 convolution on synthetic
 * data of the given shape; debug 0 = real kernel, 1 = operand copies skipped, 2 = MMAs skipped (timing experiments). */
int eegldm_bench_conv(int B, int T, int Cin, int Cout, int k, int with_res, int math, int debug, int reps, float* ms_out,
                      void* stream);
/* Timing hook (tools/attn_timeline.py): average milliseconds of `reps` launches of the tcgen05 attention kernel (f16x3, output
 * written as proj_out's operand image) on synthetic q, k, v; timeline_out (optional, [8]) receives per-CTA cycle averages
 * {1 S phase, 2 softmax, 3 PV + epilogue, 4 total, 7 number of CTAs}. */
int eegldm_bench_attention(int B, int T, int H, int ch, int reps, float* ms_out, double* timeline_out, void* stream);
/* Same, followed by one launch with per-CTA cycle counters: timeline_out[16] receives, averaged over the CTAs that ran,
 * {0 total cycles, 1 MMA warp waiting for a free accumulator, 2 ... for an activation stage, 3 ... for a weight stage,
 *  4 epilogue waiting for a full accumulator, 5 epilogue busy, 6 producer waiting for a free stage, 7 producer busy,
 *  8 loader waiting for a free weight stage, 9 tiles per CTA, ..., 15 CTAs}.  tools/conv_timeline.py */
int eegldm_bench_conv_timeline(int B, int T, int Cin, int Cout, int k, int with_res, int math, int debug, int reps, float* ms_out,
                               double* timeline_out, void* stream);

/* Test hook: ONE attention launch, QKVAttentionLegacy.forward (unet.py:107-125), on channels-last
 * qkv [B][T][H*3*ch] (legacy head layout) -> out [B][T][H*ch].  Synchronises the stream. */
int eegldm_test_attention(const float* qkv_dev, int B, int T, int H, int ch, int math, float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EEGLDM_H_ */
