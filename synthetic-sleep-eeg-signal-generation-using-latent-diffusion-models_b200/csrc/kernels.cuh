// Kernel-level interface of the eegldm engine (internal; the public boundary is include/eegldm.h).
//
// Data layout in HBM: every activation is fp32 channels-last  [B][T][C]  (C contiguous).
// For the 1-channel signals / latents at the API boundary this is bit-identical to the
// reference's NCL layout; multi-channel latents are transposed once at the boundary.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <string>
#include <vector>

namespace eegldm {

// message returned by eegldm_last_error() (thread-local, engine.cu); for translation units other than engine.cu
void set_last_error(const std::string& msg);
// kernels launched by this library since process start (eegldm_launch_count)
extern std::atomic<long long> g_launch_count;

enum Resample : int { RS_NONE = 0, RS_AVGPOOL2 = 1, RS_NEAREST2 = 2 };

// One K-segment of the implicit GEMM  out[b,t,co] += sum_{ci,k} W[ci,k,co] * u[b, t*stride+k-pad, ci]
// where u = resample(act(concat(src0,src1))) and act(x) = silu?(scale*x + shift).
struct ConvSeg {
    const float* src0;   // [B][Tin][C0]
    const float* src1;   // [B][Tin][C1] or null (virtual channel concat, reference unet.py:553)
    int C0, C1;
    const float* scale;  // [B][C0+C1] GroupNorm scale (gamma*rstd) or null
    const float* shift;  // [B][C0+C1] GroupNorm shift (beta-mean*gamma*rstd)
    int silu;            // activation after the affine: 0 none, 1 SiLU, 2 LeakyReLU(0.2)
    int resample;        // Resample applied AFTER act (unet.py:308-313)
    int Tin;             // source length (before resample)
    const float* w;      // packed [(ci*taps+k)][Cout]
    int taps;            // 1 or 3
};

struct ConvParams {
    ConvSeg seg[2];
    int nseg;
    int Cout, Tout;
    int Tc;              // conv-input length after resample
    int stride;          // 1 or 2 (seg[0] only; seg[1] is a 1x1 at output resolution)
    int pad_left;        // 1 for "same" k3, 0 for k1 and for the AEKL pad-right-1/stride-2 downsample
    const float* bias;   // [Cout] (already summed over segments) or null
    const float* temb;   // temb[b*temb_stride + co] or null
    int temb_stride;
    const float* res;    // identity residual [B][res_Tin][Cout] or null
    int res_mode;        // Resample of the residual
    int res_Tin;
    float* out;          // [B][Tout][Cout]
    const float* ddim_x;     // if non-null: out = coef[0]*ddim_x + coef[1]*value   (DDIM step epilogue)
    const float* ddim_coef;  // device [2]
    int B;
    float* gn_rec;           // optional, narrow-input conv with Cout == 128 and Tout % 16 == 0 only (conv_narrow_in_gn_ok): GroupNorm
                             // records (64, mean, M2) of the output per (sample, 16-position segment, 4-channel group): [B][Tout/16][32][3]
    int gn_rec_tile;         // 1 (Tout % 128 == 0): one record (512, mean, M2) per 128 positions instead: [B][Tout/128][32][3]
};
bool conv_narrow_in_gn_ok(const ConvParams& p);

struct GnParams {
    const float* src0; const float* src1; int C0, C1;
    int T, G;
    const float* gamma; const float* beta; float eps;
    float* scale; float* shift;   // [B][C]
    float* partial;               // [B][nsplit][G][3] scratch
    int nsplit;
    int B;
    float* mean_out;              // optional [B][G] (training: saved for the backward pass)
    float* rstd_out;
    // launch_groupnorm_finalize only: records written by the producing convs' epilogues at THEIR group width.  The normalised
    // tensor is concat(src0, src1); source i has rec_G[i] record groups per split (rec_G = 0: `partial` is already at this
    // norm's G, single source).  Every group of this norm must be a whole number of records of ONE source.
    const float* partial1;
    int rec_G0, rec_G1;
    int nsplit1;                  // records per sample of source 1 (0: the same as `nsplit`)
};

struct AttnParams {
    const float* qkv;   // [B][T][H*3*ch]  (legacy head layout, unet.py:116-118)
    float* out;         // [B][T][H*ch]
    int T, H, ch, B;
};

cudaError_t launch_conv_simt(const ConvParams& p, cudaStream_t st);

// ---- tcgen05 path (conv_tc.cu) ---------------------------------------------------------------
constexpr int TC_BK = 32;               // input channels per k-step
constexpr int TC_U_HALF_BYTES = 9216;   // one 16-bit (hi or lo) activation tile image [8 seg x 18 slots x TC_BK]
// act_split: fp32 activations -> tile images U[m_tile][k-step][hi|lo][kc][slot][segment][8 ch]
struct ActSplitParams {
    const float* src0; const float* src1; int C0, C1;   // (virtual concat of) fp32 sources [B][Tin][C]
    const float* scale; const float* shift;             // [B][C0+C1] GroupNorm scale/shift or null
    int silu, resample, Tin;
    int Tout, nsegs16, nks;                             // conv-input length after resample; B*Tout/16; (C0+C1)/TC_BK
    uint8_t* U;
    uint8_t* U_raw;   // optional second image: the same rows WITHOUT affine / SiLU (skip_connection operand, unet.py:327)
    int* range_flag;  // optional: set to 1 when a value to be split is outside the f16x3 operand range (|x| >= 65504 or NaN)
};
size_t act_split_bytes(int nsegs16, int Cin);
cudaError_t launch_act_split(const ActSplitParams& p, bool x3, cudaStream_t st);
struct TcSeg {
    const uint8_t* U;   // pre-pass form: activation tile images written by act_split
    const uint8_t* w;   // packed by pack_conv_tc: [k-step][tap][hi|lo][Cout/8][4][8][8] 16-bit (tile-width agnostic)
    int taps, nks;      // nks = Cin/TC_BK
    // fused-producer form (TcConvParams.direct): the conv reads the fp32 source itself (resample: RS_NONE | RS_NEAREST2)
    const float* src0; const float* src1; int C0, C1;
    const float* scale; const float* shift; int silu, resample, Tin;
};
struct TcConvParams {
    TcSeg seg[2];
    int nseg;
    int Cout, Tout;     // stride 1, "same" padding: conv-input length == Tout
    int bn;             // output channels per tile: 128 or 256 (conv_tc_bn)
    int nsegs16;        // B*Tout/16 segments of 16 positions (8 per CTA)
    const float* bias; const float* temb; int temb_stride;
    const float* res; int res_mode; int res_Tin;
    float* out;
    uint8_t* qkv16;     // optional (then `out` is unused): the output is an AttentionBlock's qkv tensor -- write it as the fp16 hi/lo
    int qkv_H, qkv_ch;  //   q/k/v operand images of attn_tc.cu (layout of launch_qkv_split) for qkv_H heads of qkv_ch channels; f16x3 only
    float* gn_partial;  // optional: GroupNorm statistics of `out`, one (count, mean, M2) record per (sample, 16-position segment, group):
    int gn_cpg;         //   [B][Tout/16][Cout/gn_cpg][3]; gn_cpg = channels per group in {4, 8, 16, 32} (conv_tc_gn_ok)
    int gn_tile;        // 1 (needs Tout % 128 == 0: a 128-row tile never straddles samples; two-warpgroup epilogue): ONE record per
                        //   (tile, group) instead of eight per-segment ones -- [B][Tout/128][Cout/gn_cpg][3]: an eighth of the record
                        //   traffic and of gn_finalize's loop
    int direct;         // 1: activation operands produced inside the conv kernel from TcSeg.src0/src1 (no act_split pre-pass)
    int debug;          // timing experiments only (eegldm_bench_conv): 1 = no operand copies, 2 = no MMAs; results are garbage
    int* range_flag;    // optional (fused producer): set to 1 when an operand is outside the f16x3 range (|x| >= 65504 or NaN)
    int cat;            // f16x3, bn == 128: issue a_hi x [w_hi | w_lo] as one N = 256 MMA (set by launch_conv_tc from g_conv_tc_cat)
    int poly;           // 1: polyphase form of "nearest x2 upsample -> 3-tap conv" (ResBlock(up=True).in_layers, unet.py:263-268): the conv
                        //    runs on the LOW-resolution input; Cout = 2 x the real channel count = [even-output phase | odd-output phase],
                        //    weights (w0, w1+w2, 0) | (0, w0+w1, w2) (pack_conv_tc_poly): an N tile lies in one phase and skips that
                        //    phase's zero tap (2 taps of MMAs instead of 3); output position t of phase ph is row 2t + ph of
                        //    out [B][2 Tout][Cout/2]; bias / temb / GroupNorm records are indexed with the real channel.  No residual,
                        //    two-warpgroup epilogue only, bn == 256, (Cout/2) % 256 == 0.
    unsigned long long* timeline;   // optional (eegldm_bench_conv_timeline): per-CTA cycle counters, TC_TL_N per CTA
    // CTA-pair form (g_conv_tc_pair; filled by launch_conv_tc): 2-D tensor maps over the weight image ([rows of 256 u16], box = the
    // BN/2 columns of one CTA) and, in the pre-pass form, over the U image (box = one 18 KB stage) of each segment -- the pair's loads are
    // cp.async.bulk.tensor.cta_group::2, whose completion can signal the LEADER's mbarrier from either CTA
    CUtensorMap tmap_w[2];
    CUtensorMap tmap_u[2];
};
// per-CTA counters of the conv kernel's warp roles (cycles, summed over the CTA's tiles)
enum TcTimeline : int { TC_TL_TOTAL = 0, TC_TL_MMA_WAIT_ACC, TC_TL_MMA_WAIT_A, TC_TL_MMA_WAIT_B, TC_TL_EPI_WAIT, TC_TL_EPI_BUSY,
                        TC_TL_PROD_WAIT, TC_TL_PROD_BUSY, TC_TL_LOAD_WAIT_B, TC_TL_TILES, TC_TL_N = 16 };
bool conv_tc_eligible(int Cin0, int Cin1, int Cout, int Tout, int taps, int stride);
bool conv_tc_gn_ok(int Cout, int G);         // can the conv epilogue emit the GroupNorm(G) statistics of its output?
int conv_tc_bn(int Cout, int weight_stages);   // weight_stages = sum over segments of (Cin/32)*taps
void pack_conv_tc(const float* w, int Cout, int Cin, int k, bool x3, std::vector<uint16_t>& out);
void pack_conv_tc_poly(const float* w, int Cout, int Cin, bool x3, std::vector<uint16_t>& out);   // TcConvParams.poly image (2 Cout columns)
cudaError_t launch_conv_tc(const TcConvParams& p, bool x3, cudaStream_t st);
extern int g_conv_tc_cluster;        // CTAs per cluster sharing weight stages (1, 2, 4) when g_conv_tc_pair == 0
extern int g_conv_tc_pair;           // cta_group::2 CTA pairs (default 1)
extern int g_conv_tc_bn256_stages;   // N=256 tiles from this many weight stages per tile
extern int g_conv_tc_epi8;           // two epilogue warpgroups (eegldm_set_conv_tuning bit 6 = off)
extern int g_conv_tc_cat;            // N=128 f16x3 tiles: hi x [hi | lo] as one N=256 MMA (eegldm_set_conv_tuning bit 5)
cudaError_t launch_groupnorm(const GnParams& p, cudaStream_t st);
cudaError_t launch_groupnorm_finalize(const GnParams& p, cudaStream_t st);   // statistics already in p.partial (p.nsplit records)
int groupnorm_nsplit(int C, int T, int G);
cudaError_t launch_attention_simt(const AttnParams& p, cudaStream_t st);
// ---- tcgen05 attention (attn_tc.cu) ----------------------------------------------------------
struct AttnTcParams {
    const uint8_t* qkv16;   // fp16 hi/lo images written by launch_qkv_split
    float* out;             // [B][T][H*ch] fp32
    int T, H, ch, B;
    float scale_log2e;      // ch^-1/2 * log2(e)
    uint8_t* out_u;         // optional (then `out` is unused): write the result as the U operand image of the following 1x1
                            // proj_out conv (conv_tc.cu layout, T % 16 == 0) instead of fp32
    const float* qkv32;     // optional (then `qkv16` is unused): fp32 qkv [B][T][H*3*ch]; q, k, v are split to fp16 hi/lo inside the
                            // kernel (attn_direct_eligible), no launch_qkv_split pass
    unsigned long long* timeline;   // optional (eegldm_bench_attention): 8 cycle counters per CTA, see attn_tc.cu
    // long sequences (T > 256: the raw-signal DM variant, sample_trials_ddpm.py, T = 768): keys are processed in blocks of Tk
    // (attn_tc_key_block) by separate CTAs; each writes its softmax-normalised partial output and the row's (max, sum) of its block,
    // launch_attention_tc then merges the blocks (flash-decoding style split over the keys).  Set by launch_attention_tc.
    int Tk;                 // keys per CTA (0 / T: all of them, no merge)
    float* part_out;        // [T/Tk][B][T][H*ch] fp32
    float2* part_ml;        // [T/Tk][B][H][T]: (row max in log2 units, row sum) of the block
};
bool attn_tc_eligible(int T, int ch);
int attn_tc_key_block(int T);                               // keys per CTA: T for T <= 256, else the largest of 256 / 192 / 128 dividing T
size_t attn_tc_scratch_bytes(int B, int T, int H, int ch);  // part_out + part_ml for T > 256 (0 otherwise); pass as AttnTcParams.part_out
bool attn_direct_eligible(int T, int ch);   // in-kernel fp32 -> fp16 hi/lo split of q, k, v (T <= 208)
size_t attn_qkv16_bytes(int B, int T, int H, int ch);
cudaError_t launch_qkv_split(const float* qkv, uint8_t* dst, int B, int T, int H, int ch, cudaStream_t st, int* range_flag = nullptr);
cudaError_t launch_attention_tc(const AttnTcParams& p, bool x3, cudaStream_t st);
// y[r][o] = bias[o] + sum_i act(x[r][i]) * W[o][i];  act = SiLU if silu_in
cudaError_t launch_linear(const float* x, const float* W, const float* bias, float* y, int R, int I, int O,
                          int silu_in, cudaStream_t st);
// timestep_embedding (unet.py:12-36) of device-resident timesteps: out[nt][dim], same rounding points as the host version
cudaError_t launch_timestep_embedding(const float* t_dev, int nt, int dim, float* out, cudaStream_t st);
// NCL <-> NLC transposes for multi-channel boundary tensors
cudaError_t launch_transpose_ncl_to_nlc(const float* in, float* out, int B, int C, int T, cudaStream_t st);
cudaError_t launch_transpose_nlc_to_ncl(const float* in, float* out, int B, int C, int T, cudaStream_t st);
// standalone Downsample/Upsample without conv (unet.py:195,221) on [B][Tin][C]
cudaError_t launch_resample(const float* in, float* out, int B, int Tin, int C, int mode, cudaStream_t st);
// dst[i] = src[i] * alpha
cudaError_t launch_scale(const float* src, float* dst, float alpha, size_t n, cudaStream_t st);
// copies row `*step` of the per-step tables into the "current" buffers and increments *step
cudaError_t launch_step_advance(const float* temb_table, int temb_row, float* temb_cur, const float* coef_table,
                                float* coef_cur, int* step, cudaStream_t st);
// AEKL sampling: z = mu + eps * sigma ; sigma = exp(clamp(logvar,-30,20)/2)
cudaError_t launch_kl_sigma(const float* logvar, float* sigma, size_t n, cudaStream_t st);
cudaError_t launch_axpy_sampling(const float* mu, const float* sigma, const float* eps, float* z, size_t n,
                                 cudaStream_t st);

int spectral_loss(const float* input, const float* target, int B, int N, int reduction, float loss_weight, float* loss_dev,
                  float* grad_dev, float grad_weight, int grad_accumulate, cudaStream_t st, std::string* err);
// ---- AutoencoderKL training step (train_kernels.cu, spectral.cu) --------------------------------
cudaError_t launch_norm_act_fwd(const float* x, const float* scale, const float* shift, float* a, int B, int T, int C, int silu,
                                cudaStream_t st);
struct ConvGradParams {   // y = conv(upsample?(a)): taps, stride, left pad; Tc = conv-input length, Tin = length of a
    const float* dy; const float* a; const float* w;   // w: SIMT image [(ci*taps + k)][Cout]
    float* da; float* dw; float* db;
    int Cin, Cout, taps, stride, pad, ups, Tin, Tc, Tout, B, accumulate;
    float* dres;           // optional: gradient slot of the conv's identity-residual input, dres (+)= dy (launch_conv_bwd only)
    int dres_accumulate;
};
cudaError_t launch_conv_bwd_data(const ConvGradParams& p, cudaStream_t st);
cudaError_t launch_conv_bwd_weight(const ConvGradParams& p, cudaStream_t st);   // dw, db are accumulated (atomicAdd)
cudaError_t launch_conv_bwd(const ConvGradParams& p, cudaStream_t st);          // both (da == null: weights only)
struct NormGradParams {
    const float* da; const float* x; const float* mean; const float* rstd; const float* gamma; const float* beta;
    float* m12; float* dgamma; float* dbeta; float* dx;
    int C, T, G, B, silu, accumulate;
};
cudaError_t launch_norm_act_bwd(const NormGradParams& p, cudaStream_t st);
cudaError_t launch_gn_act_fwd(const GnParams& p, float* a, int silu, cudaStream_t st);   // statistics + apply (+ SiLU), mean / rstd saved
cudaError_t launch_axpy(const float* src, float* dst, float alpha, int accumulate, size_t n, cudaStream_t st);
cudaError_t launch_l1_loss(const float* r, const float* x, float* dr, float* loss, float weight, size_t n, cudaStream_t st);
cudaError_t launch_latent(const float* mu, const float* lv, const float* eps, float* sigma, float* z, const float* dz, float* dmu,
                          float* dlv, float* kl_loss, float kl_weight, int B, size_t n, cudaStream_t st, const float* dmu_ext = nullptr,
                          const float* dsigma_ext = nullptr);
// ---- PatchDiscriminator / adversarial loss (train_kernels.cu): BatchNorm1d in training mode over channels-last rows [N][C]
cudaError_t launch_bn_stats(const float* h, size_t N, int C, const float* gamma, const float* beta, float eps, int B, double* sums /*[2C]*/,
                            float* scale /*[B][C]*/, float* shift, float* mean /*[C]*/, float* rstd, float* run_mean, float* run_var,
                            float momentum, int n_updates, cudaStream_t st);
cudaError_t launch_bn_lrelu_bwd(const float* da, const float* h, const float* mean, const float* rstd, const float* gamma,
                                const float* beta, size_t N, int C, float slope, double* sums, float* dh, float* dgamma /*nullable: +=*/,
                                float* dbeta, cudaStream_t st);
cudaError_t launch_lrelu_bwd(const float* da, const float* h, float slope, float* dh, size_t n, cudaStream_t st);
cudaError_t launch_affine_lrelu(const float* h, const float* scale /*[C] or null*/, const float* shift, int C, float slope, float* a,
                                size_t n, cudaStream_t st);
cudaError_t launch_adv_loss(const float* logits, size_t n, float target, float act_slope, float loss_weight, float* loss /*+=*/,
                            float grad_weight, float* dlogits, cudaStream_t st);
cudaError_t launch_dgrad_weights(const float* w, int Cin, int Cout, int stride, float* wd, cudaStream_t st);
// ---- tensor-pipe training pieces (train_tc.cu)
constexpr int WG_CHUNK = 32;   // positions per K chunk of the weight-gradient GEMM
size_t conv_tc_image_u16(int Cin, int Cout, int k);   // 16-bit elements of a tcgen05 weight image
cudaError_t launch_pack_conv_tc_dev(const float* w_simt, int Cin, int Cout, int k, uint16_t* out, cudaStream_t st);   // f16x3 image from a SIMT image
cudaError_t launch_s2z_weights(const float* w, int Cin, int Cout, float* wv, cudaStream_t st);          // stride-2 conv as a conv over row pairs
cudaError_t launch_s2z_grad_fold(const float* dwv, int Cin, int Cout, float* dw, cudaStream_t st);
cudaError_t launch_colsum(const float* x, size_t N, int C, float* out /*+=*/, cudaStream_t st);
bool wgrad_tc_eligible(int Cin, int Cout, int T);
size_t wgrad_image_bytes(size_t rows_total, int C, int halo);
cudaError_t launch_wgrad_split(const float* src, const float* scale, const float* shift, int ss_bstride, int act, int B, int T, int C,
                               int halo, uint8_t* img, cudaStream_t st);
cudaError_t launch_wgrad_tc(const uint8_t* a_img, const uint8_t* dy_img, float* dw /*[(ci*ktot+tap)][Cout], +=*/, int B, int T, int Cin,
                            int Cout, int ktot, int tap0, int ntaps, cudaStream_t st, int dw_tap_shift = 0);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, float lr, float b1, float b2, float eps, int step, size_t n,
                        cudaStream_t st);

// ---- latent-diffusion training step (unet_train_kernels.cu): everything around the convolutions
struct BGemm {   // C[m][n] (+)= alpha * sum_k A[m][k] B[k][n], element strides; batch index = outer * nb_inner + inner
    const float* A; const float* B; float* C;
    int M, N, K;
    long long a_m, a_k, b_k, b_n, c_m, c_n;
    int batch, nb_inner;
    long long a_bo, a_bi, b_bo, b_bi, c_bo, c_bi;
    float alpha; int accumulate;
};
cudaError_t launch_bgemm(const BGemm& g, cudaStream_t st);
cudaError_t launch_softmax_rows(float* s, size_t rows, int n, cudaStream_t st);                            // in place
cudaError_t launch_softmax_bwd_rows(const float* p, float* dp, size_t rows, int n, cudaStream_t st);       // dP -> dS in place
cudaError_t launch_rowsum_bt(const float* x, int B, int T, int C, float* out, int out_stride, int accumulate, cudaStream_t st);
cudaError_t launch_resample_bwd(const float* dy, float* dx, int B, int Tin, int C, int mode, int accumulate, cudaStream_t st);
cudaError_t launch_concat(const float* x0, int C0, const float* x1, int C1, float* out, size_t rows, cudaStream_t st);
cudaError_t launch_split_add(const float* dcat, float* d0, int C0, int acc0, float* d1, int C1, int acc1, size_t rows, cudaStream_t st);
cudaError_t launch_silu_fwd(const float* x, float* y, size_t n, cudaStream_t st);
cudaError_t launch_silu_bwd(const float* dy, const float* x, float* dx, size_t n, cudaStream_t st);
cudaError_t launch_ldm_inputs(const float* z0, const float* eps, const long long* t, const float* acp, int n_train, float* noisy, float* target,
                              float* t_f, int v_pred, int B, size_t per, cudaStream_t st);
cudaError_t launch_mse_loss(const float* p, const float* y, float* dp, float* loss, size_t n, cudaStream_t st);
cudaError_t launch_dgrad_weights_any(const float* w, int Cin, int Cout, int taps, float* wd, cudaStream_t st);

}  // namespace eegldm
