// Tensor-pipe pieces of the training steps (PatchDiscriminator, train_autoencoderkl.py:213-234; UNet backward,
// training/training.py:402-450): everything a convolution's backward pass needs on tcgen05 in the f16x3 arithmetic of conv_tc.cu.
//
//   data gradient     = the FORWARD kernel (conv_tc.cu) on transformed weights: launch_dgrad_weights (train_kernels.cu) builds them
//                       in the SIMT image, pack_conv_tc_dev (here) turns any SIMT image into the tcgen05 hi/lo stage image on the
//                       device (weights change every optimiser step, so the host packer of the inference path is not usable).
//   stride-2 forward  = a stride-1 conv over the input's row pairs seen as one row of 2 Cin channels (channels-last: the same
//                       memory): y[t] = [0 | W0] z[t-1] + [W1 | W2] z[t]   (s2z_weights_kernel builds that virtual 3-tap image).
//   weight gradient   = wgrad_tc_kernel: dW[tap][ci][co] = sum_{b,t} a[b, t + tap - 1, ci] * dy[b, t, co], a GEMM whose reduction runs
//                       over the B*T positions.  Both operands are "MN-major" for the tensor core (channels contiguous); a pre-pass
//                       (wgrad_split_kernel) splits them to fp16 hi/lo once and stores them as [chunk][hi|lo][C/8][rows][8 ch]: per
//                       8-channel core a contiguous column of position rows, 16 bytes each.  A tap shift is then +16 bytes on the
//                       descriptor start address (rows are linear in K across core matrices because LBO = 128 = 8 rows), so the 3
//                       taps read ONE staged tile with a 2-row halo, like the forward kernel's phase-strided layout does along M.
//                       CTA = (tap, 128 input channels, BNW output channels, K split); 4 stages of 32 positions fed by bulk copies;
//                       f16x3: hi*hi -> accumulator 0, hi*lo + lo*hi -> accumulator 1; epilogue adds the tile into dW with fp32
//                       reductions (split-K).
#include <cuda_fp16.h>

#include <algorithm>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace eegldm {
using namespace tc;
namespace {

constexpr int CH = WG_CHUNK;          // positions per K chunk
constexpr int RA = CH + 2;            // rows of an activation column (1 halo row each side)
constexpr int WG_STAGES = 4;
constexpr int WG_THREADS = 192;

// ------------------------------------------------------------------------------------------------ weight images on the device
// SIMT image w[(ci*k + tap)][Cout] fp32  ->  tcgen05 image [k-step][tap][hi|lo][Cout/8][kc 4][8][8] fp16 (pack_conv_tc's layout)
__global__ void pack_conv_tc_dev_kernel(const float* __restrict__ w, int Cin, int Cout, int k, uint16_t* __restrict__ out) {
    const size_t total = (size_t)(Cin / TC_BK) * k * Cout * 4;        // one thread per (k-step, tap, co, kc): 8 channels
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int co = (int)(i % Cout);
    size_t rest = i / Cout;
    const int kc = (int)(rest % 4); rest /= 4;
    const int tap = (int)(rest % k);
    const int ks = (int)(rest / k);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = w[((size_t)(ks * TC_BK + kc * 8 + e) * k + tap) * Cout + co];
    uint4 hi, lo;
    split8_f16(v, hi, lo);
    const size_t half = (size_t)Cout * TC_BK;                         // u16 elements per half
    uint16_t* h = out + ((size_t)ks * k + tap) * 2 * half + (size_t)(co >> 3) * 256 + kc * 64 + (co & 7) * 8;
    *reinterpret_cast<uint4*>(h) = hi;
    *reinterpret_cast<uint4*>(h + half) = lo;
}

// virtual forward weights of a k3 / stride 2 / padding 1 conv over row pairs: Wv[(c2*3 + tap)][Cout], c2 in [0, 2 Cin)
//   tap 0 (z[t-1]): c2 >= Cin -> W[c2-Cin][0];   tap 1 (z[t]): c2 < Cin -> W[c2][1], else W[c2-Cin][2];   tap 2: 0
__global__ void s2z_weights_kernel(const float* __restrict__ w, int Cin, int Cout, float* __restrict__ wv) {
    const size_t total = (size_t)2 * Cin * 3 * Cout;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int co = (int)(i % Cout);
    const int tap = (int)((i / Cout) % 3);
    const int c2 = (int)(i / ((size_t)Cout * 3));
    float v = 0.f;
    if (tap == 0) { if (c2 >= Cin) v = w[(size_t)((c2 - Cin) * 3 + 0) * Cout + co]; }
    else if (tap == 1) v = c2 < Cin ? w[(size_t)(c2 * 3 + 1) * Cout + co] : w[(size_t)((c2 - Cin) * 3 + 2) * Cout + co];
    wv[i] = v;
}
// fold the gradient of the virtual weights back: dW[ci][0] += dWv[ci+Cin][0], dW[ci][1] += dWv[ci][1], dW[ci][2] += dWv[ci+Cin][1]
__global__ void s2z_grad_fold_kernel(const float* __restrict__ dwv, int Cin, int Cout, float* __restrict__ dw) {
    const size_t total = (size_t)Cin * 3 * Cout;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int co = (int)(i % Cout);
    const int k = (int)((i / Cout) % 3);
    const int ci = (int)(i / ((size_t)Cout * 3));
    const int c2 = k == 1 ? ci : ci + Cin, tap = k == 0 ? 0 : 1;
    dw[i] += dwv[(size_t)(c2 * 3 + tap) * Cout + co];
}

// ------------------------------------------------------------------------------------------------ operand pre-pass of the weight gradient
// src [B][T][C] fp32 (optionally act(scale[c] * x + shift[c])) -> img[chunk][hi|lo][C/8][R][8] fp16, R = CH + 2 with a halo row on
// each side (zero at the sample edges = the conv padding) or R = CH.  One thread per (chunk, 8-channel core, row).
__global__ void __launch_bounds__(256) wgrad_split_kernel(const float* __restrict__ src, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, int ss_bstride, int act, int T, int C,
                                                           int halo, size_t nchunks, uint8_t* __restrict__ img) {
    const int R = halo ? RA : CH, ncore = C / 8;
    const size_t total = nchunks * ncore * R;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r = (int)(i % R);
    const int core = (int)((i / R) % ncore);
    const size_t q = i / ((size_t)R * ncore);
    const int cpt = T / CH;                                   // chunks per sample
    const size_t b = q / cpt;
    const int t = (int)(q % cpt) * CH + r - (halo ? 1 : 0);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (t >= 0 && t < T) {
        const float* p = src + ((size_t)b * T + t) * C + core * 8;
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(p)), x1 = __ldg(reinterpret_cast<const float4*>(p + 4));
        v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
        if (scale) {   // per-channel (BatchNorm: ss_bstride = 0) or per-(sample, channel) (GroupNorm: ss_bstride = C) affine
            const size_t so = b * (size_t)ss_bstride + core * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = fmaf(scale[so + e], v[e], shift[so + e]);
        }
        if (act == 2) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = v[e] > 0.f ? v[e] : 0.2f * v[e];
        } else if (act == 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = v[e] / (1.f + __expf(-v[e]));
        }
    }
    uint4 hi, lo;
    split8_f16(v, hi, lo);
    const size_t half = (size_t)ncore * R * 16;
    uint8_t* o = img + q * 2 * half + ((size_t)core * R + r) * 16;
    *reinterpret_cast<uint4*>(o) = hi;
    *reinterpret_cast<uint4*>(o + half) = lo;
}
// bias gradient: db[c] += sum over rows of dy [N][C]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, size_t N, int C, float* __restrict__ out) {
    const size_t r0 = (size_t)blockIdx.x * 128, r1 = min(N, r0 + 128);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (size_t r = r0; r < r1; ++r) s += x[r * C + c];
        atomicAdd(out + c, s);
    }
}

// ------------------------------------------------------------------------------------------------ weight-gradient GEMM
struct WgradParams {
    const uint8_t* A;   // activations  [chunk][hi|lo][Cin/8][RA][8]
    const uint8_t* Bm;  // dy           [chunk][hi|lo][Cout/8][CH][8]
    float* dw;          // [(ci*ktot + tap)][Cout] fp32, += (split-K reductions)
    int Cin, Cout, ntaps, tap0, ktot;   // taps tap0 .. tap0+ntaps-1 of a ktot-tap weight image (row offset of tap j: j - 1 + ... see below)
    int nchunks, nsplit;
    int dw_tap_shift;   // tap index in dW = tap - dw_tap_shift (a 1x1 conv reads halo row 1 = "tap 1" and has the single tap 0)
};

template <int BNW>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const WgradParams p) {
    constexpr int A_HALF = 16 * RA * 16;            // 128 input channels: 16 cores x RA rows x 16 B
    constexpr int B_HALF = (BNW / 8) * CH * 16;
    constexpr int STAGE = 2 * A_HALF + 2 * B_HALF;
    static_assert(WG_STAGES * STAGE + 256 <= 227 * 1024, "stage ring exceeds shared memory");
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + WG_STAGES * STAGE;
    const uint32_t barFull = bars, barEmpty = bars + 8 * WG_STAGES, barAcc = bars + 16 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WG_STAGES * STAGE + 16 * WG_STAGES + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // work unit: (tap, ci tile, co tile) x K split
    const int n_ci = p.Cin / 128, n_co = p.Cout / BNW;
    int u = blockIdx.x / p.nsplit;
    const int split = blockIdx.x % p.nsplit;
    const int co_t = u % n_co; u /= n_co;
    const int ci_t = u % n_ci; u /= n_ci;
    const int tap = p.tap0 + u;                     // tap index in the weight image: reads rows (position + tap - 1) -> halo row offset tap
    const int c0 = (int)((long long)p.nchunks * split / p.nsplit), c1 = (int)((long long)p.nchunks * (split + 1) / p.nsplit);
    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(barFull + 8 * i, 1); mbar_init(barEmpty + 8 * i, 1); }
        mbar_init(barAcc, 1);
        fence_mbar_init();
    }
    constexpr uint32_t TCOLS = 2 * BNW;             // accumulator 0 | accumulator 1
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const size_t a_chunk = (size_t)2 * (p.Cin / 8) * RA * 16, a_half_g = (size_t)(p.Cin / 8) * RA * 16;
    const size_t b_chunk = (size_t)2 * (p.Cout / 8) * CH * 16, b_half_g = (size_t)(p.Cout / 8) * CH * 16;
    if (warp == 4) {
        // ================================================================ loader: 4 bulk copies per stage
        int it = 0;
        for (int c = c0; c < c1; ++c, ++it) {
            const int s = it % WG_STAGES;
            mbar_wait(barEmpty + 8 * s, ((it / WG_STAGES) & 1) ^ 1);
            if (elect_one()) {
                const uint32_t dst = sbase + s * STAGE, bar = barFull + 8 * s;
                mbar_arrive_expect_tx(bar, STAGE);
                const uint8_t* a = p.A + (size_t)c * a_chunk + (size_t)ci_t * A_HALF;
                const uint8_t* b = p.Bm + (size_t)c * b_chunk + (size_t)co_t * B_HALF;
                bulk_copy_g2s(dst, a, A_HALF, bar);
                bulk_copy_g2s(dst + A_HALF, a + a_half_g, A_HALF, bar);
                bulk_copy_g2s(dst + 2 * A_HALF, b, B_HALF, bar);
                bulk_copy_g2s(dst + 2 * A_HALF + B_HALF, b + b_half_g, B_HALF, bar);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ================================================================ MMA issuer
        // both operands MN-major (bits 15 / 16 of the instruction descriptor): core matrix = 8 K rows x 16 B of MN
        constexpr uint32_t IDESC = make_idesc(0u, 128u, (uint32_t)BNW, 1u) | (1u << 15);
        int it = 0;
        uint32_t acc = 0;
        for (int c = c0; c < c1; ++c, ++it) {
            const int s = it % WG_STAGES;
            mbar_wait(barFull + 8 * s, (it / WG_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_hi = sbase + s * STAGE + tap * 16, a_lo = a_hi + A_HALF;
                const uint32_t b_hi = sbase + s * STAGE + 2 * A_HALF, b_lo = b_hi + B_HALF;
#pragma unroll
                for (int kk = 0; kk < CH / 16; ++kk) {
                    const uint64_t dah = make_desc(a_hi + kk * 256, 128, RA * 16), dal = make_desc(a_lo + kk * 256, 128, RA * 16);
                    const uint64_t dbh = make_desc(b_hi + kk * 256, 128, CH * 16), dbl = make_desc(b_lo + kk * 256, 128, CH * 16);
                    umma_bf16(tmem, dah, dbh, IDESC, acc);
                    umma_bf16(tmem + BNW, dah, dbl, IDESC, acc);
                    umma_bf16(tmem + BNW, dal, dbh, IDESC, 1u);
                    acc = 1;
                }
                umma_commit(barEmpty + 8 * s);
                if (c == c1 - 1) umma_commit(barAcc);
            }
            __syncwarp();
        }
    } else if (c1 > c0) {
        // ================================================================ epilogue: accumulators -> fp32 reductions into dW
        mbar_wait(barAcc, 0);
        tc_fence_after();
        const int ci = ci_t * 128 + warp * 32 + lane;                    // TMEM lane = input channel
        float* row = p.dw + ((size_t)ci * p.ktot + (tap - p.dw_tap_shift)) * p.Cout + (size_t)co_t * BNW;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int cb = 0; cb < BNW; cb += 32) {
            uint32_t v[32], c2[32];
            tmem_ld32_async(taddr + (uint32_t)cb, v);
            tmem_ld32_async(taddr + (uint32_t)(BNW + cb), c2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(row + cb + j, fmaf(__uint_as_float(c2[j]), 1.0f / LO_SCALE, __uint_as_float(v[j])));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, TCOLS);
}

unsigned nblk(size_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

size_t conv_tc_image_u16(int Cin, int Cout, int k) { return (size_t)(Cin / TC_BK) * k * 2 * Cout * TC_BK; }

cudaError_t launch_pack_conv_tc_dev(const float* w_simt, int Cin, int Cout, int k, uint16_t* out, cudaStream_t st) {
    const size_t total = (size_t)(Cin / TC_BK) * k * Cout * 4;
    if (!total) return cudaSuccess;
    pack_conv_tc_dev_kernel<<<nblk(total), 256, 0, st>>>(w_simt, Cin, Cout, k, out);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_s2z_weights(const float* w, int Cin, int Cout, float* wv, cudaStream_t st) {
    s2z_weights_kernel<<<nblk((size_t)2 * Cin * 3 * Cout), 256, 0, st>>>(w, Cin, Cout, wv);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_s2z_grad_fold(const float* dwv, int Cin, int Cout, float* dw, cudaStream_t st) {
    s2z_grad_fold_kernel<<<nblk((size_t)Cin * 3 * Cout), 256, 0, st>>>(dwv, Cin, Cout, dw);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_colsum(const float* x, size_t N, int C, float* out, cudaStream_t st) {
    if (!N) return cudaSuccess;
    colsum_kernel<<<(unsigned)((N + 127) / 128), 256, 0, st>>>(x, N, C, out);
    g_launch_count += 1;
    return cudaGetLastError();
}

bool wgrad_tc_eligible(int Cin, int Cout, int T) { return Cin > 0 && Cin % 128 == 0 && Cout > 0 && Cout % 128 == 0 && T > 0 && T % CH == 0; }
size_t wgrad_image_bytes(size_t rows_total, int C, int halo) { return rows_total / CH * 2 * (size_t)(C / 8) * (halo ? RA : CH) * 16; }

cudaError_t launch_wgrad_split(const float* src, const float* scale, const float* shift, int ss_bstride, int act, int B, int T, int C,
                               int halo, uint8_t* img, cudaStream_t st) {
    const size_t nchunks = (size_t)B * T / CH;
    const size_t total = nchunks * (C / 8) * (halo ? RA : CH);
    if (!total) return cudaSuccess;
    wgrad_split_kernel<<<nblk(total), 256, 0, st>>>(src, scale, shift, ss_bstride, act, T, C, halo, nchunks, img);
    g_launch_count += 1;
    return cudaGetLastError();
}

// dw[(ci*ktot + tap)][Cout] += sum over positions, taps tap0 .. tap0 + ntaps - 1 (tap j reads position t + j - 1)
cudaError_t launch_wgrad_tc(const uint8_t* a_img, const uint8_t* dy_img, float* dw, int B, int T, int Cin, int Cout, int ktot, int tap0,
                            int ntaps, cudaStream_t st, int dw_tap_shift) {
    if (!wgrad_tc_eligible(Cin, Cout, T) || ntaps < 1 || tap0 < 0 || tap0 + ntaps > 3) return cudaErrorInvalidValue;
    WgradParams p{a_img, dy_img, dw, Cin, Cout, ntaps, tap0, ktot, (int)((size_t)B * T / CH), 1, dw_tap_shift};
    const int bnw = Cout % 256 == 0 ? 256 : 128;
    const int units = ntaps * (Cin / 128) * (Cout / bnw);
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    p.nsplit = std::max(1, std::min(p.nchunks, num_sms / units));
    const size_t smem_a = (size_t)2 * 16 * RA * 16;
    cudaError_t e;
    if (bnw == 256) {
        const size_t smem = WG_STAGES * (smem_a + (size_t)2 * 32 * CH * 16) + 256;
        static bool set = false;
        if (!set) { e = cudaFuncSetAttribute(wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; set = true; }
        wgrad_tc_kernel<256><<<units * p.nsplit, WG_THREADS, smem, st>>>(p);
    } else {
        const size_t smem = WG_STAGES * (smem_a + (size_t)2 * 16 * CH * 16) + 256;
        static bool set = false;
        if (!set) { e = cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; set = true; }
        wgrad_tc_kernel<128><<<units * p.nsplit, WG_THREADS, smem, st>>>(p);
    }
    g_launch_count += 1;
    return cudaGetLastError();
}

}  // namespace eegldm
