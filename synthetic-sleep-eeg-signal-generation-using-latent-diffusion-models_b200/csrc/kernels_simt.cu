// fp32 SIMT kernels of the eegldm engine: the exact-arithmetic ("parity") path and every op that is
// too narrow for the tensor pipe (AEKL 2-2-4 convs, GroupNorm statistics, softmax, time MLP).
//
// Reference semantics (file:line under /root/reference):
//   fused conv prologue  = Normalize -> SiLU -> {AvgPool1d | nearest x2}   src/models/unet.py:71-74,262,308-313
//   conv                 = nn.Conv1d k in {1,3}                            unet.py:263,291,302,385,504
//   epilogue             = + emb_out[..., None] (unet.py:325), skip(x)+h (unet.py:327), x + proj (unet.py:174)
//   attention            = QKVAttentionLegacy.forward                      unet.py:107-125
#include "kernels.cuh"

#include <algorithm>

namespace eegldm {
namespace {

constexpr int BN = 128;        // output positions per CTA
constexpr int KC = 16;         // input channels per K-chunk
constexpr int XP = 2 * BN + 4; // smem pitch of the activation tile (covers the stride-2 case)

__device__ __forceinline__ float silu_f(float v) { return v / (1.f + expf(-v)); }

// prologue activation after the affine: 0 none, 1 SiLU (UNet / autoencoder), 2 LeakyReLU(0.2) (PatchDiscriminator)
__device__ __forceinline__ float act1(float x, float a, float s, int silu) {
    float v = fmaf(a, x, s);
    return silu == 1 ? silu_f(v) : (silu == 2 ? (v > 0.f ? v : 0.2f * v) : v);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------------------------------------
// Activation tile loader: Xs[kk][x] = u[b, u0+x, c0+kk], zero outside [0,Tc) and beyond Cin.
__device__ __forceinline__ void load_x_tile(const ConvSeg& sg, int b, int c0, int u0, int ncols, int Tc,
                                            float (*Xs)[XP]) {
    const int Cin = sg.C0 + sg.C1;
    const bool vec = ((sg.C0 & 3) == 0) && ((sg.C1 & 3) == 0);
    const bool aff = sg.scale != nullptr;
    if (vec) {
        for (int idx = threadIdx.x; idx < ncols * 4; idx += 256) {
            const int x = idx >> 2, cq = idx & 3;
            const int c = c0 + cq * 4;
            const int u = u0 + x;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < Cin && u >= 0 && u < Tc) {
                const float* src; int cc, Cs;
                if (c < sg.C0) { src = sg.src0; cc = c; Cs = sg.C0; }
                else           { src = sg.src1; cc = c - sg.C0; Cs = sg.C1; }
                float4 a4 = make_float4(1.f, 1.f, 1.f, 1.f), s4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (aff) {
                    a4 = ld4(sg.scale + (size_t)b * Cin + c);
                    s4 = ld4(sg.shift + (size_t)b * Cin + c);
                }
                const float* base = src + (size_t)b * sg.Tin * Cs + cc;
                if (sg.resample == RS_NONE) {
                    float4 r = ld4(base + (size_t)u * Cs);
                    v.x = act1(r.x, a4.x, s4.x, sg.silu); v.y = act1(r.y, a4.y, s4.y, sg.silu);
                    v.z = act1(r.z, a4.z, s4.z, sg.silu); v.w = act1(r.w, a4.w, s4.w, sg.silu);
                } else if (sg.resample == RS_AVGPOOL2) {
                    float4 r0 = ld4(base + (size_t)(2 * u) * Cs);
                    float4 r1 = ld4(base + (size_t)(2 * u + 1) * Cs);
                    v.x = 0.5f * (act1(r0.x, a4.x, s4.x, sg.silu) + act1(r1.x, a4.x, s4.x, sg.silu));
                    v.y = 0.5f * (act1(r0.y, a4.y, s4.y, sg.silu) + act1(r1.y, a4.y, s4.y, sg.silu));
                    v.z = 0.5f * (act1(r0.z, a4.z, s4.z, sg.silu) + act1(r1.z, a4.z, s4.z, sg.silu));
                    v.w = 0.5f * (act1(r0.w, a4.w, s4.w, sg.silu) + act1(r1.w, a4.w, s4.w, sg.silu));
                } else {
                    float4 r = ld4(base + (size_t)(u >> 1) * Cs);
                    v.x = act1(r.x, a4.x, s4.x, sg.silu); v.y = act1(r.y, a4.y, s4.y, sg.silu);
                    v.z = act1(r.z, a4.z, s4.z, sg.silu); v.w = act1(r.w, a4.w, s4.w, sg.silu);
                }
            }
            Xs[cq * 4 + 0][x] = v.x; Xs[cq * 4 + 1][x] = v.y;
            Xs[cq * 4 + 2][x] = v.z; Xs[cq * 4 + 3][x] = v.w;
        }
    } else {
        for (int idx = threadIdx.x; idx < ncols * KC; idx += 256) {
            const int x = idx / KC, kk = idx % KC;
            const int c = c0 + kk;
            const int u = u0 + x;
            float v = 0.f;
            if (c < Cin && u >= 0 && u < Tc) {
                const float* src; int cc, Cs;
                if (c < sg.C0) { src = sg.src0; cc = c; Cs = sg.C0; }
                else           { src = sg.src1; cc = c - sg.C0; Cs = sg.C1; }
                float a = 1.f, s = 0.f;
                if (aff) { a = sg.scale[(size_t)b * Cin + c]; s = sg.shift[(size_t)b * Cin + c]; }
                const float* base = src + (size_t)b * sg.Tin * Cs + cc;
                if (sg.resample == RS_NONE) v = act1(base[(size_t)u * Cs], a, s, sg.silu);
                else if (sg.resample == RS_AVGPOOL2)
                    v = 0.5f * (act1(base[(size_t)(2 * u) * Cs], a, s, sg.silu) +
                                act1(base[(size_t)(2 * u + 1) * Cs], a, s, sg.silu));
                else v = act1(base[(size_t)(u >> 1) * Cs], a, s, sg.silu);
            }
            Xs[kk][x] = v;
        }
    }
}

template <int BM, int TAPS>
__device__ __forceinline__ void load_w_tile(const float* __restrict__ w, int Cin, int Cout, int c0, int co0,
                                            float (*Ws)[BM]) {
    constexpr int ROWS = KC * TAPS;
    const int grow0 = c0 * TAPS;
    const int grows = Cin * TAPS;
    if ((Cout & 3) == 0) {
        constexpr int QPR = BM / 4;
        for (int idx = threadIdx.x; idx < ROWS * QPR; idx += 256) {
            const int r = idx / QPR, q = idx % QPR;
            const int co = co0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow0 + r < grows && co < Cout) v = ld4(w + (size_t)(grow0 + r) * Cout + co);
            *reinterpret_cast<float4*>(&Ws[r][q * 4]) = v;
        }
    } else {
        for (int idx = threadIdx.x; idx < ROWS * BM; idx += 256) {
            const int r = idx / BM, m = idx % BM;
            const int co = co0 + m;
            float v = 0.f;
            if (grow0 + r < grows && co < Cout) v = w[(size_t)(grow0 + r) * Cout + co];
            Ws[r][m] = v;
        }
    }
}

// Accumulate one K-segment.  acc[i][j]: i -> output channel m_i, j -> position n_j with
//   n_j = tx*4 + (j&3) + 64*(j>>2);  m_i = ty*4 + (i&3) + 64*(i>>2)  (TM=8) | ty*4+i (TM=4) | ty (TM=1)
template <int TM, int TAPS>
__device__ __forceinline__ void conv_segment(const ConvSeg& sg, const ConvParams& p, int b, int t0, int co0,
                                             int stride, int pad_left, float (&acc)[TM][8],
                                             float (*Wsraw)[16 * TM], float (*Xs)[XP]) {
    constexpr int BM = 16 * TM;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int Cin = sg.C0 + sg.C1;
    const int ncols = (stride == 1) ? (BN + TAPS - 1) : (2 * BN + 1);
    const int u0 = t0 * stride - pad_left;
    for (int c0 = 0; c0 < Cin; c0 += KC) {
        __syncthreads();
        load_x_tile(sg, b, c0, u0, ncols, p.Tc, Xs);
        load_w_tile<BM, TAPS>(sg.w, Cin, p.Cout, c0, co0, Wsraw);
        __syncthreads();
        if (stride == 1) {
#pragma unroll 4
            for (int kk = 0; kk < KC; ++kk) {
                float xv[2][6];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float* xr = &Xs[kk][h * 64 + tx * 4];
                    float4 x0 = *reinterpret_cast<const float4*>(xr);
                    xv[h][0] = x0.x; xv[h][1] = x0.y; xv[h][2] = x0.z; xv[h][3] = x0.w;
                    if (TAPS == 3) {
                        float2 x1 = *reinterpret_cast<const float2*>(xr + 4);
                        xv[h][4] = x1.x; xv[h][5] = x1.y;
                    }
                }
#pragma unroll
                for (int k = 0; k < TAPS; ++k) {
                    float a[TM];
                    const float* wr = &Wsraw[kk * TAPS + k][0];
                    if constexpr (TM == 8) {
                        float4 a0 = *reinterpret_cast<const float4*>(wr + ty * 4);
                        float4 a1 = *reinterpret_cast<const float4*>(wr + 64 + ty * 4);
                        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                    } else if constexpr (TM == 4) {
                        float4 a0 = *reinterpret_cast<const float4*>(wr + ty * 4);
                        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                    } else {
                        a[0] = wr[ty];
                    }
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], xv[j >> 2][(j & 3) + k], acc[i][j]);
                }
            }
        } else {  // stride 2 (AEKL downsample, UNet Downsample(use_conv=True)): generic indexed reads
            for (int kk = 0; kk < KC; ++kk) {
#pragma unroll
                for (int k = 0; k < TAPS; ++k) {
                    float a[TM];
                    const float* wr = &Wsraw[kk * TAPS + k][0];
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        const int m = (TM == 8) ? (ty * 4 + (i & 3) + 64 * (i >> 2)) : (TM == 4 ? ty * 4 + i : ty);
                        a[i] = wr[m];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int n = tx * 4 + (j & 3) + 64 * (j >> 2);
                        const float xvv = Xs[kk][2 * n + k];
#pragma unroll
                        for (int i = 0; i < TM; ++i) acc[i][j] = fmaf(a[i], xvv, acc[i][j]);
                    }
                }
            }
        }
    }
}

template <int TM>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
    constexpr int BM = 16 * TM;
    __shared__ __align__(16) float Ws[KC * 3][BM];
    __shared__ __align__(16) float Xs[KC][XP];
    const int b = blockIdx.z, t0 = blockIdx.x * BN, co0 = blockIdx.y * BM;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    float acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    if (p.seg[0].taps == 3) conv_segment<TM, 3>(p.seg[0], p, b, t0, co0, p.stride, p.pad_left, acc, Ws, Xs);
    else                    conv_segment<TM, 1>(p.seg[0], p, b, t0, co0, p.stride, p.pad_left, acc, Ws, Xs);
    if (p.nseg > 1)         conv_segment<TM, 1>(p.seg[1], p, b, t0, co0, 1, 0, acc, Ws, Xs);

    // ---- epilogue: bias + temb + residual (+ DDIM update), channels-last store
    const int Cout = p.Cout;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int t = t0 + tx * 4 + (j & 3) + 64 * (j >> 2);
        if (t >= p.Tout) continue;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = (TM == 8) ? (ty * 4 + (i & 3) + 64 * (i >> 2)) : (TM == 4 ? ty * 4 + i : ty);
            const int co = co0 + m;
            if (co >= Cout) continue;
            float v = acc[i][j];
            if (p.bias) v += p.bias[co];
            if (p.temb) v += p.temb[(size_t)b * p.temb_stride + co];
            if (p.res) {
                const float* rb = p.res + (size_t)b * p.res_Tin * Cout + co;
                if (p.res_mode == RS_NONE) v += rb[(size_t)t * Cout];
                else if (p.res_mode == RS_AVGPOOL2)
                    v += 0.5f * (rb[(size_t)(2 * t) * Cout] + rb[(size_t)(2 * t + 1) * Cout]);
                else v += rb[(size_t)(t >> 1) * Cout];
            }
            const size_t o = ((size_t)b * p.Tout + t) * Cout + co;
            if (p.ddim_x) v = p.ddim_coef[0] * p.ddim_x[o] + p.ddim_coef[1] * v;
            p.out[o] = v;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Narrow-input conv (UNet input conv, unet.py:385: Conv1d(z -> model_channels, 3, padding=1), z <= 4): pure HBM-write bound.
// One thread = one position x 4 output channels; weights/bias in shared memory.
__global__ void __launch_bounds__(256) conv_narrow_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out, int Cin, int Cout,
                                                              int T, size_t total4) {
    extern __shared__ float ws[];   // [Cin*3][Cout] + [Cout]
    for (int i = threadIdx.x; i < Cin * 3 * Cout; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) ws[Cin * 3 * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    // thread = fixed group of 4 output channels, positions strided over the grid (no 64-bit divisions in the loop);
    // requires (Cout/4) | blockDim.x, checked by the launcher
    const int q = Cout >> 2;
    const int co = (int)(threadIdx.x % q) * 4;
    const unsigned ppb = blockDim.x / q;                       // positions per block per iteration
    const unsigned npos = (unsigned)(total4 / q);
    for (unsigned pos = blockIdx.x * ppb + threadIdx.x / q; pos < npos; pos += gridDim.x * ppb) {
        const size_t bt = pos;
        const int t = (int)(pos % (unsigned)T);
        float4 acc = *reinterpret_cast<const float4*>(ws + Cin * 3 * Cout + co);
        for (int k = 0; k < 3; ++k) {
            const int u = t + k - 1;
            if (u < 0 || u >= T) continue;
            const float* xr = x + (bt + k - 1) * Cin;
            for (int ci = 0; ci < Cin; ++ci) {
                const float xv = __ldg(xr + ci);
                const float4 wv = *reinterpret_cast<const float4*>(ws + (ci * 3 + k) * Cout + co);
                acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
            }
        }
        *reinterpret_cast<float4*>(out + bt * Cout + co) = acc;
    }
}

// The same conv for Cout = 128 emitting the GroupNorm records of its output (the layout the tcgen05 conv epilogue writes: one
// (count, mean, M2) per sample, 16-position segment and 4-channel group), so that input_blocks.1's norm and the last output block's
// concat norm need no pass over the tensor.  One warp = one segment: lane = 4-channel group, 16 positions in sequence (512-byte
// row stores), shifted one-pass sums per lane -- no cross-lane combination at all.
__global__ void __launch_bounds__(256) conv_narrow_in_gn_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ out,
                                                                 float* __restrict__ rec, int Cin, int T, int nsegs, int per) {
    constexpr int Cout = 128;
    extern __shared__ float ws[];   // [Cin*3][Cout] + [Cout]
    for (int i = threadIdx.x; i < Cin * 3 * Cout; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) ws[Cin * 3 * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, co = lane * 4;
    const int wpb = blockDim.x >> 5;
    const float4 b4 = *reinterpret_cast<const float4*>(ws + Cin * 3 * Cout + co);
    // per = 16 (a record per 16-position segment) or 128 (T % 128 == 0: a record per 128 positions); nsegs counts records
    for (int seg = blockIdx.x * wpb + (threadIdx.x >> 5); seg < nsegs; seg += gridDim.x * wpb) {
        const size_t bt0 = (size_t)seg * per;
        const int t0 = (int)(bt0 % (size_t)T);                 // T % per == 0: a record never straddles samples
        float K = 0.f, s = 0.f, q = 0.f;
#pragma unroll 4
        for (int i = 0; i < per; ++i) {
            float4 acc = b4;
            for (int k = 0; k < 3; ++k) {
                const int u = t0 + i + k - 1;
                if (u < 0 || u >= T) continue;
                const float* xr = x + (bt0 + i + k - 1) * Cin;
                for (int ci = 0; ci < Cin; ++ci) {
                    const float xv = __ldg(xr + ci);
                    const float4 wv = *reinterpret_cast<const float4*>(ws + (ci * 3 + k) * Cout + co);
                    acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
                }
            }
            *reinterpret_cast<float4*>(out + (bt0 + i) * Cout + co) = acc;
            if (i == 0) K = acc.x;
            const float d0 = acc.x - K, d1 = acc.y - K, d2 = acc.z - K, d3 = acc.w - K;
            s += (d0 + d1) + (d2 + d3);
            q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, q))));
        }
        float* o = rec + ((size_t)seg * 32 + lane) * 3;
        const float cnt = 4.f * per, icnt = 1.f / cnt;
        o[0] = cnt; o[1] = fmaf(s, icnt, K); o[2] = fmaxf(q - s * s * icnt, 0.f);
    }
}

// Narrow-output conv (UNet output head, unet.py:501-505: Conv1d(model_channels -> z, 3, padding=1) on SiLU(GN(h)), z <= 4)
// with the DDIM update in the epilogue: pure HBM-read bound.  One warp walks a strip of positions; each lane owns 4 of every
// 128 input channels, applies the GroupNorm affine + SiLU once per element and scatters it into the 3 output rows it feeds;
// a rolling window of 3 partial rows is warp-reduced as each row completes.
constexpr int NO_STRIP = 32;
template <int CO>
__global__ void __launch_bounds__(256) conv_narrow_out_kernel(const ConvParams p) {
    const ConvSeg& sg = p.seg[0];
    const int Cin = sg.C0, T = p.Tout;
    const int lane = threadIdx.x & 31;
    const int strips = (T + NO_STRIP - 1) / NO_STRIP;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)p.B * strips) return;
    const int b = (int)(wid / strips), t0 = (int)(wid % strips) * NO_STRIP;
    const int t1 = min(T, t0 + NO_STRIP);
    float acc[3][CO];   // partial sums of output rows r-1, r, r+1 (rolling)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
    const float* hb = sg.src0 + (size_t)b * T * Cin;
    // Cin == 128 (the UNet head): this lane's 4 channels never change -> affine and the 12*CO weights live in registers
    const bool hoist = Cin == 128;
    float4 ha = make_float4(1.f, 1.f, 1.f, 1.f), hs = make_float4(0.f, 0.f, 0.f, 0.f);
    float hw[4][3][CO];
    if (hoist) {
        if (sg.scale) { ha = ld4(sg.scale + (size_t)b * Cin + lane * 4); hs = ld4(sg.shift + (size_t)b * Cin + lane * 4); }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int c = 0; c < CO; ++c) hw[e][k][c] = c < p.Cout ? __ldg(sg.w + (size_t)((lane * 4 + e) * 3 + k) * p.Cout + c) : 0.f;
    }
    // (hoisted form: four rows of loads in flight ahead of the row being reduced -- the loop is a chain of load -> SiLU -> 5-step
    // shuffle reduction per row otherwise, 0.28 ms for a 403 MB read at B = 1024)
    auto ldrow = [&](int rr) -> float4 {
        return hoist && rr >= 0 && rr < T && rr <= t1 ? ld4(hb + (size_t)rr * Cin + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 pf0 = ldrow(t0 - 1), pf1 = ldrow(t0), pf2 = ldrow(t0 + 1), pf3 = ldrow(t0 + 2);
    for (int r = t0 - 1; r <= t1; ++r) {   // input rows feeding outputs t0..t1-1
        const float4 hv_pf = pf0;
        pf0 = pf1; pf1 = pf2; pf2 = pf3; pf3 = ldrow(r + 4);
        if (r >= 0 && r < T) {
            if (hoist) {
                const float4 hv = hv_pf;
                float v[4] = {hv.x, hv.y, hv.z, hv.w};
                if (sg.scale) {
                    v[0] = act1(v[0], ha.x, hs.x, sg.silu); v[1] = act1(v[1], ha.y, hs.y, sg.silu);
                    v[2] = act1(v[2], ha.z, hs.z, sg.silu); v[3] = act1(v[3], ha.w, hs.w, sg.silu);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int c = 0; c < CO; ++c) acc[2 - k][c] = fmaf(v[e], hw[e][k][c], acc[2 - k][c]);
            } else
            for (int c0 = lane * 4; c0 < Cin; c0 += 128) {
                const float4 hv = ld4(hb + (size_t)r * Cin + c0);
                float v[4] = {hv.x, hv.y, hv.z, hv.w};
                if (sg.scale) {
                    const float4 a4 = ld4(sg.scale + (size_t)b * Cin + c0), s4 = ld4(sg.shift + (size_t)b * Cin + c0);
                    v[0] = act1(v[0], a4.x, s4.x, sg.silu); v[1] = act1(v[1], a4.y, s4.y, sg.silu);
                    v[2] = act1(v[2], a4.z, s4.z, sg.silu); v[3] = act1(v[3], a4.w, s4.w, sg.silu);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {   // input row r is tap k of output row r + 1 - k  ->  window slot 2 - k
                        const float* wr = sg.w + (size_t)((c0 + e) * 3 + k) * p.Cout;
#pragma unroll
                        for (int c = 0; c < CO; ++c) if (c < p.Cout) acc[2 - k][c] = fmaf(v[e], __ldg(wr + c), acc[2 - k][c]);
                    }
            }
        }
        // output row r-1 is complete once input row r has been added
        const int to = r - 1;
        if (to >= t0 && to < t1) {
#pragma unroll
            for (int c = 0; c < CO; ++c) {
                float s = acc[0][c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == c && c < p.Cout) {
                    float val = s + (p.bias ? p.bias[c] : 0.f);
                    const size_t oi = ((size_t)b * T + to) * p.Cout + c;
                    if (p.ddim_x) val = p.ddim_coef[0] * p.ddim_x[oi] + p.ddim_coef[1] * val;
                    p.out[oi] = val;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CO; ++c) { acc[0][c] = acc[1][c]; acc[1][c] = acc[2][c]; acc[2][c] = 0.f; }
    }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics -> per-(sample, channel) scale/shift.  Shifted one-pass sums per thread,
// Chan's pairwise combination across threads / splits (no E[x^2]-E[x]^2 cancellation).
struct Mom { float n, mean, m2; };
__device__ __forceinline__ Mom mom_combine(Mom a, Mom b) {
    if (b.n == 0.f) return a;
    if (a.n == 0.f) return b;
    Mom r;
    r.n = a.n + b.n;
    const float d = b.mean - a.mean;
    r.mean = a.mean + d * (b.n / r.n);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
    return r;
}

// grid (nsplit, B); block (C/4, R).  Requires C0%4==0, C1%4==0, (C/G)%4==0, C/4<=256, G<=blockDim.x*blockDim.y.
__global__ void gn_partial_vec_kernel(const GnParams p) {
    __shared__ Mom sm[256];
    const int C = p.C0 + p.C1;
    const int b = blockIdx.y, split = blockIdx.x;
    const int q = threadIdx.x, c = q * 4;
    const float* src; int cc, Cs;
    if (c < p.C0) { src = p.src0; cc = c; Cs = p.C0; } else { src = p.src1; cc = c - p.C0; Cs = p.C1; }
    const int rps = (p.T + p.nsplit - 1) / p.nsplit;
    const int t_lo = split * rps, t_hi = min(p.T, t_lo + rps);
    const float* base = src + (size_t)b * p.T * Cs + cc;
    float K = 0.f, s = 0.f, ss = 0.f, n = 0.f;
    for (int t = t_lo + threadIdx.y; t < t_hi; t += blockDim.y) {
        float4 v = ld4(base + (size_t)t * Cs);
        if (n == 0.f) K = v.x;
        float d0 = v.x - K, d1 = v.y - K, d2 = v.z - K, d3 = v.w - K;
        s += (d0 + d1) + (d2 + d3);
        ss += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        n += 4.f;
    }
    Mom m;
    m.n = n;
    m.mean = n > 0.f ? K + s / n : 0.f;
    m.m2 = n > 0.f ? fmaxf(ss - s * s / n, 0.f) : 0.f;
    const int nx = blockDim.x;
    sm[threadIdx.y * nx + q] = m;
    __syncthreads();
    const int lin = threadIdx.y * nx + threadIdx.x;
    if (lin < p.G) {
        const int qpg = (C / p.G) / 4;
        Mom r; r.n = 0.f; r.mean = 0.f; r.m2 = 0.f;
        for (int y = 0; y < (int)blockDim.y; ++y)
            for (int k = 0; k < qpg; ++k) r = mom_combine(r, sm[y * nx + lin * qpg + k]);
        float* o = p.partial + (((size_t)b * p.nsplit + split) * p.G + lin) * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}

// grid (nsplit, G, B); block 256.  Any C, G.
__global__ void gn_partial_generic_kernel(const GnParams p) {
    __shared__ Mom sm[256];
    const int C = p.C0 + p.C1;
    const int cpg = C / p.G;
    const int split = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const int rps = (p.T + p.nsplit - 1) / p.nsplit;
    const int t_lo = split * rps, t_hi = min(p.T, t_lo + rps);
    const int total = max(t_hi - t_lo, 0) * cpg;
    float K = 0.f, s = 0.f, ss = 0.f, n = 0.f;
    for (int idx = threadIdx.x; idx < total; idx += 256) {
        const int t = t_lo + idx / cpg, c = g * cpg + idx % cpg;
        float v = (c < p.C0) ? p.src0[((size_t)b * p.T + t) * p.C0 + c]
                             : p.src1[((size_t)b * p.T + t) * p.C1 + (c - p.C0)];
        if (n == 0.f) K = v;
        const float d = v - K;
        s += d; ss += d * d; n += 1.f;
    }
    Mom m;
    m.n = n;
    m.mean = n > 0.f ? K + s / n : 0.f;
    m.m2 = n > 0.f ? fmaxf(ss - s * s / n, 0.f) : 0.f;
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sm[threadIdx.x] = mom_combine(sm[threadIdx.x], sm[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float* o = p.partial + (((size_t)b * p.nsplit + split) * p.G + g) * 3;
        o[0] = sm[0].n; o[1] = sm[0].mean; o[2] = sm[0].m2;
    }
}

// grid (B); block 256.  The nsplit partial records of a group are combined by 256/G' threads in parallel (G' = G rounded
// up to a power of two <= 256), then pairwise through shared memory.
__global__ void gn_finalize_kernel(const GnParams p) {
    __shared__ Mom sm[256];
    __shared__ float s_mean[256], s_rstd[256];
    const int C = p.C0 + p.C1, b = blockIdx.x, cpg = C / p.G;
    int gp = 1;
    while (gp < p.G) gp <<= 1;
    if (gp <= 256) {
        const int lanes = 256 / gp;                       // threads per group
        const int g = threadIdx.x % gp, part = threadIdx.x / gp;
        Mom r; r.n = 0.f; r.mean = 0.f; r.m2 = 0.f;
        if (g < p.G) {
            // where this group's records live: its own column of `partial`, or whole records of the sources at THEIR record
            // width -- a group of a virtual concat may take its first channels from source 0 and the rest from source 1
            // (768 = 512 + 256 channels in 32 groups of 24: group 21 is one 8-channel record of source 0 + four 4-channel
            // records of source 1)
            if (!p.rec_G0) {
                for (int sp = part; sp < p.nsplit; sp += lanes) {
                    const float* o = p.partial + (((size_t)b * p.nsplit + sp) * p.G + g) * 3;
                    Mom m; m.n = o[0]; m.mean = o[1]; m.m2 = o[2];
                    r = mom_combine(r, m);
                }
            } else {
                const int c0 = g * cpg, c1 = c0 + cpg;    // channel range of the group in the concatenated tensor
#pragma unroll
                for (int src = 0; src < 2; ++src) {
                    const int cb = src ? p.C0 : 0, cs = src ? p.C1 : p.C0, recG = src ? p.rec_G1 : p.rec_G0;
                    const int lo = max(c0, cb), hi = min(c1, cb + cs);
                    if (lo >= hi || recG <= 0) continue;
                    const float* base = src ? p.partial1 : p.partial;
                    const int w = cs / recG, first = (lo - cb) / w, nrec = (hi - lo) / w;
                    const int ns = src && p.nsplit1 ? p.nsplit1 : p.nsplit;     // records per sample of this source
                    for (int sp = part; sp < ns; sp += lanes)
                        for (int k = 0; k < nrec; ++k) {
                            const float* o = base + (((size_t)b * ns + sp) * recG + first + k) * 3;
                            Mom m; m.n = o[0]; m.mean = o[1]; m.m2 = o[2];
                            r = mom_combine(r, m);
                        }
                }
            }
        }
        sm[threadIdx.x] = r;
        __syncthreads();
        for (int off = lanes / 2; off > 0; off >>= 1) {
            if (part < off) sm[threadIdx.x] = mom_combine(sm[threadIdx.x], sm[threadIdx.x + off * gp]);
            __syncthreads();
        }
        if (part == 0 && g < p.G) {
            r = sm[g];
            s_mean[g] = r.mean;
            s_rstd[g] = 1.0f / sqrtf(r.m2 / r.n + p.eps);   // biased variance, as nn.GroupNorm
            if (p.mean_out) { p.mean_out[(size_t)b * p.G + g] = s_mean[g]; p.rstd_out[(size_t)b * p.G + g] = s_rstd[g]; }
        }
    } else {
        for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
            Mom r; r.n = 0.f; r.mean = 0.f; r.m2 = 0.f;
            for (int sp = 0; sp < p.nsplit; ++sp) {
                const float* o = p.partial + (((size_t)b * p.nsplit + sp) * p.G + g) * 3;
                Mom m; m.n = o[0]; m.mean = o[1]; m.m2 = o[2];
                r = mom_combine(r, m);
            }
            s_mean[g % 256] = r.mean;   // G > 256 is not used by any model here; scale/shift written directly below
            s_rstd[g % 256] = 1.0f / sqrtf(r.m2 / r.n + p.eps);
            if (p.mean_out) { p.mean_out[(size_t)b * p.G + g] = r.mean; p.rstd_out[(size_t)b * p.G + g] = s_rstd[g % 256]; }
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                const float a = p.gamma[c] * s_rstd[g % 256];
                p.scale[(size_t)b * C + c] = a;
                p.shift[(size_t)b * C + c] = p.beta[c] - r.mean * a;
            }
        }
        return;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float a = p.gamma[c] * s_rstd[g];
        p.scale[(size_t)b * C + c] = a;
        p.shift[(size_t)b * C + c] = p.beta[c] - s_mean[g] * a;
    }
}

// ------------------------------------------------------------------------------------------------
// Attention (legacy qkv head layout).  grid (ceil(T/32), H, B), block 256, dynamic smem.
constexpr int TQ = 32;
template <int NJ>
__global__ void __launch_bounds__(256) attention_simt_kernel(const AttnParams p) {
    extern __shared__ float smem[];
    const int T = p.T, ch = p.ch;
    const int SP = T + 1;
    float* S = smem;                 // [TQ][SP]
    float* Kc = S + TQ * SP;         // [T][33]
    float* Qc = Kc + (size_t)T * 33; // [TQ][33]
    const int b = blockIdx.z, h = blockIdx.y, tq0 = blockIdx.x * TQ;
    const int W3 = p.H * 3 * ch;
    const float* qbase = p.qkv + (size_t)b * T * W3 + (size_t)h * 3 * ch;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;

    float acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < ch; c0 += 32) {
        __syncthreads();
        for (int idx = tid; idx < TQ * 32; idx += 256) {
            const int t = idx >> 5, cc = idx & 31;
            float v = 0.f;
            if (tq0 + t < T && c0 + cc < ch) v = qbase[(size_t)(tq0 + t) * W3 + c0 + cc];
            Qc[t * 33 + cc] = v;
        }
        for (int idx = tid; idx < T * 32; idx += 256) {
            const int s = idx >> 5, cc = idx & 31;
            float v = 0.f;
            if (c0 + cc < ch) v = qbase[(size_t)s * W3 + ch + c0 + cc];
            Kc[s * 33 + cc] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int cc = 0; cc < 32; ++cc) {
            float qv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = Qc[(ty * 4 + i) * 33 + cc];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int s = tx + 32 * j;
                const float kv = (s < T) ? Kc[s * 33 + cc] : 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(qv[i], kv, acc[i][j]);
            }
        }
    }
    // scale = ch^-1/4 applied to both q and k (unet.py:119-121)  ->  scores * ch^-1/2
    const float sc = 1.0f / sqrtf((float)ch);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int s = tx + 32 * j;
            if (s < T) S[(ty * 4 + i) * SP + s] = acc[i][j] * sc;
        }
    __syncthreads();
    // softmax over s (fp32, unet.py:123)
    for (int r = ty; r < TQ; r += 8) {
        float* row = S + r * SP;
        float m = -INFINITY;
        for (int s = tx; s < T; s += 32) m = fmaxf(m, row[s]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
        for (int s = tx; s < T; s += 32) { const float e = expf(row[s] - m); row[s] = e; sum += e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        for (int s = tx; s < T; s += 32) row[s] *= inv;
    }
    __syncthreads();
    // a[t][c] = sum_s P[t][s] * v[s][c]
    const int cl = tid & 63, tg = tid >> 6;  // 4 groups of 8 query rows
    const float* vbase = qbase + 2 * ch;
    float* obase = p.out + (size_t)b * T * (p.H * ch) + (size_t)h * ch;
    for (int cb = 0; cb < ch; cb += 64) {
        const int c = cb + cl;
        const bool cv = c < ch;
        float o8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o8[i] = 0.f;
#pragma unroll 4
        for (int s = 0; s < T; ++s) {
            const float v = cv ? vbase[(size_t)s * W3 + c] : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) o8[i] = fmaf(S[(tg * 8 + i) * SP + s], v, o8[i]);
        }
        if (cv) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = tq0 + tg * 8 + i;
                if (t < T) obase[(size_t)t * (p.H * ch) + c] = o8[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tiny-channel convolution (the 2-2-4 autoencoder of config_aekl_eeg_2_2_4_spec.yaml: 1, 2 or 4 channels everywhere).
// One thread computes P = 4 consecutive output positions x all CO channels of one sample from (P-1)*S + TAPS input rows
// held in registers (vector loads of CI floats per row); weights, bias and the optional 1x1 skip weights sit in shared
// memory (broadcast reads).  Same fused contract as conv_simt_kernel: GroupNorm apply + SiLU prologue, nearest x2 upsample
// folded into the row index, stride 1 | 2 with explicit left pad, + bias + identity residual | + 1x1 skip conv of the raw
// block input (nin_shortcut).  HBM-bound: every input row is read once per 4 outputs, every output written once.
template <int C> struct RowVec;
template <> struct RowVec<1> { using T = float; };
template <> struct RowVec<2> { using T = float2; };
template <> struct RowVec<4> { using T = float4; };
template <int C>
__device__ __forceinline__ void load_row(const float* p, float (&v)[C]) {
    const typename RowVec<C>::T r = *reinterpret_cast<const typename RowVec<C>::T*>(p);
    const float* f = reinterpret_cast<const float*>(&r);
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = f[c];
}
template <int C>
__device__ __forceinline__ void store_row(float* p, const float (&v)[C]) {
    typename RowVec<C>::T r;
    float* f = reinterpret_cast<float*>(&r);
#pragma unroll
    for (int c = 0; c < C; ++c) f[c] = v[c];
    *reinterpret_cast<typename RowVec<C>::T*>(p) = r;
}

constexpr int TINY_P = 4;
template <int CI, int CO, int TAPS, int S>
__global__ void __launch_bounds__(256) conv_tiny_kernel(const ConvParams p) {
    constexpr int NR = (TINY_P - 1) * S + TAPS;
    __shared__ float ws[CI * TAPS * CO], bs[CO], w2[4 * CO];
    const ConvSeg& s0 = p.seg[0];
    for (int i = threadIdx.x; i < CI * TAPS * CO; i += blockDim.x) ws[i] = s0.w[i];
    for (int i = threadIdx.x; i < CO; i += blockDim.x) bs[i] = p.bias ? p.bias[i] : 0.f;
    const int C2 = p.nseg > 1 ? p.seg[1].C0 : 0;
    for (int i = threadIdx.x; i < C2 * CO; i += blockDim.x) w2[i] = p.seg[1].w[i];
    __syncthreads();
    const int groups = (p.Tout + TINY_P - 1) / TINY_P;
    const size_t total = (size_t)p.B * groups;
    const bool ups = s0.resample == RS_NEAREST2;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / groups), t0 = (int)(idx % groups) * TINY_P;
        float a[CI], sh[CI];
#pragma unroll
        for (int c = 0; c < CI; ++c) {
            a[c] = s0.scale ? s0.scale[(size_t)b * CI + c] : 1.f;
            sh[c] = s0.scale ? s0.shift[(size_t)b * CI + c] : 0.f;
        }
        float x[NR][CI];
        const int u0 = t0 * S - p.pad_left;
        const float* src = s0.src0 + (size_t)b * s0.Tin * CI;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int u = u0 + r;
            if (u >= 0 && u < p.Tc) {
                load_row<CI>(src + (size_t)(ups ? (u >> 1) : u) * CI, x[r]);
#pragma unroll
                for (int c = 0; c < CI; ++c) x[r][c] = act1(x[r][c], a[c], sh[c], s0.silu);
            } else {
#pragma unroll
                for (int c = 0; c < CI; ++c) x[r][c] = 0.f;   // the conv's zero padding (applied after the activation)
            }
        }
#pragma unroll
        for (int j = 0; j < TINY_P; ++j) {
            const int t = t0 + j;
            if (t >= p.Tout) break;
            float acc[CO];
#pragma unroll
            for (int co = 0; co < CO; ++co) acc[co] = bs[co];
#pragma unroll
            for (int c = 0; c < CI; ++c)
#pragma unroll
                for (int k = 0; k < TAPS; ++k)
#pragma unroll
                    for (int co = 0; co < CO; ++co) acc[co] = fmaf(x[j * S + k][c], ws[(c * TAPS + k) * CO + co], acc[co]);
            if (p.res) {
                float r[CO];
                load_row<CO>(p.res + ((size_t)b * p.res_Tin + t) * CO, r);
#pragma unroll
                for (int co = 0; co < CO; ++co) acc[co] += r[co];
            }
            if (C2) {
                const float* x2 = p.seg[1].src0 + ((size_t)b * p.seg[1].Tin + t) * C2;
                for (int c = 0; c < C2; ++c) {
                    const float v = x2[c];
#pragma unroll
                    for (int co = 0; co < CO; ++co) acc[co] = fmaf(v, w2[c * CO + co], acc[co]);
                }
            }
            store_row<CO>(p.out + ((size_t)b * p.Tout + t) * CO, acc);
        }
    }
}

template <int CI, int CO>
cudaError_t launch_conv_tiny_t(const ConvParams& p, cudaStream_t st) {
    const size_t total = (size_t)p.B * ((p.Tout + TINY_P - 1) / TINY_P);
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
    const int taps = p.seg[0].taps;
    if (taps == 3 && p.stride == 1) conv_tiny_kernel<CI, CO, 3, 1><<<blocks, 256, 0, st>>>(p);
    else if (taps == 3 && p.stride == 2) conv_tiny_kernel<CI, CO, 3, 2><<<blocks, 256, 0, st>>>(p);
    else if (taps == 1 && p.stride == 1) conv_tiny_kernel<CI, CO, 1, 1><<<blocks, 256, 0, st>>>(p);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
bool conv_tiny_ok(const ConvParams& p) {
    const ConvSeg& a = p.seg[0];
    auto c124 = [](int c) { return c == 1 || c == 2 || c == 4; };
    if (!c124(a.C0) || a.C1 || !c124(p.Cout) || p.temb || p.ddim_x) return false;
    if (!(a.taps == 3 || (a.taps == 1 && p.stride == 1)) || (p.stride != 1 && p.stride != 2)) return false;
    if (a.resample != RS_NONE && a.resample != RS_NEAREST2) return false;
    if (p.res && (p.res_mode != RS_NONE || p.nseg > 1)) return false;
    if (p.nseg > 1) {
        const ConvSeg& q = p.seg[1];
        if (q.taps != 1 || q.C1 || q.C0 > 4 || q.scale || q.silu || q.resample != RS_NONE || q.Tin != p.Tout) return false;
    }
    return true;
}
cudaError_t launch_conv_tiny(const ConvParams& p, cudaStream_t st) {
    const int ci = p.seg[0].C0, co = p.Cout;
#define EEGLDM_TINY(CI, CO) if (ci == CI && co == CO) return launch_conv_tiny_t<CI, CO>(p, st);
    EEGLDM_TINY(1, 1) EEGLDM_TINY(1, 2) EEGLDM_TINY(1, 4) EEGLDM_TINY(2, 1) EEGLDM_TINY(2, 2) EEGLDM_TINY(2, 4)
    EEGLDM_TINY(4, 1) EEGLDM_TINY(4, 2) EEGLDM_TINY(4, 4)
#undef EEGLDM_TINY
    return cudaErrorInvalidValue;
}

// GroupNorm statistics for tensors of at most 8 channels (single source): grid (nsplit, B), block 256.  Per-channel shifted
// sums per thread, Chan combination through warp shuffles and shared memory, channels folded into groups by thread 0.
__global__ void __launch_bounds__(256) gn_partial_tiny_kernel(const GnParams p) {
    __shared__ Mom sm[8][8];
    const int C = p.C0, b = blockIdx.y, split = blockIdx.x;
    const int rps = (p.T + p.nsplit - 1) / p.nsplit;
    const int t_lo = split * rps, t_hi = min(p.T, t_lo + rps);
    const float* base = p.src0 + (size_t)b * p.T * C;
    float K[8], s[8], ss[8];
    float n = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) { K[c] = 0.f; s[c] = 0.f; ss[c] = 0.f; }
    for (int t = t_lo + threadIdx.x; t < t_hi; t += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c < C) {
                const float v = base[(size_t)t * C + c];
                if (n == 0.f) K[c] = v;
                const float d = v - K[c];
                s[c] += d; ss[c] = fmaf(d, d, ss[c]);
            }
        }
        n += 1.f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (c >= C) break;
        Mom m;
        m.n = n; m.mean = n > 0.f ? K[c] + s[c] / n : 0.f; m.m2 = n > 0.f ? fmaxf(ss[c] - s[c] * s[c] / n, 0.f) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Mom q;
            q.n = __shfl_xor_sync(0xffffffffu, m.n, o); q.mean = __shfl_xor_sync(0xffffffffu, m.mean, o);
            q.m2 = __shfl_xor_sync(0xffffffffu, m.m2, o);
            m = mom_combine(m, q);
        }
        if (lane == 0) sm[c][warp] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int cpg = C / p.G, nw = blockDim.x >> 5;
        for (int g = 0; g < p.G; ++g) {
            Mom r; r.n = 0.f; r.mean = 0.f; r.m2 = 0.f;
            for (int c = g * cpg; c < (g + 1) * cpg; ++c)
                for (int w = 0; w < nw; ++w) r = mom_combine(r, sm[c][w]);
            float* o = p.partial + (((size_t)b * p.nsplit + split) * p.G + g) * 3;
            o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// y[r][o] = bias[o] + sum_i act(x[r][i]) W[o][i]   -- one warp per output
__global__ void linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                              const float* __restrict__ bias, float* __restrict__ y, int R, int I, int O,
                              int silu_in) {
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)R * O) return;
    const int r = (int)(wid / O), o = (int)(wid % O);
    const float* xr = x + (size_t)r * I;
    const float* wr = W + (size_t)o * I;
    float s = 0.f;
    for (int i = lane; i < I; i += 32) {
        float v = xr[i];
        if (silu_in) v = silu_f(v);
        s = fmaf(v, wr[i], s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[(size_t)r * O + o] = s + (bias ? bias[o] : 0.f);
}

__global__ void transpose_ncl_to_nlc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T,
                                            size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const size_t bt = i / C;
    const int t = (int)(bt % T);
    const size_t b = bt / T;
    out[i] = in[(b * C + c) * T + t];
}
__global__ void transpose_nlc_to_ncl_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T,
                                            size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int t = (int)(i % T);
    const size_t bc = i / T;
    const int c = (int)(bc % C);
    const size_t b = bc / C;
    out[i] = in[(b * T + t) * C + c];
}
__global__ void resample_kernel(const float* __restrict__ in, float* __restrict__ out, int Tin, int Tout, int C, int mode,
                                size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const size_t bt = i / C;
    const int t = (int)(bt % Tout);
    const size_t b = bt / Tout;
    const float* base = in + (b * Tin) * C + c;
    out[i] = mode == RS_AVGPOOL2 ? 0.5f * (base[(size_t)(2 * t) * C] + base[(size_t)(2 * t + 1) * C]) : base[(size_t)(t >> 1) * C];
}
__global__ void scale_kernel(const float* __restrict__ src, float* __restrict__ dst, float alpha, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] * alpha;
}
// timestep_embedding (unet.py:12-36) for device-resident timesteps; double precision with fp32 rounding at the points the
// reference rounds (freqs, t*freq, cos / sin), identical to the host version in engine.cu
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int nt, int dim, float* __restrict__ out) {
    const int half = dim / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nt * half) return;
    const int i = idx / half, j = idx % half;
    const float fr = (float)exp((double)(-(float)log(10000.0) * (float)j / (float)half));
    const float arg = t[i] * fr;
    out[(size_t)i * dim + j] = (float)cos((double)arg);
    out[(size_t)i * dim + half + j] = (float)sin((double)arg);
    if ((dim & 1) && j == 0) out[(size_t)i * dim + dim - 1] = 0.f;
}

__global__ void step_advance_kernel(const float* __restrict__ temb_table, int temb_row, float* __restrict__ temb_cur,
                                    const float* __restrict__ coef_table, float* __restrict__ coef_cur, int* step) {
    const int s = *step;
    for (int i = threadIdx.x; i < temb_row; i += blockDim.x) temb_cur[i] = temb_table[(size_t)s * temb_row + i];
    if (threadIdx.x < 2) coef_cur[threadIdx.x] = coef_table[2 * s + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) *step = s + 1;
}
__global__ void kl_sigma_kernel(const float* __restrict__ logvar, float* __restrict__ sigma, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sigma[i] = expf(fminf(fmaxf(logvar[i], -30.f), 20.f) / 2.f);
}
__global__ void axpy_sampling_kernel(const float* __restrict__ mu, const float* __restrict__ sigma,
                                     const float* __restrict__ eps, float* __restrict__ z, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = mu[i] + eps[i] * sigma[i];
}

}  // namespace

// ================================================================================================ launchers
// can launch_conv_simt emit GroupNorm records for this conv (ConvParams.gn_rec)?  Mirrors the narrow-input branch below.
bool conv_narrow_in_gn_ok(const ConvParams& p) {
    const ConvSeg& s0 = p.seg[0];
    return p.nseg == 1 && s0.taps == 3 && p.stride == 1 && p.pad_left == 1 && !s0.src1 && s0.resample == RS_NONE && !s0.silu && !p.temb &&
           !p.res && p.Tc == p.Tout && s0.C0 <= 4 && !s0.scale && !p.ddim_x && p.Cout == 128 && p.Tout % 16 == 0 &&
           (size_t)p.B * p.Tout < (1u << 31);
}

cudaError_t launch_conv_simt(const ConvParams& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    const ConvSeg& s0 = p.seg[0];
    const bool plain = p.nseg == 1 && s0.taps == 3 && p.stride == 1 && p.pad_left == 1 && !s0.src1 && s0.resample == RS_NONE &&
                       p.Tc == p.Tout && !p.temb && !p.res;
    if (plain && s0.C0 <= 4 && !s0.scale && !p.ddim_x && (p.Cout & 3) == 0 && p.Cout >= 32 && 256 % (p.Cout / 4) == 0 &&
        (size_t)p.B * p.Tout < (1u << 31) && (size_t)(s0.C0 * 3 + 1) * p.Cout * sizeof(float) <= 40 * 1024) {
        const size_t total4 = (size_t)p.B * p.Tout * (p.Cout / 4);
        const size_t smem = ((size_t)s0.C0 * 3 * p.Cout + p.Cout) * sizeof(float);
        if (p.gn_rec) {
            if (p.Cout != 128 || p.Tout % 16) return cudaErrorInvalidValue;
            const int per = p.gn_rec_tile ? 128 : 16;
            if (p.Tout % per) return cudaErrorInvalidValue;
            const int nsegs = (int)((size_t)p.B * p.Tout / per);
            const unsigned nb = (unsigned)std::min<size_t>(((size_t)nsegs + 7) / 8, 148 * 16);
            conv_narrow_in_gn_kernel<<<nb, 256, smem, st>>>(s0.src0, s0.w, p.bias, p.out, p.gn_rec, s0.C0, p.Tout, nsegs, per);
            g_launch_count += 1;
            return cudaGetLastError();
        }
        const unsigned blocks = (unsigned)std::min<size_t>((total4 + 255) / 256, 148 * 16);
        conv_narrow_in_kernel<<<blocks, 256, smem, st>>>(s0.src0, s0.w, p.bias, p.out, s0.C0, p.Cout, p.Tout, total4);
        g_launch_count += 1;
        return cudaGetLastError();
    }
    if (plain && p.Cout <= 4 && s0.C0 % 128 == 0) {
        const long long warps = (long long)p.B * ((p.Tout + NO_STRIP - 1) / NO_STRIP);
        const unsigned nb = (unsigned)((warps * 32 + 255) / 256);
        if (p.Cout == 1) conv_narrow_out_kernel<1><<<nb, 256, 0, st>>>(p);
        else if (p.Cout == 2) conv_narrow_out_kernel<2><<<nb, 256, 0, st>>>(p);
        else conv_narrow_out_kernel<4><<<nb, 256, 0, st>>>(p);
        g_launch_count += 1;
        return cudaGetLastError();
    }
    if (conv_tiny_ok(p)) {
        cudaError_t e = launch_conv_tiny(p, st);
        g_launch_count += 1;
        return e;
    }
    const int gx = (p.Tout + BN - 1) / BN;
    if (p.Cout >= 96) {
        dim3 grid(gx, (p.Cout + 127) / 128, p.B);
        conv_simt_kernel<8><<<grid, 256, 0, st>>>(p);
    } else if (p.Cout >= 24) {
        dim3 grid(gx, (p.Cout + 63) / 64, p.B);
        conv_simt_kernel<4><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid(gx, (p.Cout + 15) / 16, p.B);
        conv_simt_kernel<1><<<grid, 256, 0, st>>>(p);
    }
    g_launch_count += 1;
    return cudaGetLastError();
}

int groupnorm_nsplit(int C, int T, int G) {
    (void)G;
    if (C <= 8) { int n = T / 768; return n < 1 ? 1 : n; }   // gn_partial_tiny_kernel: >= 3 rows per thread
    int n = T / 96;
    if (n < 1) n = 1;
    if (n > 32) n = 32;
    return n;
}

cudaError_t launch_groupnorm(const GnParams& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    const int C = p.C0 + p.C1;
    const int cpg = C / p.G;
    const bool vec = (p.C0 % 4 == 0) && (p.C1 % 4 == 0) && (cpg % 4 == 0) && (C / 4 <= 256) && (p.G <= 32);
    if (C <= 8 && !p.src1) {
        dim3 grid(p.nsplit, p.B);
        gn_partial_tiny_kernel<<<grid, 256, 0, st>>>(p);
    } else if (vec) {
        const int nx = C / 4;
        int ny = 256 / nx;
        if (ny < 1) ny = 1;
        while (nx * ny < p.G) ++ny;
        dim3 grid(p.nsplit, p.B), block(nx, ny);
        gn_partial_vec_kernel<<<grid, block, 0, st>>>(p);
    } else {
        dim3 grid(p.nsplit, p.G, p.B);
        gn_partial_generic_kernel<<<grid, 256, 0, st>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    gn_finalize_kernel<<<p.B, 256, 0, st>>>(p);
    g_launch_count += 2;
    return cudaGetLastError();
}

cudaError_t launch_groupnorm_finalize(const GnParams& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    gn_finalize_kernel<<<p.B, 256, 0, st>>>(p);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_attention_simt(const AttnParams& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    const size_t smem = ((size_t)TQ * (p.T + 1) + (size_t)p.T * 33 + TQ * 33) * sizeof(float);
    dim3 grid((p.T + TQ - 1) / TQ, p.H, p.B);
    const int nj = (p.T + 31) / 32;
    cudaError_t e;
#define EEGLDM_ATTN(NJ)                                                                                        \
    do {                                                                                                       \
        static bool attr_set = false;                                                                          \
        if (!attr_set) {                                                                                       \
            e = cudaFuncSetAttribute(attention_simt_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                     227 * 1024);                                                              \
            if (e != cudaSuccess) return e;                                                                    \
            attr_set = true;                                                                                   \
        }                                                                                                      \
        attention_simt_kernel<NJ><<<grid, 256, smem, st>>>(p);                                                 \
    } while (0)
    if (nj <= 3) EEGLDM_ATTN(3);
    else if (nj <= 6) EEGLDM_ATTN(6);
    else if (nj <= 12) EEGLDM_ATTN(12);
    else if (nj <= 24) EEGLDM_ATTN(24);
    else return cudaErrorInvalidValue;
#undef EEGLDM_ATTN
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_linear(const float* x, const float* W, const float* bias, float* y, int R, int I, int O,
                          int silu_in, cudaStream_t st) {
    const long long warps = (long long)R * O;
    if (warps <= 0) return cudaSuccess;
    const long long blocks = (warps * 32 + 255) / 256;
    linear_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, W, bias, y, R, I, O, silu_in);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_transpose_ncl_to_nlc(const float* in, float* out, int B, int C, int T, cudaStream_t st) {
    const size_t total = (size_t)B * C * T;
    if (!total) return cudaSuccess;
    transpose_ncl_to_nlc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, C, T, total);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_transpose_nlc_to_ncl(const float* in, float* out, int B, int C, int T, cudaStream_t st) {
    const size_t total = (size_t)B * C * T;
    if (!total) return cudaSuccess;
    transpose_nlc_to_ncl_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, C, T, total);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_resample(const float* in, float* out, int B, int Tin, int C, int mode, cudaStream_t st) {
    const int Tout = mode == RS_AVGPOOL2 ? Tin / 2 : Tin * 2;
    const size_t total = (size_t)B * Tout * C;
    if (!total) return cudaSuccess;
    resample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, Tin, Tout, C, mode, total);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_scale(const float* src, float* dst, float alpha, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, alpha, n);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_timestep_embedding(const float* t_dev, int nt, int dim, float* out, cudaStream_t st) {
    const int n = nt * (dim / 2);
    if (n <= 0) return cudaSuccess;
    timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, st>>>(t_dev, nt, dim, out);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_step_advance(const float* temb_table, int temb_row, float* temb_cur, const float* coef_table,
                                float* coef_cur, int* step, cudaStream_t st) {
    step_advance_kernel<<<1, 256, 0, st>>>(temb_table, temb_row, temb_cur, coef_table, coef_cur, step);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_kl_sigma(const float* logvar, float* sigma, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    kl_sigma_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logvar, sigma, n);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_axpy_sampling(const float* mu, const float* sigma, const float* eps, float* z, size_t n,
                                 cudaStream_t st) {
    if (!n) return cudaSuccess;
    axpy_sampling_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mu, sigma, eps, z, n);
    g_launch_count += 1;
    return cudaGetLastError();
}

}  // namespace eegldm
