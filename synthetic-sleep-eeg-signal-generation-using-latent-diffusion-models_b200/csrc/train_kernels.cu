// Backward / loss / optimiser kernels of the AutoencoderKL training step (src/train_autoencoderkl.py:204-220):
//   recon, mu, sigma = model(x);  loss = L1(recon, x) + kl_w * KL(mu, sigma) + spec_w * Jukebox(recon, x);  Adam step.
// The autoencoder's channels are narrow (2..64), so everything here is fp32 SIMT and HBM / latency bound; tensors are
// channels-last [B][T][C] like the rest of the engine, conv weights use the SIMT image [(ci*taps + k)][Cout].
// (The cuFFT side of the spectral loss is in spectral.cu.)
#include "kernels.cuh"

#include <algorithm>

namespace eegldm {
namespace {

__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }

template <int N>
__device__ __forceinline__ void block_reduce_sum(float (&v)[N], float* sm /* [N][32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
        if (lane == 0) sm[i * 32 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float s = (threadIdx.x < nw) ? sm[i * 32 + threadIdx.x] : 0.f;
        if (warp == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        }
        v[i] = s;   // valid in thread 0
    }
    __syncthreads();
}

// a = silu?(scale[b][c] * x + shift[b][c])
__global__ void norm_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                    float* __restrict__ a, int C, int T, int silu, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const size_t b = i / ((size_t)C * T);
    float v = fmaf(scale[b * C + c], x[i], shift[b * C + c]);
    if (silu) v = v * sigmoid_f(v);
    a[i] = v;
}

// the same, four channels per thread (C % 4 == 0: the UNet's tensors) -- identical per-element arithmetic, a quarter of the index
// divisions and 128-bit accesses
__global__ void norm_act_fwd4_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                     float* __restrict__ a, int C, int T, int silu, size_t total4) {
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const size_t i = i4 * 4;
    const int c = (int)(i % C);
    const size_t b = i / ((size_t)C * T);
    const float4 xv = *reinterpret_cast<const float4*>(x + i);
    const float4 sc = *reinterpret_cast<const float4*>(scale + b * C + c), sh = *reinterpret_cast<const float4*>(shift + b * C + c);
    float v[4] = {fmaf(sc.x, xv.x, sh.x), fmaf(sc.y, xv.y, sh.y), fmaf(sc.z, xv.z, sh.z), fmaf(sc.w, xv.w, sh.w)};
    if (silu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = v[e] * sigmoid_f(v[e]);
    }
    *reinterpret_cast<float4*>(a + i) = make_float4(v[0], v[1], v[2], v[3]);
}

// Conv1d backward w.r.t. its input.  The conv consumed u = upsample?(a) with length Tc, stride s, left pad p:
//   y[t][co] = sum_{k,ci} W[ci][k][co] * u[t*s + k - p][ci]        da[b][i][ci] (+)= sum over the conv-input rows that read a[i]
__global__ void conv_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ da, int Cin,
                                     int Cout, int taps, int stride, int pad, int ups, int Tin, int Tc, int Tout, int accumulate,
                                     size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ci = (int)(idx % Cin);
    const size_t bt = idx / Cin;
    const int i = (int)(bt % Tin);
    const size_t b = bt / Tin;
    float acc = 0.f;
    const int n_uc = ups ? 2 : 1;
    for (int j = 0; j < n_uc; ++j) {
        const int uc = ups ? 2 * i + j : i;
        if (uc >= Tc) continue;
        for (int k = 0; k < taps; ++k) {
            const int num = uc + pad - k;
            if (num < 0 || num % stride) continue;
            const int t = num / stride;
            if (t >= Tout) continue;
            const float* dyr = dy + (b * Tout + t) * Cout;
            const float* wr = w + (size_t)(ci * taps + k) * Cout;
            for (int co = 0; co < Cout; ++co) acc = fmaf(dyr[co], __ldg(wr + co), acc);
        }
    }
    da[idx] = accumulate ? da[idx] + acc : acc;
}

// Conv1d backward w.r.t. weight and bias; grid (ceil(Tout/P), B), dynamic smem: dy tile [P][Cout] + input tile [P*stride+taps-1][Cin]
// (P = WG_P positions per block: 64, halved by the launcher until the two tiles fit shared memory)
__global__ void __launch_bounds__(256) conv_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ a,
                                                               float* __restrict__ dw, float* __restrict__ db, int Cin, int Cout,
                                                               int taps, int stride, int pad, int ups, int Tin, int Tc, int Tout, int WG_P,
                                                               int ntiles_t, int ntiles) {
    // Persistent over (sample, position tile): narrow convs (Cin * taps * Cout <= 2048: the 1-channel in / out convs of the UNet and
    // of the discriminator) keep their weight-gradient partial sums in registers across all the block's tiles and issue ONE
    // atomicAdd per weight per block -- with one atomic per weight per tile the 10^4 tiles of such a layer serialise on a few
    // hundred addresses.  Wider convs (fp32 fallback only; the tensor-pipe path has wgrad_tc_kernel) add per tile.
    extern __shared__ float sm[];
    float* dys = sm;                       // [P][Cout]
    float* as = sm + WG_P * Cout;          // [rows][Cin]
    const int nw = Cin * taps * Cout;
    const bool in_regs = nw <= 256 * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accb = 0.f;
    const int rows = (WG_P - 1) * stride + taps;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / ntiles_t, t0 = (tile - b * ntiles_t) * WG_P;
        const int np = min(WG_P, Tout - t0);
        const int u0 = t0 * stride - pad;
        __syncthreads();   // the previous tile's readers are done
        for (int i = threadIdx.x; i < WG_P * Cout; i += blockDim.x) {
            const int p = i / Cout, co = i % Cout;
            dys[i] = p < np ? dy[((size_t)b * Tout + t0 + p) * Cout + co] : 0.f;
        }
        for (int i = threadIdx.x; i < rows * Cin; i += blockDim.x) {
            const int r = i / Cin, ci = i % Cin;
            const int uc = u0 + r;
            float v = 0.f;
            if (uc >= 0 && uc < Tc) v = a[((size_t)b * Tin + (ups ? (uc >> 1) : uc)) * Cin + ci];
            as[i] = v;
        }
        __syncthreads();
        if (in_regs) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = threadIdx.x + 256 * j;
                if (i < nw) {
                    const int co = i % Cout, ck = i / Cout, k = ck % taps, ci = ck / taps;   // i == packed weight index
                    float s = 0.f;
                    for (int p = 0; p < np; ++p) s = fmaf(dys[p * Cout + co], as[(p * stride + k) * Cin + ci], s);
                    acc[j] += s;
                }
            }
        } else {
            for (int i = threadIdx.x; i < nw; i += blockDim.x) {
                const int co = i % Cout, ck = i / Cout, k = ck % taps, ci = ck / taps;
                float s = 0.f;
                for (int p = 0; p < np; ++p) s = fmaf(dys[p * Cout + co], as[(p * stride + k) * Cin + ci], s);
                atomicAdd(dw + i, s);
            }
        }
        if (db)
            for (int co = threadIdx.x; co < Cout; co += blockDim.x) {   // Cout <= 256 in the register form: thread co owns db[co]
                float s = 0.f;
                for (int p = 0; p < np; ++p) s += dys[p * Cout + co];
                if (in_regs && Cout <= 256) accb += s; else atomicAdd(db + co, s);
            }
    }
    if (in_regs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = threadIdx.x + 256 * j;
            if (i < nw) atomicAdd(dw + i, acc[j]);
        }
        if (db && Cout <= 256 && (int)threadIdx.x < Cout) atomicAdd(db + threadIdx.x, accb);
    }
}

// ---- tiny-channel (1 / 2 / 4) variants for the 2-2-4 autoencoder: rows are vector loads, weights live in shared memory,
// ---- the weight gradient is reduced in registers -> warp shuffles -> shared memory -> ONE atomicAdd per weight per block.
template <int C> struct RowVecT;
template <> struct RowVecT<1> { using T = float; };
template <> struct RowVecT<2> { using T = float2; };
template <> struct RowVecT<4> { using T = float4; };
template <int C>
__device__ __forceinline__ void ld_row(const float* p, float (&v)[C]) {
    const typename RowVecT<C>::T r = *reinterpret_cast<const typename RowVecT<C>::T*>(p);
    const float* f = reinterpret_cast<const float*>(&r);
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = f[c];
}
template <int C>
__device__ __forceinline__ void st_row(float* p, const float (&v)[C]) {
    typename RowVecT<C>::T r;
    float* f = reinterpret_cast<float*>(&r);
#pragma unroll
    for (int c = 0; c < C; ++c) f[c] = v[c];
    *reinterpret_cast<typename RowVecT<C>::T*>(p) = r;
}

// one thread per (sample, input position): da[b][i][:] over all CI channels
template <int CI, int CO, int TAPS>
__global__ void __launch_bounds__(256) conv_bwd_data_tiny_kernel(const ConvGradParams p) {
    __shared__ float ws[CI * TAPS * CO];
    for (int i = threadIdx.x; i < CI * TAPS * CO; i += blockDim.x) ws[i] = p.w[i];
    __syncthreads();
    const size_t total = (size_t)p.B * p.Tin;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % p.Tin);
        const size_t b = idx / p.Tin;
        float acc[CI];
#pragma unroll
        for (int c = 0; c < CI; ++c) acc[c] = 0.f;
        const int n_uc = p.ups ? 2 : 1;
        for (int j = 0; j < n_uc; ++j) {
            const int uc = p.ups ? 2 * i + j : i;
            if (uc >= p.Tc) continue;
#pragma unroll
            for (int k = 0; k < TAPS; ++k) {
                const int num = uc + p.pad - k;
                if (num < 0 || num % p.stride) continue;
                const int t = num / p.stride;
                if (t >= p.Tout) continue;
                float d[CO];
                ld_row<CO>(p.dy + (b * p.Tout + t) * CO, d);
#pragma unroll
                for (int c = 0; c < CI; ++c)
#pragma unroll
                    for (int co = 0; co < CO; ++co) acc[c] = fmaf(d[co], ws[(c * TAPS + k) * CO + co], acc[c]);
            }
        }
        float* o = p.da + idx * CI;
        if (p.accumulate) {
            float old[CI];
            ld_row<CI>(o, old);
#pragma unroll
            for (int c = 0; c < CI; ++c) acc[c] += old[c];
        }
        st_row<CI>(o, acc);
    }
}

// grid-stride over (sample, output position); dw [(ci*TAPS + k)][CO] and db [CO] accumulated with atomics, one per block
template <int CI, int CO, int TAPS>
__global__ void __launch_bounds__(256) conv_bwd_weight_tiny_kernel(const ConvGradParams p) {
    constexpr int NW = CI * TAPS * CO, NV = NW + CO;
    __shared__ float red[8][NV];
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    const size_t total = (size_t)p.B * p.Tout;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(idx % p.Tout);
        const size_t b = idx / p.Tout;
        float d[CO];
        ld_row<CO>(p.dy + idx * CO, d);
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[NW + co] += d[co];
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
            const int uc = t * p.stride + k - p.pad;
            if (uc < 0 || uc >= p.Tc) continue;
            float a[CI];
            ld_row<CI>(p.a + (b * p.Tin + (p.ups ? (uc >> 1) : uc)) * CI, a);
#pragma unroll
            for (int c = 0; c < CI; ++c)
#pragma unroll
                for (int co = 0; co < CO; ++co) acc[(c * TAPS + k) * CO + co] = fmaf(d[co], a[c], acc[(c * TAPS + k) * CO + co]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NV; i += blockDim.x) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][i];
        if (i < NW) atomicAdd(p.dw + i, v);
        else if (p.db) atomicAdd(p.db + (i - NW), v);
    }
}

// stride-1, same-length convs (30 of the 37 in the 2-2-4 autoencoder): weight gradient AND input gradient from one pass.
// The thread of output position t also owns input position t: da[t] = sum_k dy[t + pad - k] . W[:, k, :], and its weight
// contribution is dy[t] (x) a[t + k - pad]; the rows t-1 .. t+1 of dy and a are each read once per thread (L1 serves the
// overlap between neighbours).
template <int CI, int CO, int TAPS>
__global__ void __launch_bounds__(256) conv_bwd_fused_tiny_kernel(const ConvGradParams p) {
    constexpr int NW = CI * TAPS * CO, NV = NW + CO;
    __shared__ float red[8][NV];
    __shared__ float ws[NW];
    for (int i = threadIdx.x; i < NW; i += blockDim.x) ws[i] = p.w[i];
    __syncthreads();
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    const int T = p.Tout;   // == Tin == Tc
    const size_t total = (size_t)p.B * T;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(idx % T);
        float d[TAPS][CO], a[TAPS][CI];   // rows t + k - pad of dy and a (zero outside the sample)
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
            const int u = t + k - p.pad;
            if (u >= 0 && u < T) {
                ld_row<CO>(p.dy + (idx + k - p.pad) * CO, d[k]);
                ld_row<CI>(p.a + (idx + k - p.pad) * CI, a[k]);
            } else {
#pragma unroll
                for (int c = 0; c < CO; ++c) d[k][c] = 0.f;
#pragma unroll
                for (int c = 0; c < CI; ++c) a[k][c] = 0.f;
            }
        }
        constexpr int MID = TAPS / 2;      // row t itself (pad == TAPS / 2 for these convs)
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[NW + co] += d[MID][co];
        if (p.dres) {                      // y = conv(a) + res  ->  dres (+)= dy
            float g[CO];
            if (p.dres_accumulate) {
                ld_row<CO>(p.dres + idx * CO, g);
#pragma unroll
                for (int co = 0; co < CO; ++co) g[co] += d[MID][co];
            } else {
#pragma unroll
                for (int co = 0; co < CO; ++co) g[co] = d[MID][co];
            }
            st_row<CO>(p.dres + idx * CO, g);
        }
#pragma unroll
        for (int k = 0; k < TAPS; ++k)
#pragma unroll
            for (int c = 0; c < CI; ++c)
#pragma unroll
                for (int co = 0; co < CO; ++co) acc[(c * TAPS + k) * CO + co] = fmaf(d[MID][co], a[k][c], acc[(c * TAPS + k) * CO + co]);
        if (p.da) {
            float g[CI];
#pragma unroll
            for (int c = 0; c < CI; ++c) g[c] = 0.f;
            // da[t] = sum_k dy[t + pad - k] W[ci][k][:]: output row t + pad - k is window row (TAPS - 1 - k)
#pragma unroll
            for (int k = 0; k < TAPS; ++k)
#pragma unroll
                for (int c = 0; c < CI; ++c)
#pragma unroll
                    for (int co = 0; co < CO; ++co) g[c] = fmaf(d[TAPS - 1 - k][co], ws[(c * TAPS + k) * CO + co], g[c]);
            float* o = p.da + idx * CI;
            if (p.accumulate) {
                float old[CI];
                ld_row<CI>(o, old);
#pragma unroll
                for (int c = 0; c < CI; ++c) g[c] += old[c];
            }
            st_row<CI>(o, g);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NV; i += blockDim.x) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][i];
        if (i < NW) atomicAdd(p.dw + i, v);
        else if (p.db) atomicAdd(p.db + (i - NW), v);
    }
}
template <int CI, int CO>
cudaError_t launch_conv_bwd_fused_tiny_t(const ConvGradParams& p, cudaStream_t st) {
    const size_t total = (size_t)p.B * p.Tout;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
    if (p.taps == 3) conv_bwd_fused_tiny_kernel<CI, CO, 3><<<blocks, 256, 0, st>>>(p);
    else conv_bwd_fused_tiny_kernel<CI, CO, 1><<<blocks, 256, 0, st>>>(p);
    return cudaGetLastError();
}

template <int CI, int CO>
cudaError_t launch_conv_grad_tiny_t(const ConvGradParams& p, bool weight, cudaStream_t st) {
    const size_t total = (size_t)p.B * (weight ? p.Tout : p.Tin);
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, weight ? 148 * 4 : 148 * 16);
    if (p.taps == 3) {
        if (weight) conv_bwd_weight_tiny_kernel<CI, CO, 3><<<blocks, 256, 0, st>>>(p);
        else conv_bwd_data_tiny_kernel<CI, CO, 3><<<blocks, 256, 0, st>>>(p);
    } else {
        if (weight) conv_bwd_weight_tiny_kernel<CI, CO, 1><<<blocks, 256, 0, st>>>(p);
        else conv_bwd_data_tiny_kernel<CI, CO, 1><<<blocks, 256, 0, st>>>(p);
    }
    return cudaGetLastError();
}
bool conv_grad_tiny_ok(const ConvGradParams& p) {
    auto c124 = [](int c) { return c == 1 || c == 2 || c == 4; };
    return c124(p.Cin) && c124(p.Cout) && (p.taps == 1 || p.taps == 3) && (p.stride == 1 || p.stride == 2);
}
cudaError_t launch_conv_grad_tiny(const ConvGradParams& p, bool weight, cudaStream_t st) {
#define EEGLDM_TINY(CI, CO) if (p.Cin == CI && p.Cout == CO) return launch_conv_grad_tiny_t<CI, CO>(p, weight, st);
    EEGLDM_TINY(1, 1) EEGLDM_TINY(1, 2) EEGLDM_TINY(1, 4) EEGLDM_TINY(2, 1) EEGLDM_TINY(2, 2) EEGLDM_TINY(2, 4)
    EEGLDM_TINY(4, 1) EEGLDM_TINY(4, 2) EEGLDM_TINY(4, 4)
#undef EEGLDM_TINY
    return cudaErrorInvalidValue;
}

// ---- one CTA per sample (tensors of the 2-2-4 autoencoder are <= 6144 elements per sample): GroupNorm statistics, apply and
// ---- SiLU in ONE kernel with the sample in registers (forward), reduce + apply in ONE kernel (backward).  SAMPLE_NT % C == 0, so a
// ---- thread's elements e = tid + SAMPLE_NT*k all belong to channel tid % C: per-thread partials are per (channel, group).
constexpr int SAMPLE_NT = 512;   // threads per sample CTA: 4 CTAs per SM, short phases
constexpr int SAMPLE_NE = 16;   // elements per thread at most (C*T <= 8192)
// Per-channel block totals of one value per thread (thread's channel = tid % C, C in {1,2,4,8} divides the warp size):
// xor-shuffles over the lanes of equal channel, one row per warp in shared memory, then channel c's total in out[c].
// All SAMPLE_NT threads call it; out[] is valid after the trailing __syncthreads().
__device__ __forceinline__ void block_channel_sums(float v, int C, float (*rows)[8] /* [SAMPLE_NT/32 warps][8] */, float* out /* [8] */) {
    for (int o = 16; o >= C; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < C) rows[warp][lane] = v;
    __syncthreads();
    if ((int)threadIdx.x < C) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < SAMPLE_NT / 32; ++w) t += rows[w][threadIdx.x];
        out[threadIdx.x] = t;
    }
    __syncthreads();
}
__device__ __forceinline__ float group_of(const float* chs, int cpg, int g) {
    float s = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) s += chs[c];
    return s;
}
__global__ void __launch_bounds__(SAMPLE_NT) gn_act_fwd_sample_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps, float* __restrict__ a,
                                                                 float* __restrict__ mean_out, float* __restrict__ rstd_out, int C, int T,
                                                                 int G, int silu) {
    __shared__ float rows[SAMPLE_NT / 32][8], chs[8];
    const int b = blockIdx.x, n = C * T, cpg = C / G, ch = threadIdx.x % C, g = ch / cpg;
    const float* xb = x + (size_t)b * n;
    const float inv = 1.f / (float)(cpg * T);
    float v[SAMPLE_NE];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < SAMPLE_NE; ++k) {
        const int e = threadIdx.x + SAMPLE_NT * k;
        v[k] = e < n ? xb[e] : 0.f;
        s += v[k];
    }
    block_channel_sums(s, C, rows, chs);
    const float mu = group_of(chs, cpg, g) * inv;
    __syncthreads();                       // chs is rewritten below
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < SAMPLE_NE; ++k) {
        const int e = threadIdx.x + SAMPLE_NT * k;
        if (e < n) { const float d = v[k] - mu; q = fmaf(d, d, q); }
    }
    block_channel_sums(q, C, rows, chs);
    const float rs = 1.0f / sqrtf(group_of(chs, cpg, g) * inv + eps);   // biased variance, as nn.GroupNorm
    if ((int)threadIdx.x < C && ch % cpg == 0) {
        mean_out[(size_t)b * G + g] = mu;
        rstd_out[(size_t)b * G + g] = rs;
    }
    const float sc = gamma[ch] * rs, sh = beta[ch] - mu * sc;
    float* ab = a + (size_t)b * n;
#pragma unroll
    for (int k = 0; k < SAMPLE_NE; ++k) {
        const int e = threadIdx.x + SAMPLE_NT * k;
        if (e < n) {
            float y = fmaf(sc, v[k], sh);
            if (silu) y = y * sigmoid_f(y);
            ab[e] = y;
        }
    }
}

__global__ void __launch_bounds__(SAMPLE_NT) norm_act_bwd_sample_kernel(const NormGradParams p) {
    __shared__ float rows[SAMPLE_NT / 32][8], c1[8], c2[8];
    const int b = blockIdx.x, C = p.C, n = C * p.T, cpg = C / p.G, ch = threadIdx.x % C, g = ch / cpg;
    const float mu = p.mean[(size_t)b * p.G + g], rs = p.rstd[(size_t)b * p.G + g], ga = p.gamma[ch], be = p.beta[ch];
    const float* xb = p.x + (size_t)b * n;
    const float* db = p.da + (size_t)b * n;
    float gq[SAMPLE_NE], xh[SAMPLE_NE];
    float s1 = 0.f, s2 = 0.f, dg = 0.f, dbv = 0.f;
#pragma unroll
    for (int k = 0; k < SAMPLE_NE; ++k) {
        const int e = threadIdx.x + SAMPLE_NT * k;
        gq[k] = 0.f; xh[k] = 0.f;
        if (e < n) {
            xh[k] = (xb[e] - mu) * rs;
            const float v = fmaf(xh[k], ga, be);
            float dv = db[e];
            if (p.silu) { const float sg = sigmoid_f(v); dv *= sg * (1.f + v * (1.f - sg)); }
            gq[k] = dv * ga;
            s1 += gq[k]; s2 = fmaf(gq[k], xh[k], s2);
            dg = fmaf(dv, xh[k], dg); dbv += dv;
        }
    }
    block_channel_sums(s1, C, rows, c1);
    block_channel_sums(s2, C, rows, c2);
    const float inv = 1.f / (float)(cpg * p.T);
    const float m1 = group_of(c1, cpg, g) * inv, m2 = group_of(c2, cpg, g) * inv;
    __syncthreads();                       // c1 / c2 are rewritten below
    float* dxb = p.dx + (size_t)b * n;
#pragma unroll
    for (int k = 0; k < SAMPLE_NE; ++k) {
        const int e = threadIdx.x + SAMPLE_NT * k;
        if (e < n) {
            const float r = rs * (gq[k] - m1 - xh[k] * m2);
            dxb[e] = p.accumulate ? dxb[e] + r : r;
        }
    }
    block_channel_sums(dg, C, rows, c1);   // per-channel dgamma / dbeta of this sample
    block_channel_sums(dbv, C, rows, c2);
    if ((int)threadIdx.x < C) {
        atomicAdd(p.dgamma + threadIdx.x, c1[threadIdx.x]);
        atomicAdd(p.dbeta + threadIdx.x, c2[threadIdx.x]);
    }
}
bool norm_sample_ok(int C, int T, int G) { return C >= 1 && C <= 8 && SAMPLE_NT % C == 0 && G >= 1 && G <= 8 && C % G == 0 && (size_t)C * T <= SAMPLE_NT * SAMPLE_NE; }

// GroupNorm(+SiLU) backward, pass 1: per (sample, group) means of g = dv*gamma and g*xhat; per-channel dgamma / dbeta.
//   v = xhat*gamma + beta,  a = silu?(v),  dv = da * silu'(v)
// grid (G, B), block 256; requires cpg = C/G <= 256.
__global__ void __launch_bounds__(256) norm_act_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ x,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float* __restrict__ m12 /* [B][G][2] */, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta, int C, int T, int G, int silu) {
    __shared__ float red[2 * 32];
    __shared__ float chs[2 * 256];
    const int g = blockIdx.x, b = blockIdx.y, cpg = C / G;
    const float mu = mean[b * G + g], rs = rstd[b * G + g];
    const int cl = threadIdx.x % cpg, tl = threadIdx.x / cpg, tstep = blockDim.x / cpg;
    const int c = g * cpg + cl;
    const float ga = gamma[c], be = beta[c];
    float s[2] = {0.f, 0.f};
    float dg = 0.f, dbv = 0.f;
    if (tl < tstep)
        for (int t = tl; t < T; t += tstep) {
            const size_t i = ((size_t)b * T + t) * C + c;
            const float xh = (x[i] - mu) * rs;
            const float v = fmaf(xh, ga, be);
            float dv = da[i];
            if (silu) { const float sg = sigmoid_f(v); dv *= sg * (1.f + v * (1.f - sg)); }
            s[0] += dv * ga;
            s[1] += dv * ga * xh;
            dg += dv * xh;
            dbv += dv;
        }
    chs[threadIdx.x] = dg;
    chs[256 + threadIdx.x] = dbv;
    block_reduce_sum<2>(s, red);   // contains __syncthreads
    if (threadIdx.x == 0) {
        const float n = (float)cpg * (float)T;
        m12[((size_t)b * G + g) * 2 + 0] = s[0] / n;
        m12[((size_t)b * G + g) * 2 + 1] = s[1] / n;
    }
    if ((int)threadIdx.x < cpg) {
        float a0 = 0.f, a1 = 0.f;
        for (int j = threadIdx.x; j < tstep * cpg; j += cpg) { a0 += chs[j]; a1 += chs[256 + j]; }
        atomicAdd(dgamma + c, a0);
        atomicAdd(dbeta + c, a1);
    }
}

// pass 2: dx (+)= rstd * (dv*gamma - m1 - xhat*m2)
__global__ void norm_act_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ x, const float* __restrict__ mean,
                                          const float* __restrict__ rstd, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, const float* __restrict__ m12, float* __restrict__ dx, int C,
                                          int T, int G, int silu, int accumulate, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const size_t b = i / ((size_t)C * T);
    const int g = c / (C / G);
    const float mu = mean[b * G + g], rs = rstd[b * G + g], ga = gamma[c];
    const float xh = (x[i] - mu) * rs;
    const float v = fmaf(xh, ga, beta[c]);
    float dv = da[i];
    if (silu) { const float sg = sigmoid_f(v); dv *= sg * (1.f + v * (1.f - sg)); }
    const float r = rs * (dv * ga - m12[(b * G + g) * 2] - xh * m12[(b * G + g) * 2 + 1]);
    dx[i] = accumulate ? dx[i] + r : r;
}

// the same, four channels per thread (C % 4 == 0 and (C/G) % 4 == 0: the four channels share a group)
__global__ void norm_act_bwd_apply4_kernel(const float* __restrict__ da, const float* __restrict__ x, const float* __restrict__ mean,
                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                           const float* __restrict__ beta, const float* __restrict__ m12, float* __restrict__ dx, int C,
                                           int T, int G, int silu, int accumulate, size_t total4) {
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const size_t i = i4 * 4;
    const int c = (int)(i % C);
    const size_t b = i / ((size_t)C * T);
    const int g = c / (C / G);
    const float mu = mean[b * G + g], rs = rstd[b * G + g], m1 = m12[(b * G + g) * 2], m2 = m12[(b * G + g) * 2 + 1];
    const float4 xv = *reinterpret_cast<const float4*>(x + i), dv4 = *reinterpret_cast<const float4*>(da + i);
    const float4 ga4 = *reinterpret_cast<const float4*>(gamma + c), be4 = *reinterpret_cast<const float4*>(beta + c);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
    const float gs[4] = {ga4.x, ga4.y, ga4.z, ga4.w}, bs[4] = {be4.x, be4.y, be4.z, be4.w};
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float xh = (xs[e] - mu) * rs;
        const float v = fmaf(xh, gs[e], bs[e]);
        float dv = ds[e];
        if (silu) { const float sg = sigmoid_f(v); dv *= sg * (1.f + v * (1.f - sg)); }
        r[e] = rs * (dv * gs[e] - m1 - xh * m2);
    }
    float4 o = make_float4(r[0], r[1], r[2], r[3]);
    if (accumulate) {
        const float4 p = *reinterpret_cast<const float4*>(dx + i);
        o = make_float4(p.x + o.x, p.y + o.y, p.z + o.z, p.w + o.w);
    }
    *reinterpret_cast<float4*>(dx + i) = o;
}

__global__ void axpy_kernel(const float* __restrict__ src, float* __restrict__ dst, float alpha, int accumulate, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = accumulate ? dst[i] + alpha * src[i] : alpha * src[i];
}

// L1Loss(mean): loss += sum|r-x| / n ;  drecon = w * sign(r-x)/n
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ r, const float* __restrict__ x, float* __restrict__ dr,
                                                       float* __restrict__ loss, float weight, size_t n) {
    __shared__ float red[32];
    float s[1] = {0.f};
    const float inv = 1.f / (float)n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = r[i] - x[i];
        s[0] += fabsf(d);
        dr[i] = weight * inv * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    block_reduce_sum<1>(s, red);
    if (threadIdx.x == 0) atomicAdd(loss, s[0] * inv);
}

// KL term of train_autoencoderkl.py:210-211 and the reparameterisation z = mu + eps*sigma, sigma = exp(clamp(lv,-30,20)/2).
// forward (dz == null): writes sigma, z.   backward (dz != null): dmu = dz + kl_w*mu/B ; dlv = (dz*eps + kl_w*(sigma - 1/sigma)/B) * sigma/2 (0 outside the clamp)
__global__ void __launch_bounds__(256) latent_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                                                      float* __restrict__ sigma, float* __restrict__ z, const float* __restrict__ dz,
                                                      float* __restrict__ dmu, float* __restrict__ dlv, float* __restrict__ kl_loss,
                                                      float kl_weight, int B, size_t n, const float* __restrict__ dmu_ext,
                                                      const float* __restrict__ dsigma_ext) {
    __shared__ float red[32];
    float s[1] = {0.f};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float raw = lv[i];
        const float sg = expf(fminf(fmaxf(raw, -30.f), 20.f) * 0.5f);
        if (!dz) {
            sigma[i] = sg;
            z[i] = fmaf(eps[i], sg, mu[i]);
            const float m = mu[i];
            s[0] += 0.5f * (m * m + sg * sg - logf(sg * sg) - 1.f);
        } else {
            const float m = mu[i];
            // dmu_ext / dsigma_ext: gradients arriving at the z_mu / z_sigma OUTPUTS of AutoencoderKL.forward (autograd boundary,
            // eegldm_aekl_backward: the caller's own KL term), added to the reparameterisation path
            dmu[i] = dz[i] + kl_weight * m / (float)B + (dmu_ext ? dmu_ext[i] : 0.f);
            const float dsg = dz[i] * eps[i] + kl_weight * (sg - 1.f / sg) / (float)B + (dsigma_ext ? dsigma_ext[i] : 0.f);
            dlv[i] = (raw > -30.f && raw < 20.f) ? dsg * sg * 0.5f : 0.f;
        }
    }
    if (!dz) {
        block_reduce_sum<1>(s, red);
        if (threadIdx.x == 0) atomicAdd(kl_loss, s[0] / (float)B);
    }
}

// torch.optim.Adam (no weight decay, amsgrad off)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float lr,
                            float b1, float b2, float eps, float bc1, float bc2_sqrt, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
}

unsigned blocks_for(size_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// ---- PatchDiscriminator pieces (train_autoencoderkl.py:213-234): BatchNorm1d in training mode over rows h [N = B*T][C]
// ---- (channels-last), LeakyReLU, the least-squares adversarial loss.  Per-channel sums are accumulated in double (one atomic
// ---- per channel per block): var = E[x^2] - E[x]^2 is then free of the fp32 cancellation.
constexpr int BN_ROWS = 64;   // rows per block
__device__ __forceinline__ float lrelu_grad(float v, float slope) { return v > 0.f ? 1.f : slope; }

__global__ void __launch_bounds__(256) bn_sums_kernel(const float* __restrict__ h, size_t N, int C, double* __restrict__ sums) {
    const size_t r0 = (size_t)blockIdx.x * BN_ROWS, r1 = min(N, r0 + BN_ROWS);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f, q = 0.f;
        for (size_t r = r0; r < r1; ++r) { const float v = h[r * C + c]; s += v; q = fmaf(v, v, q); }
        atomicAdd(sums + c, (double)s);
        atomicAdd(sums + C + c, (double)q);
    }
}
// mean / rstd [C]; scale = gamma * rstd, shift = beta - mean * scale replicated over the B samples (the conv prologues index
// [b][c]); running statistics updated n_updates times with momentum (unbiased variance), as nn.BatchNorm1d does per forward call
__global__ void bn_finalize_kernel(const double* __restrict__ sums, size_t N, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, int B, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ run_mean,
                                   float* __restrict__ run_var, float momentum, int n_updates) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mu = sums[c] / (double)N;
    double var = sums[C + c] / (double)N - mu * mu;
    if (var < 0.0) var = 0.0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    mean_out[c] = (float)mu; rstd_out[c] = rs;
    const float sc = gamma[c] * rs, sh = beta[c] - (float)mu * sc;
    for (int b = 0; b < B; ++b) { scale[(size_t)b * C + c] = sc; shift[(size_t)b * C + c] = sh; }
    if (run_mean) {
        const float uv = (float)(N > 1 ? var * (double)N / (double)(N - 1) : var);
        float rm = run_mean[c], rv = run_var[c];
        for (int i = 0; i < n_updates; ++i) { rm = (1.f - momentum) * rm + momentum * (float)mu; rv = (1.f - momentum) * rv + momentum * uv; }
        run_mean[c] = rm; run_var[c] = rv;
    }
}
// backward of a = lrelu(xhat * gamma + beta): pass 1, per-channel sums of dv and dv * xhat (dv = da * lrelu'(v))
__global__ void __launch_bounds__(256) bn_bwd_sums_kernel(const float* __restrict__ da, const float* __restrict__ h,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, size_t N, int C,
                                                           float slope, double* __restrict__ sums) {
    const size_t r0 = (size_t)blockIdx.x * BN_ROWS, r1 = min(N, r0 + BN_ROWS);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float mu = mean[c], rs = rstd[c], ga = gamma[c], be = beta[c];
        float s1 = 0.f, s2 = 0.f;
        for (size_t r = r0; r < r1; ++r) {
            const float xh = (h[r * C + c] - mu) * rs;
            const float dv = da[r * C + c] * lrelu_grad(fmaf(xh, ga, be), slope);
            s1 += dv; s2 = fmaf(dv, xh, s2);
        }
        atomicAdd(sums + c, (double)s1);
        atomicAdd(sums + C + c, (double)s2);
    }
}
// pass 2: dh = gamma * rstd * (dv - mean(dv) - xhat * mean(dv * xhat));  block 0 also adds dgamma = sum dv*xhat, dbeta = sum dv
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ h,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const double* __restrict__ sums, size_t N, int C, float slope,
                                                            float* __restrict__ dh, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const size_t r0 = (size_t)blockIdx.x * BN_ROWS, r1 = min(N, r0 + BN_ROWS);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float mu = mean[c], rs = rstd[c], ga = gamma[c], be = beta[c];
        const float m1 = (float)(sums[c] / (double)N), m2 = (float)(sums[C + c] / (double)N);
        for (size_t r = r0; r < r1; ++r) {
            const float xh = (h[r * C + c] - mu) * rs;
            const float dv = da[r * C + c] * lrelu_grad(fmaf(xh, ga, be), slope);
            dh[r * C + c] = ga * rs * (dv - m1 - xh * m2);
        }
        if (blockIdx.x == 0 && dgamma) { dgamma[c] += (float)sums[C + c]; dbeta[c] += (float)sums[c]; }
    }
}
// dh = da * lrelu'(h)   (activation without a norm: the discriminator's initial_conv)
__global__ void lrelu_bwd_kernel(const float* __restrict__ da, const float* __restrict__ h, float slope, float* __restrict__ dh, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dh[i] = da[i] * lrelu_grad(h[i], slope);
}
// a = lrelu(scale[c] * h + shift[c])   (materialised conv input for the weight-gradient kernels; scale == null: a = lrelu(h))
__global__ void affine_lrelu_kernel(const float* __restrict__ h, const float* __restrict__ scale, const float* __restrict__ shift, int C,
                                    float slope, float* __restrict__ a, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C);
    const float v = scale ? fmaf(scale[c], h[i], shift[c]) : h[i];
    a[i] = v > 0.f ? v : slope * v;
}
// PatchAdversarialLoss(criterion="least_squares"): y = lrelu_{act_slope}(logit) (act_slope = 1: no activation);
// loss += loss_weight * mean((y - target)^2);  dlogit = grad_weight * 2 (y - target) / n * lrelu'(logit)
__global__ void __launch_bounds__(256) adv_loss_kernel(const float* __restrict__ logits, size_t n, float target, float act_slope,
                                                        float loss_weight, float* __restrict__ loss, float grad_weight,
                                                        float* __restrict__ dlogits) {
    __shared__ float red[32];
    float s[1] = {0.f};
    const float inv = 1.f / (float)n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float z = logits[i];
        const float y = z > 0.f ? z : act_slope * z;
        const float d = y - target;
        s[0] = fmaf(d, d, s[0]);
        if (dlogits) dlogits[i] = grad_weight * 2.f * d * inv * (z > 0.f ? 1.f : act_slope);
    }
    block_reduce_sum<1>(s, red);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, loss_weight * s[0] * inv);
}
// Weights of the convolution that computes a conv's input gradient, built from the forward weights (SIMT image
// [(ci*3 + k)][Cout], k = 3, padding 1) so that the data gradient runs through the forward conv kernels:
//   stride 1:  da = conv3_same(dy, Wd),  Wd[(co*3 + k)][ci] = W[(ci*3 + 2-k)][co]                       (Cin_d = Cout, Cout_d = Cin)
//   stride 2:  rows 2t, 2t+1 of the input seen as ONE row of 2 Cin channels (channels-last: the same memory):
//              dz[t][0:Cin] = W1^T dy[t],  dz[t][Cin:2Cin] = W2^T dy[t] + W0^T dy[t+1]                     (Cin_d = Cout, Cout_d = 2 Cin)
__global__ void dgrad_weights_kernel(const float* __restrict__ w, int Cin, int Cout, int stride, float* __restrict__ wd) {
    const int cod = stride == 1 ? Cin : 2 * Cin;
    const size_t total = (size_t)Cout * 3 * cod;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c2 = (int)(i % cod);
    const int k = (int)((i / cod) % 3);
    const int co = (int)(i / ((size_t)cod * 3));
    float v = 0.f;
    if (stride == 1) v = w[(size_t)(c2 * 3 + (2 - k)) * Cout + co];
    else if (k == 1) v = c2 < Cin ? w[(size_t)(c2 * 3 + 1) * Cout + co] : w[(size_t)((c2 - Cin) * 3 + 2) * Cout + co];
    else if (k == 2) v = c2 >= Cin ? w[(size_t)((c2 - Cin) * 3 + 0) * Cout + co] : 0.f;
    wd[i] = v;
}

}  // namespace

cudaError_t launch_norm_act_fwd(const float* x, const float* scale, const float* shift, float* a, int B, int T, int C, int silu,
                                cudaStream_t st) {
    const size_t total = (size_t)B * T * C;
    if (!total) return cudaSuccess;
    const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(scale) |
                        reinterpret_cast<uintptr_t>(shift)) & 15) == 0;
    if (C % 4 == 0 && al16) norm_act_fwd4_kernel<<<blocks_for(total / 4), 256, 0, st>>>(x, scale, shift, a, C, T, silu, total / 4);
    else norm_act_fwd_kernel<<<blocks_for(total), 256, 0, st>>>(x, scale, shift, a, C, T, silu, total);
    g_launch_count += 1;
    return cudaGetLastError();
}

// weight gradient and (if p.da) input gradient of one conv; one fused launch for the tiny stride-1 same-length case
cudaError_t launch_conv_bwd(const ConvGradParams& p, cudaStream_t st) {
    if (p.B <= 0 || p.Tout <= 0) return cudaSuccess;
    if (conv_grad_tiny_ok(p) && p.stride == 1 && !p.ups && p.Tin == p.Tout && p.pad == p.taps / 2) {
        g_launch_count += 1;
#define EEGLDM_TINY(CI, CO) if (p.Cin == CI && p.Cout == CO) return launch_conv_bwd_fused_tiny_t<CI, CO>(p, st);
        EEGLDM_TINY(1, 1) EEGLDM_TINY(1, 2) EEGLDM_TINY(1, 4) EEGLDM_TINY(2, 1) EEGLDM_TINY(2, 2) EEGLDM_TINY(2, 4)
        EEGLDM_TINY(4, 1) EEGLDM_TINY(4, 2) EEGLDM_TINY(4, 4)
#undef EEGLDM_TINY
        return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaSuccess;
    if (p.dres) e = launch_axpy(p.dy, p.dres, 1.f, p.dres_accumulate, (size_t)p.B * p.Tout * p.Cout, st);
    if (e == cudaSuccess) e = launch_conv_bwd_weight(p, st);
    if (e == cudaSuccess && p.da) e = launch_conv_bwd_data(p, st);
    return e;
}

cudaError_t launch_conv_bwd_data(const ConvGradParams& p, cudaStream_t st) {
    const size_t total = (size_t)p.B * p.Tin * p.Cin;
    if (!total) return cudaSuccess;
    if (conv_grad_tiny_ok(p)) {
        g_launch_count += 1;
        return launch_conv_grad_tiny(p, false, st);
    }
    conv_bwd_data_kernel<<<blocks_for(total), 256, 0, st>>>(p.dy, p.w, p.da, p.Cin, p.Cout, p.taps, p.stride, p.pad, p.ups, p.Tin, p.Tc,
                                                            p.Tout, p.accumulate, total);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_conv_bwd_weight(const ConvGradParams& p, cudaStream_t st) {
    if (p.B <= 0 || p.Tout <= 0) return cudaSuccess;
    if (conv_grad_tiny_ok(p)) {
        g_launch_count += 1;
        return launch_conv_grad_tiny(p, true, st);
    }
    int WG_P = 64;
    auto smem_for = [&](int P) { return ((size_t)P * p.Cout + (size_t)((P - 1) * p.stride + p.taps) * p.Cin) * sizeof(float); };
    // narrow convs (register form of the kernel): small tiles so that several blocks share an SM -- their loads are latency-bound
    const size_t smem_cap = (size_t)p.Cin * p.taps * p.Cout <= 2048 ? 40 * 1024 : 200 * 1024;
    while (WG_P > 8 && smem_for(WG_P) > smem_cap) WG_P >>= 1;
    while (WG_P > 1 && smem_for(WG_P) > 200 * 1024) WG_P >>= 1;
    const size_t smem = smem_for(WG_P);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(conv_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr = 200 * 1024;
    }
    const int ntiles_t = (p.Tout + WG_P - 1) / WG_P;
    const long long ntiles = (long long)ntiles_t * p.B;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
    const unsigned grid = (unsigned)std::min<long long>(ntiles, 148LL * per_sm);
    conv_bwd_weight_kernel<<<grid, 256, smem, st>>>(p.dy, p.a, p.dw, p.db, p.Cin, p.Cout, p.taps, p.stride, p.pad, p.ups, p.Tin, p.Tc, p.Tout, WG_P,
                                                    ntiles_t, (int)ntiles);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_norm_act_bwd(const NormGradParams& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    if (norm_sample_ok(p.C, p.T, p.G)) {
        norm_act_bwd_sample_kernel<<<p.B, SAMPLE_NT, 0, st>>>(p);
        g_launch_count += 1;
        return cudaGetLastError();
    }
    if (p.C / p.G > 256) return cudaErrorInvalidValue;
    dim3 grid(p.G, p.B);
    norm_act_bwd_reduce_kernel<<<grid, 256, 0, st>>>(p.da, p.x, p.mean, p.rstd, p.gamma, p.beta, p.m12, p.dgamma, p.dbeta, p.C, p.T, p.G, p.silu);
    const size_t total = (size_t)p.B * p.T * p.C;
    const bool al16 = ((reinterpret_cast<uintptr_t>(p.da) | reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.dx) |
                        reinterpret_cast<uintptr_t>(p.gamma) | reinterpret_cast<uintptr_t>(p.beta)) & 15) == 0;
    if (p.C % 4 == 0 && (p.C / p.G) % 4 == 0 && al16)
        norm_act_bwd_apply4_kernel<<<blocks_for(total / 4), 256, 0, st>>>(p.da, p.x, p.mean, p.rstd, p.gamma, p.beta, p.m12, p.dx, p.C, p.T, p.G,
                                                                          p.silu, p.accumulate, total / 4);
    else
        norm_act_bwd_apply_kernel<<<blocks_for(total), 256, 0, st>>>(p.da, p.x, p.mean, p.rstd, p.gamma, p.beta, p.m12, p.dx, p.C, p.T, p.G, p.silu,
                                                                    p.accumulate, total);
    g_launch_count += 2;
    return cudaGetLastError();
}

// GroupNorm statistics + apply (+ SiLU) of a training forward pass: a = silu?(GN(x)), mean / rstd [B][G] saved for the backward
// pass.  Tiny tensors: one kernel; otherwise statistics (launch_groupnorm with p prepared by the caller) + launch_norm_act_fwd.
cudaError_t launch_gn_act_fwd(const GnParams& p, float* a, int silu, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    if (!p.src1 && norm_sample_ok(p.C0, p.T, p.G) && p.mean_out && p.rstd_out) {
        gn_act_fwd_sample_kernel<<<p.B, SAMPLE_NT, 0, st>>>(p.src0, p.gamma, p.beta, p.eps, a, p.mean_out, p.rstd_out, p.C0, p.T, p.G, silu);
        g_launch_count += 1;
        return cudaGetLastError();
    }
    cudaError_t e = launch_groupnorm(p, st);
    if (e != cudaSuccess) return e;
    return launch_norm_act_fwd(p.src0, p.scale, p.shift, a, p.B, p.T, p.C0, silu, st);
}

cudaError_t launch_axpy(const float* src, float* dst, float alpha, int accumulate, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    axpy_kernel<<<blocks_for(n), 256, 0, st>>>(src, dst, alpha, accumulate, n);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_l1_loss(const float* r, const float* x, float* dr, float* loss, float weight, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
    l1_loss_kernel<<<blocks, 256, 0, st>>>(r, x, dr, loss, weight, n);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_latent(const float* mu, const float* lv, const float* eps, float* sigma, float* z, const float* dz, float* dmu,
                          float* dlv, float* kl_loss, float kl_weight, int B, size_t n, cudaStream_t st, const float* dmu_ext,
                          const float* dsigma_ext) {
    if (!n) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
    latent_kernel<<<blocks, 256, 0, st>>>(mu, lv, eps, sigma, z, dz, dmu, dlv, kl_loss, kl_weight, B, n, dmu_ext, dsigma_ext);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_adam(float* p, const float* g, float* m, float* v, float lr, float b1, float b2, float eps, int step, size_t n,
                        cudaStream_t st) {
    if (!n) return cudaSuccess;
    const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
    adam_kernel<<<blocks_for(n), 256, 0, st>>>(p, g, m, v, lr, b1, b2, eps, bc1, sqrtf(bc2), n);
    g_launch_count += 1;
    return cudaGetLastError();
}

}  // namespace eegldm

namespace eegldm {

cudaError_t launch_bn_stats(const float* h, size_t N, int C, const float* gamma, const float* beta, float eps, int B, double* sums,
                            float* scale, float* shift, float* mean, float* rstd, float* run_mean, float* run_var, float momentum,
                            int n_updates, cudaStream_t st) {
    if (!N || C <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(double), st);
    if (e != cudaSuccess) return e;
    bn_sums_kernel<<<(unsigned)((N + BN_ROWS - 1) / BN_ROWS), 256, 0, st>>>(h, N, C, sums);
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, N, C, gamma, beta, eps, B, scale, shift, mean, rstd, run_mean, run_var, momentum,
                                                        n_updates);
    g_launch_count += 2;
    return cudaGetLastError();
}

cudaError_t launch_bn_lrelu_bwd(const float* da, const float* h, const float* mean, const float* rstd, const float* gamma,
                                const float* beta, size_t N, int C, float slope, double* sums, float* dh, float* dgamma, float* dbeta,
                                cudaStream_t st) {
    if (!N || C <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const unsigned nb = (unsigned)((N + BN_ROWS - 1) / BN_ROWS);
    bn_bwd_sums_kernel<<<nb, 256, 0, st>>>(da, h, mean, rstd, gamma, beta, N, C, slope, sums);
    bn_bwd_apply_kernel<<<nb, 256, 0, st>>>(da, h, mean, rstd, gamma, beta, sums, N, C, slope, dh, dgamma, dbeta);
    g_launch_count += 2;
    return cudaGetLastError();
}

cudaError_t launch_lrelu_bwd(const float* da, const float* h, float slope, float* dh, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    lrelu_bwd_kernel<<<blocks_for(n), 256, 0, st>>>(da, h, slope, dh, n);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_affine_lrelu(const float* h, const float* scale, const float* shift, int C, float slope, float* a, size_t n,
                                cudaStream_t st) {
    if (!n) return cudaSuccess;
    affine_lrelu_kernel<<<blocks_for(n), 256, 0, st>>>(h, scale, shift, C, slope, a, n);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_adv_loss(const float* logits, size_t n, float target, float act_slope, float loss_weight, float* loss,
                            float grad_weight, float* dlogits, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 4);
    adv_loss_kernel<<<blocks, 256, 0, st>>>(logits, n, target, act_slope, loss_weight, loss, grad_weight, dlogits);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_dgrad_weights(const float* w, int Cin, int Cout, int stride, float* wd, cudaStream_t st) {
    const size_t total = (size_t)Cout * 3 * (stride == 1 ? Cin : 2 * Cin);
    if (!total) return cudaSuccess;
    dgrad_weights_kernel<<<blocks_for(total), 256, 0, st>>>(w, Cin, Cout, stride, wd);
    g_launch_count += 1;
    return cudaGetLastError();
}

}  // namespace eegldm
