// Kernels of the latent-diffusion training step that are not convolutions (training/training.py:402-450 around
// models/unet.py): the attention block's forward / backward GEMMs and softmax (unet.py:107-125), the time-embedding MLP's
// backward (unet.py:372-377, 277-285), nearest / average-pool resample adjoints (unet.py:177-224), channel concat / split
// (unet.py:553), DDPM add_noise / get_velocity (training.py:429-436) and the mean-squared-error loss (training.py:437).
// Everything here is fp32 SIMT: the convolutions (96 % of the step's FLOPs) run on the tensor pipe (conv_tc.cu, train_tc.cu).
#include <algorithm>

#include "kernels.cuh"

namespace eegldm {
namespace {

unsigned nblk(size_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

// ------------------------------------------------------------------------------------------------ batched GEMM
// C[m][n] (+)= alpha * sum_k A[m][k] B[k][n] with arbitrary element strides; batch index = outer * nb_inner + inner.
// 64 x 64 tile, 16-deep k-steps, 256 threads x (4 x 4) outputs.
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256) bgemm_kernel(const BGemm g) {
    __shared__ float As[GK][GT + 4], Bs[GK][GT + 4];
    const int bz = blockIdx.z, bo = bz / g.nb_inner, bi = bz % g.nb_inner;
    const float* A = g.A + bo * g.a_bo + bi * g.a_bi;
    const float* B = g.B + bo * g.b_bo + bi * g.b_bi;
    float* C = g.C + bo * g.c_bo + bi * g.c_bi;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};
    // load mappings: consecutive threads along the unit-stride axis of each operand
    const bool a_mfast = g.a_m == 1, b_nfast = g.b_n == 1;
    for (int k0 = 0; k0 < g.K; k0 += GK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = tid + 256 * j;
            const int mm = a_mfast ? (e & 63) : (e >> 4), kk = a_mfast ? (e >> 6) : (e & 15);
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < g.K) ? A[(long long)m * g.a_m + (long long)k * g.a_k] : 0.f;
            const int nn = b_nfast ? (e & 63) : (e >> 4), kb = b_nfast ? (e >> 6) : (e & 15);
            const int n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < g.N && k2 < g.K) ? B[(long long)k2 * g.b_k + (long long)n * g.b_n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < g.M && n < g.N) {
                float* c = C + (long long)m * g.c_m + (long long)n * g.c_n;
                const float v = g.alpha * acc[i][j];
                *c = g.accumulate ? *c + v : v;
            }
        }
}

// ------------------------------------------------------------------------------------------------ softmax over rows of length n
// one warp per row; in place
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ s, size_t rows, int n) {
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float* r = s + row * n;
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, r[i]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) { const float e = __expf(r[i] - mx); r[i] = e; sum += e; }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int i = lane; i < n; i += 32) r[i] *= inv;
}
// dS = P o (dP - sum_s dP o P), in place on dP
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float* __restrict__ p, float* __restrict__ dp, size_t rows, int n) {
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* pr = p + row * n;
    float* dr = dp + row * n;
    float dot = 0.f;
    for (int i = lane; i < n; i += 32) dot = fmaf(pr[i], dr[i], dot);
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int i = lane; i < n; i += 32) dr[i] = pr[i] * (dr[i] - dot);
}

// ------------------------------------------------------------------------------------------------ per-sample column sums
// out[b * out_stride + c] (+)= sum_t x[b][t][c]        grid (ceil(C/32), B), block 256 = 8 row groups x 32 channels
__global__ void __launch_bounds__(256) rowsum_bt_kernel(const float* __restrict__ x, int T, int C, float* __restrict__ out, int out_stride,
                                                         int accumulate) {
    __shared__ float red[8][33];
    const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), rg = threadIdx.x >> 5;
    float s = 0.f;
    if (c < C)
        for (int t = rg; t < T; t += 8) s += x[((size_t)b * T + t) * C + c];
    red[rg][threadIdx.x & 31] = s;
    __syncthreads();
    if (rg == 0 && c < C) {
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) tot += red[j][threadIdx.x];
        float* o = out + (size_t)b * out_stride + c;
        *o = accumulate ? *o + tot : tot;
    }
}

// ------------------------------------------------------------------------------------------------ resample adjoints
// forward y = resample(x): AvgPool1d(2,2) (y[t] = (x[2t] + x[2t+1]) / 2) or nearest x2 (y[2t] = y[2t+1] = x[t])
__global__ void resample_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int Tin, int C, int mode, int accumulate, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over dx [B][Tin][C]
    if (i >= total) return;
    const int c = (int)(i % C);
    const size_t bt = i / C;
    const int t = (int)(bt % Tin);
    const size_t b = bt / Tin;
    float v;
    if (mode == RS_AVGPOOL2) v = 0.5f * dy[((size_t)b * (Tin / 2) + (t >> 1)) * C + c];
    else v = dy[((size_t)b * (2 * Tin) + 2 * t) * C + c] + dy[((size_t)b * (2 * Tin) + 2 * t + 1) * C + c];
    dx[i] = accumulate ? dx[i] + v : v;
}

// ------------------------------------------------------------------------------------------------ channel concat / split
__global__ void concat_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1, float* __restrict__ out, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int C = C0 + C1, c = (int)(i % C);
    const size_t r = i / C;
    out[i] = c < C0 ? x0[r * C0 + c] : x1[r * C1 + (c - C0)];
}
__global__ void split_add_kernel(const float* __restrict__ dcat, float* __restrict__ d0, int C0, int acc0, float* __restrict__ d1, int C1, int acc1,
                                 size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int C = C0 + C1, c = (int)(i % C);
    const size_t r = i / C;
    const float v = dcat[i];
    if (c < C0) { float* o = d0 + r * C0 + c; *o = acc0 ? *o + v : v; }
    else { float* o = d1 + r * C1 + (c - C0); *o = acc1 ? *o + v : v; }
}

// ------------------------------------------------------------------------------------------------ SiLU on small tensors (time MLP)
__global__ void silu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = x[i]; y[i] = v * sigmoidf_(v); }
}
__global__ void silu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = x[i], sg = sigmoidf_(v); dx[i] = dy[i] * sg * (1.f + v * (1.f - sg)); }
}

// ------------------------------------------------------------------------------------------------ DDPM training inputs
// noisy = sqrt(abar_t) z0 + sqrt(1 - abar_t) eps (add_noise, training.py:429); target = eps (epsilon) or
// sqrt(abar_t) eps - sqrt(1 - abar_t) z0 (get_velocity, training.py:432-434); t_f[b] = float(t[b]) for the timestep embedding
__global__ void ldm_inputs_kernel(const float* __restrict__ z0, const float* __restrict__ eps, const long long* __restrict__ t,
                                  const float* __restrict__ acp, int n_train, float* __restrict__ noisy, float* __restrict__ target,
                                  float* __restrict__ t_f, int v_pred, size_t per, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t b = i / per;
    long long ti = t[b];
    ti = ti < 0 ? 0 : (ti >= n_train ? n_train - 1 : ti);
    const float ab = acp[ti];
    const float sa = sqrtf(ab), sb = sqrtf(1.f - ab);
    const float z = z0[i], e = eps[i];
    noisy[i] = sa * z + sb * e;
    target[i] = v_pred ? sa * e - sb * z : e;
    if (i % per == 0) t_f[b] = (float)t[b];
}

// F.mse_loss (mean): loss += sum (p - y)^2 / n;  dp = 2 (p - y) / n
__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ p, const float* __restrict__ y, float* __restrict__ dp,
                                                        float* __restrict__ loss, size_t n) {
    __shared__ float red[8];
    float s = 0.f;
    const float inv = 1.f / (float)n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = p[i] - y[i];
        s = fmaf(d, d, s);
        dp[i] = 2.f * d * inv;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int j = 0; j < 8; ++j) tot += red[j];
        atomicAdd(loss, tot * inv);
    }
}

// data-gradient weights for any tap count (SIMT images): Wd[(co*taps + k)][ci] = W[(ci*taps + taps-1-k)][co]
__global__ void dgrad_weights_any_kernel(const float* __restrict__ w, int Cin, int Cout, int taps, float* __restrict__ wd) {
    const size_t total = (size_t)Cout * taps * Cin;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ci = (int)(i % Cin);
    const int k = (int)((i / Cin) % taps);
    const int co = (int)(i / ((size_t)Cin * taps));
    wd[i] = w[((size_t)ci * taps + (taps - 1 - k)) * Cout + co];
}

}  // namespace

cudaError_t launch_bgemm(const BGemm& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return cudaSuccess;
    dim3 grid((g.N + GT - 1) / GT, (g.M + GT - 1) / GT, g.batch);
    bgemm_kernel<<<grid, 256, 0, st>>>(g);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_softmax_rows(float* s, size_t rows, int n, cudaStream_t st) {
    if (!rows) return cudaSuccess;
    softmax_rows_kernel<<<nblk(rows, 8), 256, 0, st>>>(s, rows, n);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_softmax_bwd_rows(const float* p, float* dp, size_t rows, int n, cudaStream_t st) {
    if (!rows) return cudaSuccess;
    softmax_bwd_rows_kernel<<<nblk(rows, 8), 256, 0, st>>>(p, dp, rows, n);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_rowsum_bt(const float* x, int B, int T, int C, float* out, int out_stride, int accumulate, cudaStream_t st) {
    if (B <= 0 || C <= 0) return cudaSuccess;
    dim3 grid((C + 31) / 32, B);
    rowsum_bt_kernel<<<grid, 256, 0, st>>>(x, T, C, out, out_stride, accumulate);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_resample_bwd(const float* dy, float* dx, int B, int Tin, int C, int mode, int accumulate, cudaStream_t st) {
    const size_t total = (size_t)B * Tin * C;
    if (!total) return cudaSuccess;
    if (mode != RS_AVGPOOL2 && mode != RS_NEAREST2) return cudaErrorInvalidValue;
    resample_bwd_kernel<<<nblk(total), 256, 0, st>>>(dy, dx, Tin, C, mode, accumulate, total);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_concat(const float* x0, int C0, const float* x1, int C1, float* out, size_t rows, cudaStream_t st) {
    const size_t total = rows * (size_t)(C0 + C1);
    if (!total) return cudaSuccess;
    concat_kernel<<<nblk(total), 256, 0, st>>>(x0, C0, x1, C1, out, total);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_split_add(const float* dcat, float* d0, int C0, int acc0, float* d1, int C1, int acc1, size_t rows, cudaStream_t st) {
    const size_t total = rows * (size_t)(C0 + C1);
    if (!total) return cudaSuccess;
    split_add_kernel<<<nblk(total), 256, 0, st>>>(dcat, d0, C0, acc0, d1, C1, acc1, total);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_silu_fwd(const float* x, float* y, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    silu_fwd_kernel<<<nblk(n), 256, 0, st>>>(x, y, n);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_silu_bwd(const float* dy, const float* x, float* dx, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    silu_bwd_kernel<<<nblk(n), 256, 0, st>>>(dy, x, dx, n);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_ldm_inputs(const float* z0, const float* eps, const long long* t, const float* acp, int n_train, float* noisy, float* target,
                              float* t_f, int v_pred, int B, size_t per, cudaStream_t st) {
    const size_t total = (size_t)B * per;
    if (!total) return cudaSuccess;
    ldm_inputs_kernel<<<nblk(total), 256, 0, st>>>(z0, eps, t, acp, n_train, noisy, target, t_f, v_pred, per, total);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_mse_loss(const float* p, const float* y, float* dp, float* loss, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
    mse_loss_kernel<<<blocks, 256, 0, st>>>(p, y, dp, loss, n);
    g_launch_count += 1;
    return cudaGetLastError();
}
cudaError_t launch_dgrad_weights_any(const float* w, int Cin, int Cout, int taps, float* wd, cudaStream_t st) {
    const size_t total = (size_t)Cout * taps * Cin;
    if (!total) return cudaSuccess;
    dgrad_weights_any_kernel<<<nblk(total), 256, 0, st>>>(w, Cin, Cout, taps, wd);
    g_launch_count += 1;
    return cudaGetLastError();
}

}  // namespace eegldm
