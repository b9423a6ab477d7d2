// Spectral (Jukebox) loss: generative.losses.JukeboxLoss(spatial_dims=1, reduction="sum") [monai-generative], as called at
// src/train_autoencoderkl.py:158,208:   A(x) = |fftn(x, dim=(1,2), norm="ortho")| ;  loss = reduce((A(target) - A(input))^2).
// For the reference's single-channel signals the (1,2)-FFT is a 1-D FFT of length N per sample.  Real input => Hermitian
// spectrum, so one batched cuFFT R2C per signal and Hermitian weights (1 for bins 0 and N/2, 2 otherwise) replace the
// reference's full complex transform; the backward pass is one C2R:
//   dL/dinput = C2R(H) / sqrt(N),  H_k = -2 (A_t[k] - A_i[k]) * F_i[k] / |F_i[k]|     (F unnormalised, A = |F| / sqrt(N))
// cuFFT is loaded with dlopen at first use so that libeegldm.so has no link-time dependency on it.
#include "cufft_api.h"

#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "kernels.cuh"

namespace eegldm {
namespace {

// cuFFT plans and their work areas belong to the device they were created on, and a scratch buffer may only be reused by
// work ordered on the same stream: plans are keyed by (device, stream, N, batch), scratch by (device, stream).  The mutex covers the
// host-side bookkeeping; launches on different streams use different scratch and do not race.
struct SpectralScratch {
    float2* freq = nullptr; size_t freq_cap = 0;    // [2][B][N/2+1]
    float* time = nullptr; size_t time_cap = 0;     // [B][N]
};
struct SpectralState {
    CufftApi& api = cufft_api();
    std::map<std::tuple<int, cudaStream_t, int, int>, std::pair<cufftHandle, cufftHandle>> plans;   // (device, stream, N, batch) -> (r2c, c2r)
    std::map<std::pair<int, cudaStream_t>, SpectralScratch> scratch;                   // (device, stream)
    std::mutex mu;
};
SpectralState g_spec;

// per bin: amplitude difference, weighted squared error, gradient spectrum (in place over Fi)
__global__ void __launch_bounds__(256) spectral_bins_kernel(float2* __restrict__ Fi, const float2* __restrict__ Ft, float* __restrict__ loss,
                                                             int N, size_t nbins_total, float loss_scale, int want_grad) {
    __shared__ float red[32];
    const int nb = N / 2 + 1;
    const float inv_sqrt_n = rsqrtf((float)N);
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins_total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % nb);
        const float2 fi = Fi[i], ft = Ft[i];
        const float mi = sqrtf(fi.x * fi.x + fi.y * fi.y), mt = sqrtf(ft.x * ft.x + ft.y * ft.y);
        const float d = (mt - mi) * inv_sqrt_n;                       // A_t - A_i
        const float wk = (k == 0 || 2 * k == N) ? 1.f : 2.f;         // Hermitian twin
        s += wk * d * d;
        if (want_grad) {
            const float c = mi > 0.f ? -2.f * d / mi : 0.f;          // d|F|/dF = F/|F|; sqrt at 0 has no usable gradient
            Fi[i] = make_float2(c * fi.x, c * fi.y);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
        s = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) atomicAdd(loss, s * loss_scale);
    }
}

__global__ void scale_store_kernel(const float* __restrict__ src, float* __restrict__ dst, float alpha, int accumulate, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = accumulate ? dst[i] + alpha * src[i] : alpha * src[i];
}

}  // namespace

CufftApi& cufft_api() { static CufftApi api; return api; }

// loss_dev += loss_weight * reduce(...);  grad_dev (+)= grad_weight * dloss/dinput   (grad_dev may be null)
// reduction: 0 = sum, 1 = mean.  Returns 0 on success, 1 = cuFFT unavailable / failed (message in *err), else a cudaError_t.
int spectral_loss(const float* input, const float* target, int B, int N, int reduction, float loss_weight, float* loss_dev,
                  float* grad_dev, float grad_weight, int grad_accumulate, cudaStream_t st, std::string* err) {
    if (B <= 0 || N <= 0) return 0;
    std::lock_guard<std::mutex> lock(g_spec.mu);
    if (!g_spec.api.load()) { if (err) *err = g_spec.api.err; return 1; }
    const int nb = N / 2 + 1;
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return (int)e; }
    auto key = std::make_tuple(dev, st, N, B);   // a plan owns a work area: one per stream
    auto it = g_spec.plans.find(key);
    if (it == g_spec.plans.end()) {
        cufftHandle r2c, c2r;
        int n[1] = {N};
        if (g_spec.api.PlanMany(&r2c, 1, n, nullptr, 1, N, nullptr, 1, nb, CUFFT_R2C, B) != CUFFT_SUCCESS ||
            g_spec.api.PlanMany(&c2r, 1, n, nullptr, 1, nb, nullptr, 1, N, CUFFT_C2R, B) != CUFFT_SUCCESS) {
            if (err) *err = "cufftPlanMany failed";
            return 1;
        }
        it = g_spec.plans.emplace(key, std::make_pair(r2c, c2r)).first;
    }
    const size_t nfreq = (size_t)2 * B * nb, ntime = (size_t)B * N;
    SpectralScratch& sc = g_spec.scratch[std::make_pair(dev, st)];
    if (nfreq > sc.freq_cap) {   // growth only (cudaFree waits for the device: once per shape)
        if (sc.freq) cudaFree(sc.freq);
        cudaError_t e = cudaMalloc((void**)&sc.freq, nfreq * sizeof(float2));
        if (e != cudaSuccess) { sc.freq = nullptr; sc.freq_cap = 0; return (int)e; }
        sc.freq_cap = nfreq;
    }
    if (grad_dev && ntime > sc.time_cap) {
        if (sc.time) cudaFree(sc.time);
        cudaError_t e = cudaMalloc((void**)&sc.time, ntime * sizeof(float));
        if (e != cudaSuccess) { sc.time = nullptr; sc.time_cap = 0; return (int)e; }
        sc.time_cap = ntime;
    }
    float2* Fi = sc.freq;
    float2* Ft = sc.freq + (size_t)B * nb;
    const cufftHandle r2c = it->second.first, c2r = it->second.second;
    if (g_spec.api.SetStream(r2c, st) != CUFFT_SUCCESS || g_spec.api.SetStream(c2r, st) != CUFFT_SUCCESS ||
        g_spec.api.ExecR2C(r2c, const_cast<float*>(input), reinterpret_cast<cufftComplex*>(Fi)) != CUFFT_SUCCESS ||
        g_spec.api.ExecR2C(r2c, const_cast<float*>(target), reinterpret_cast<cufftComplex*>(Ft)) != CUFFT_SUCCESS) {
        if (err) *err = "cufftExecR2C failed";
        return 1;
    }
    const size_t nbins = (size_t)B * nb;
    const float red = reduction == 1 ? 1.f / ((float)B * (float)N) : 1.f;
    const unsigned blocks = (unsigned)std::min<size_t>((nbins + 255) / 256, 148 * 8);
    spectral_bins_kernel<<<blocks, 256, 0, st>>>(Fi, Ft, loss_dev, N, nbins, loss_weight * red, grad_dev != nullptr);
    g_launch_count += 1;
    if (grad_dev) {
        if (g_spec.api.ExecC2R(c2r, reinterpret_cast<cufftComplex*>(Fi), sc.time) != CUFFT_SUCCESS) {
            if (err) *err = "cufftExecC2R failed";
            return 1;
        }
        scale_store_kernel<<<(unsigned)((ntime + 255) / 256), 256, 0, st>>>(sc.time, grad_dev, grad_weight * red * rsqrtf((float)N),
                                                                          grad_accumulate, ntime);
        g_launch_count += 1;
    }
    return (int)cudaGetLastError();
}

}  // namespace eegldm
