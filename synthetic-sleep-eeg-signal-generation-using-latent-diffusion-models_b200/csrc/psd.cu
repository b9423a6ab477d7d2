// Output tail of the sampling scripts (src/sample_trials.py:169-197, src/sample_trials_ddpm.py:105-128, src/util.py:66-89):
//   cropped = sample[:, :, 36:-36];  np.save(sample_i.npy);  spectrum = EpochsArray(cropped, sfreq=100).compute_psd(fmax=18);
//   psds = 10 * log10(spectrum.average().get_data());  np.save(psd_list_i.npy)
// The reference runs this per window on the host through MNE; at thousands of windows per second that tail would dominate, so
// it is one batched pass here: strided device->host copy of the cropped windows, a native .npy writer, and the PSD of ALL
// windows on the device with one batched cuFFT call.
//
// PSD estimators [upstream: mne.time_frequency, version unpinned (requirements.txt), not installed here -> parity unpinned;
// restated in oracle/psd.py from the published algorithm]:
//   multitaper (Epochs.compute_psd's default method): tapers = scipy.signal.windows.dpss(N, 4, 8, sym=False, norm=2), keep those
//     with concentration > 0.9 (low_bias); per taper X_k = rfft((x - mean) * w_k), DC (and Nyquist for even N) / sqrt 2;
//     psd = 2 * sum_k lambda_k |X_k|^2 / sum_k lambda_k      (normalization="length"; "full" divides by sfreq)
//   welch (Raw.compute_psd's default; mne psd_array_welch -> scipy.signal.spectrogram): segments of n_fft samples, step
//     n_fft - n_overlap, per segment (x - mean) * hamming_periodic, psd = |rfft|^2 / (sfreq * sum w^2), x2 except DC / Nyquist,
//     mean over segments.
// Both are "windowed segments -> batched R2C -> weighted sum of squared magnitudes", which is how they are implemented:
//   psd_window_kernel (HBM-bound, one read of x) -> cufftExecR2C -> psd_reduce_kernel (HBM-bound, one read of the spectra).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "cufft_api.h"
#include "eegldm.h"
#include "kernels.cuh"

namespace eegldm {

// ------------------------------------------------------------------------------------------------ DPSS (host, double)
// scipy.signal.windows.dpss(M, NW, Kmax, sym, norm=2, return_ratios=True): the Kmax eigenvectors with the largest eigenvalues
// of the symmetric tridiagonal matrix  diag d[t] = ((M-1-2t)/2)^2 cos(2 pi W),  off-diagonal e[t] = t (M-t) / 2,  W = NW / M
// (Percival & Walden 1993), found here by Sturm-sequence bisection + inverse iteration; the concentration ratios by the
// autocorrelation technique (P&W p. 390).  sym = 0 computes the M+1 window and drops its last sample (DFT-even).
namespace {

// number of eigenvalues of the tridiagonal (d, e) that are < x
int sturm_count(const std::vector<double>& d, const std::vector<double>& e2, double x) {
    int cnt = 0;
    double q = 1.0;
    const int n = (int)d.size();
    for (int i = 0; i < n; ++i) {
        q = d[i] - x - (i > 0 ? e2[i - 1] / q : 0.0);
        if (q == 0.0) q = 1e-300;
        if (q < 0.0) ++cnt;
    }
    return cnt;
}

// solve (T - lam I) x = b for tridiagonal T with partial pivoting (LAPACK dgtsv style); b is overwritten with x
void tridiag_solve(const std::vector<double>& d, const std::vector<double>& e, double lam, std::vector<double>& b) {
    const int n = (int)d.size();
    std::vector<double> dl(e), dd(n), du(e), du2(n > 2 ? n - 2 : 0, 0.0);
    for (int i = 0; i < n; ++i) dd[i] = d[i] - lam;
    for (int i = 0; i < n - 1; ++i) {
        if (std::fabs(dd[i]) >= std::fabs(dl[i])) {
            if (dd[i] == 0.0) dd[i] = 1e-300;
            const double f = dl[i] / dd[i];
            dd[i + 1] -= f * du[i];
            b[i + 1] -= f * b[i];
            if (i < n - 2) du2[i] = 0.0;
        } else {   // swap rows i and i+1
            const double f = dd[i] / dl[i];
            dd[i] = dl[i];
            const double t = dd[i + 1];
            dd[i + 1] = du[i] - f * t;
            if (i < n - 2) { du2[i] = du[i + 1]; du[i + 1] = -f * du2[i]; }
            du[i] = t;
            const double tb = b[i];
            b[i] = b[i + 1];
            b[i + 1] = tb - f * b[i + 1];
        }
    }
    if (dd[n - 1] == 0.0) dd[n - 1] = 1e-300;
    b[n - 1] /= dd[n - 1];
    if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / dd[n - 2];
    for (int i = n - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / dd[i];
}

}  // namespace

int dpss_host(int N, double NW, int Kmax, int sym, std::vector<double>& windows, std::vector<double>& ratios, std::string* err) {
    if (N < 2 || Kmax < 1 || Kmax > N || !(NW > 0.0) || NW >= N / 2.0) { if (err) *err = "dpss: need N >= 2, 0 < Kmax <= N, 0 < NW < N/2"; return 1; }
    const int M = sym ? N : N + 1;
    const double W = NW / M;
    std::vector<double> d(M), e(M - 1), e2(M - 1);
    const double cw = std::cos(2.0 * M_PI * W);
    for (int t = 0; t < M; ++t) { const double h = (M - 1 - 2.0 * t) / 2.0; d[t] = h * h * cw; }
    for (int t = 1; t < M; ++t) { e[t - 1] = t * (double)(M - t) / 2.0; e2[t - 1] = e[t - 1] * e[t - 1]; }
    // Gershgorin bounds
    double lo = 1e300, hi = -1e300;
    for (int i = 0; i < M; ++i) {
        const double rad = (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i < M - 1 ? std::fabs(e[i]) : 0.0);
        lo = std::min(lo, d[i] - rad); hi = std::max(hi, d[i] + rad);
    }
    windows.assign((size_t)Kmax * N, 0.0);
    ratios.assign(Kmax, 0.0);
    std::vector<std::vector<double>> vecs;
    for (int k = 0; k < Kmax; ++k) {
        // k-th largest eigenvalue = eigenvalue with index M-1-k (0-based ascending): smallest x with count(x) >= M-k
        double a = lo, b = hi;
        for (int it = 0; it < 200; ++it) {
            const double mid = 0.5 * (a + b);
            if (mid == a || mid == b) break;
            if (sturm_count(d, e2, mid) >= M - k) b = mid; else a = mid;
        }
        const double lam = 0.5 * (a + b);
        // inverse iteration from a deterministic start, re-orthogonalised against the vectors already found
        std::vector<double> v(M);
        for (int i = 0; i < M; ++i) v[i] = 1.0 + 0.37 * std::sin(1.0 + 0.618 * i * (k + 1));
        for (int it = 0; it < 4; ++it) {
            tridiag_solve(d, e, lam, v);
            for (auto& u : vecs) {
                double dot = 0.0;
                for (int i = 0; i < M; ++i) dot += u[i] * v[i];
                for (int i = 0; i < M; ++i) v[i] -= dot * u[i];
            }
            double nrm = 0.0;
            for (int i = 0; i < M; ++i) nrm += v[i] * v[i];
            nrm = std::sqrt(nrm);
            if (!(nrm > 0.0) || !std::isfinite(nrm)) { if (err) *err = "dpss: inverse iteration broke down"; return 1; }
            for (int i = 0; i < M; ++i) v[i] /= nrm;
        }
        // sign convention (Percival & Walden p. 379): even tapers have a positive mean, odd tapers start with a positive lobe
        if (k % 2 == 0) {
            double sum = 0.0;
            for (int i = 0; i < M; ++i) sum += v[i];
            if (sum < 0.0) for (auto& x : v) x = -x;
        } else {
            const double thresh = std::max(1e-7, 1.0 / M);
            for (int i = 0; i < M; ++i)
                if (v[i] * v[i] > thresh) { if (v[i] < 0.0) for (auto& x : v) x = -x; break; }
        }
        // concentration ratio: sum_lag rxx[lag] r[lag],  r[0] = 2W, r[lag] = 4W sinc(2W lag)
        double ratio = 0.0;
        for (int lag = 0; lag < M; ++lag) {
            double rxx = 0.0;
            for (int i = 0; i + lag < M; ++i) rxx += v[i] * v[i + lag];
            const double rr = lag == 0 ? 2.0 * W : 2.0 * std::sin(2.0 * M_PI * W * lag) / (M_PI * lag);
            ratio += rxx * rr;
        }
        ratios[k] = ratio;
        for (int i = 0; i < N; ++i) windows[(size_t)k * N + i] = v[i];   // sym = 0: the last sample of the M = N+1 window is dropped
        vecs.push_back(std::move(v));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ device side
namespace {

// one CTA per (window b, segment s): mean of the segment (remove_dc), then out[b*S+s][n] = (x - mean) * win[s or 0][n]
__global__ void __launch_bounds__(256) psd_window_kernel(const float* __restrict__ x, long long row_stride, int seg_len, int seg_step,
                                                          int S, const float* __restrict__ win, int win_per_seg, int remove_dc,
                                                          float* __restrict__ out) {
    __shared__ float red[8];
    __shared__ float mean_s;
    const int s = blockIdx.x, b = blockIdx.y;
    const float* xs = x + (long long)b * row_stride + (long long)s * seg_step;
    float sum = 0.f;
    if (remove_dc) {
        for (int n = threadIdx.x; n < seg_len; n += blockDim.x) sum += xs[n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
            mean_s = t / (float)seg_len;
        }
        __syncthreads();
    }
    const float mu = remove_dc ? mean_s : 0.f;
    const float* w = win + (win_per_seg ? (size_t)s * seg_len : 0);
    float* o = out + ((size_t)b * S + s) * seg_len;
    for (int n = threadIdx.x; n < seg_len; n += blockDim.x) o[n] = (xs[n] - mu) * w[n];
}

// psd[b][j] = scale_j * sum_s wgt[s] * |X[b*S+s][k0 + j]|^2, scale_j = base (x edge factor at DC / Nyquist); optional 10 log10
__global__ void __launch_bounds__(256) psd_reduce_kernel(const float2* __restrict__ X, int nb, int S, const float* __restrict__ wgt,
                                                          int k0, int nf, float base, float edge_factor, int nyq_bin, int db,
                                                          float* __restrict__ psd, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int j = (int)(i % nf);
    const size_t b = i / nf;
    const int k = k0 + j;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) {
        const float2 v = X[(b * S + s) * (size_t)nb + k];
        acc = fmaf(wgt[s], v.x * v.x + v.y * v.y, acc);
    }
    acc *= base * ((k == 0 || k == nyq_bin) ? edge_factor : 1.f);
    psd[i] = db ? 10.f * log10f(acc) : acc;
}

struct PsdPlan {
    int S = 0, seg_len = 0, seg_step = 0, win_per_seg = 0, k0 = 0, nf = 0, nyq_bin = -1;
    float base = 0.f, edge = 1.f;
    std::vector<float> win_h, wgt_h;
    float *win_d = nullptr, *wgt_d = nullptr;
};
struct PsdScratch { float* seg = nullptr; size_t seg_cap = 0; float2* spec = nullptr; size_t spec_cap = 0; };
struct PsdState {
    std::mutex mu;
    std::map<std::string, PsdPlan> plans;                                    // key: device + cfg + N
    std::map<std::tuple<int, cudaStream_t, int, int>, cufftHandle> fft;      // (device, stream, n_fft, batch)
    std::map<std::pair<int, cudaStream_t>, PsdScratch> scratch;
};
PsdState g_psd;

// frequency mask of rfftfreq(n_fft, 1/sfreq) in [fmin, fmax]: first bin k0 and count nf
void freq_range(int n_fft, float sfreq, float fmin, float fmax, int* k0, int* nf) {
    const int nb = n_fft / 2 + 1;
    int first = -1, cnt = 0;
    for (int k = 0; k < nb; ++k) {
        const double f = (double)k * (double)sfreq / (double)n_fft;   // numpy rfftfreq: k / (n * d)
        if (f >= (double)fmin && f <= (double)fmax) { if (first < 0) first = k; ++cnt; }
    }
    *k0 = first < 0 ? 0 : first;
    *nf = cnt;
}

int psd_geometry(const eegldm_psd_cfg* c, int N, int* n_fft, int* S, int* step, std::string* err) {
    if (!c || N < 2) { *err = "psd: bad argument"; return EEGLDM_ERR_INVALID; }
    if (!(c->sfreq > 0.f)) { *err = "psd: sfreq must be positive"; return EEGLDM_ERR_INVALID; }
    if (c->method == 0) { *n_fft = N; *S = 0; *step = 0; return EEGLDM_OK; }
    if (c->method != 1) { *err = "psd: method must be 0 (multitaper) or 1 (welch)"; return EEGLDM_ERR_INVALID; }
    const int nfft = c->n_fft > 0 ? c->n_fft : 256, nov = c->n_overlap;
    if (nfft > N) { *err = "psd: n_fft is larger than the signal (mne raises)"; return EEGLDM_ERR_SHAPE; }
    if (nov < 0 || nov >= nfft) { *err = "psd: n_overlap must be in [0, n_fft)"; return EEGLDM_ERR_INVALID; }
    *n_fft = nfft; *step = nfft - nov; *S = (N - nov) / *step;
    return EEGLDM_OK;
}

}  // namespace
}  // namespace eegldm

using namespace eegldm;

static int psd_fail(int code, const std::string& m);

extern "C" {

int eegldm_dpss(int N, double half_nbw, int Kmax, int sym, double* windows_out, double* ratios_out) {
    if (!windows_out || !ratios_out) return psd_fail(EEGLDM_ERR_INVALID, "null argument");
    std::vector<double> w, r;
    std::string err;
    if (dpss_host(N, half_nbw, Kmax, sym, w, r, &err)) return psd_fail(EEGLDM_ERR_INVALID, err);
    std::memcpy(windows_out, w.data(), w.size() * sizeof(double));
    std::memcpy(ratios_out, r.data(), r.size() * sizeof(double));
    return EEGLDM_OK;
}

int eegldm_psd_freqs(const eegldm_psd_cfg* cfg, int N, int* n_freqs_out, float* freqs_host) {
    int n_fft, S, step;
    std::string err;
    int r = psd_geometry(cfg, N, &n_fft, &S, &step, &err);
    if (r) return psd_fail(r, err);
    int k0, nf;
    freq_range(n_fft, cfg->sfreq, cfg->fmin, cfg->fmax, &k0, &nf);
    if (n_freqs_out) *n_freqs_out = nf;
    if (freqs_host)
        for (int j = 0; j < nf; ++j) freqs_host[j] = (float)((double)(k0 + j) * (double)cfg->sfreq / (double)n_fft);
    return EEGLDM_OK;
}

int eegldm_psd(const eegldm_psd_cfg* cfg, const float* x_dev, int B, int N, int64_t row_stride, float* psd_dev, void* stream) {
    if (!cfg || (B > 0 && (!x_dev || !psd_dev))) return psd_fail(EEGLDM_ERR_INVALID, "null argument");
    if (B < 0 || row_stride < N) return psd_fail(EEGLDM_ERR_SHAPE, "psd: bad batch / row stride");
    if (B == 0) return EEGLDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int n_fft, S, step;
    std::string err;
    int r = psd_geometry(cfg, N, &n_fft, &S, &step, &err);
    if (r) return psd_fail(r, err);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return psd_fail(EEGLDM_ERR_CUDA, "cudaGetDevice failed");
    std::lock_guard<std::mutex> lock(g_psd.mu);
    CufftApi& api = cufft_api();
    if (!api.load()) return psd_fail(EEGLDM_ERR_CUDA, "cuFFT: " + api.err);
    char keybuf[256];
    std::snprintf(keybuf, sizeof keybuf, "%d|%d|%d|%.9g|%.9g|%.9g|%.9g|%d|%d|%d|%d|%d", dev, cfg->method, N, cfg->sfreq, cfg->fmin, cfg->fmax,
                  cfg->bandwidth, cfg->low_bias, cfg->normalization, cfg->n_fft, cfg->n_overlap, cfg->remove_dc);
    auto it = g_psd.plans.find(keybuf);
    if (it == g_psd.plans.end()) {
        PsdPlan pl;
        pl.seg_len = n_fft;
        freq_range(n_fft, cfg->sfreq, cfg->fmin, cfg->fmax, &pl.k0, &pl.nf);
        pl.nyq_bin = n_fft % 2 == 0 ? n_fft / 2 : -1;
        if (cfg->method == 0) {
            // mne _compute_mt_params: half_nbw = bandwidth * n_times / (2 sfreq), default 4; n_tapers_max = int(2 half_nbw)
            const double half_nbw = cfg->bandwidth > 0.f ? (double)cfg->bandwidth * N / (2.0 * cfg->sfreq) : 4.0;
            const int kmax = (int)(2.0 * half_nbw);
            if (kmax < 1) return psd_fail(EEGLDM_ERR_INVALID, "psd: bandwidth too small (no tapers)");
            std::vector<double> w, lam;
            if (dpss_host(N, half_nbw, kmax, /*sym=*/0, w, lam, &err)) return psd_fail(EEGLDM_ERR_INVALID, err);
            std::vector<int> keep;
            for (int k = 0; k < kmax; ++k) if (!cfg->low_bias || lam[k] > 0.9) keep.push_back(k);
            if (keep.empty()) { int best = 0; for (int k = 1; k < kmax; ++k) if (lam[k] > lam[best]) best = k; keep.push_back(best); }
            pl.S = (int)keep.size(); pl.seg_step = 0; pl.win_per_seg = 1;
            double wsum = 0.0;
            for (int k : keep) wsum += lam[k];
            for (int k : keep) {
                pl.wgt_h.push_back((float)lam[k]);            // weights^2 = eigenvalues
                for (int n = 0; n < N; ++n) pl.win_h.push_back((float)w[(size_t)k * N + n]);
            }
            pl.base = (float)(2.0 / wsum) * (cfg->normalization == 1 ? 1.0f / cfg->sfreq : 1.0f);
            pl.edge = 0.5f;                                   // x_mt[..., 0] /= sqrt(2) (and Nyquist for even n_fft)
        } else {
            pl.S = S; pl.seg_step = step; pl.win_per_seg = 0;
            double w2 = 0.0;
            for (int n = 0; n < n_fft; ++n) {                 // scipy get_window("hamming", n_fft): periodic (fftbins=True)
                const double v = 0.54 - 0.46 * std::cos(2.0 * M_PI * n / n_fft);
                pl.win_h.push_back((float)v); w2 += v * v;
            }
            for (int s = 0; s < S; ++s) pl.wgt_h.push_back(1.0f / S);   // average="mean"
            pl.base = (float)(2.0 / ((double)cfg->sfreq * w2));          // scaling="density", one-sided x2 ...
            pl.edge = 0.5f;                                              // ... except DC and Nyquist
        }
        if (pl.S < 1) return psd_fail(EEGLDM_ERR_SHAPE, "psd: no segments");
        if (cudaMalloc((void**)&pl.win_d, pl.win_h.size() * sizeof(float)) != cudaSuccess ||
            cudaMalloc((void**)&pl.wgt_d, pl.wgt_h.size() * sizeof(float)) != cudaSuccess)
            return psd_fail(EEGLDM_ERR_NOMEM, "psd: cudaMalloc failed");
        cudaMemcpy(pl.win_d, pl.win_h.data(), pl.win_h.size() * sizeof(float), cudaMemcpyHostToDevice);
        cudaMemcpy(pl.wgt_d, pl.wgt_h.data(), pl.wgt_h.size() * sizeof(float), cudaMemcpyHostToDevice);
        it = g_psd.plans.emplace(keybuf, std::move(pl)).first;
    }
    PsdPlan& pl = it->second;
    if (pl.nf == 0) return EEGLDM_OK;
    const int nb = n_fft / 2 + 1;
    // windows per chunk: bound the scratch (segments + spectra, 12 bytes per sample) to ~1 GiB
    const size_t per_win = (size_t)pl.S * ((size_t)n_fft * 4 + (size_t)nb * 8);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, ((size_t)1 << 30) / per_win));
    PsdScratch& sc = g_psd.scratch[std::make_pair(dev, st)];
    const size_t need_seg = (size_t)chunk * pl.S * n_fft, need_spec = (size_t)chunk * pl.S * nb;
    if (need_seg > sc.seg_cap) {
        if (sc.seg) cudaFree(sc.seg);
        if (cudaMalloc((void**)&sc.seg, need_seg * sizeof(float)) != cudaSuccess) { sc.seg = nullptr; sc.seg_cap = 0; return psd_fail(EEGLDM_ERR_NOMEM, "psd scratch"); }
        sc.seg_cap = need_seg;
    }
    if (need_spec > sc.spec_cap) {
        if (sc.spec) cudaFree(sc.spec);
        if (cudaMalloc((void**)&sc.spec, need_spec * sizeof(float2)) != cudaSuccess) { sc.spec = nullptr; sc.spec_cap = 0; return psd_fail(EEGLDM_ERR_NOMEM, "psd scratch"); }
        sc.spec_cap = need_spec;
    }
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nbatch = std::min(chunk, B - b0);
        auto fkey = std::make_tuple(dev, st, n_fft, nbatch * pl.S);
        auto fit = g_psd.fft.find(fkey);
        if (fit == g_psd.fft.end()) {
            cufftHandle h;
            int n[1] = {n_fft};
            if (api.PlanMany(&h, 1, n, nullptr, 1, n_fft, nullptr, 1, nb, CUFFT_R2C, nbatch * pl.S) != CUFFT_SUCCESS)
                return psd_fail(EEGLDM_ERR_CUDA, "cufftPlanMany failed");
            fit = g_psd.fft.emplace(fkey, h).first;
        }
        dim3 grid(pl.S, nbatch);
        psd_window_kernel<<<grid, 256, 0, st>>>(x_dev + (size_t)b0 * row_stride, (long long)row_stride, n_fft, pl.seg_step, pl.S, pl.win_d,
                                                 pl.win_per_seg, cfg->remove_dc, sc.seg);
        g_launch_count += 1;
        if (api.SetStream(fit->second, st) != CUFFT_SUCCESS ||
            api.ExecR2C(fit->second, sc.seg, reinterpret_cast<cufftComplex*>(sc.spec)) != CUFFT_SUCCESS)
            return psd_fail(EEGLDM_ERR_CUDA, "cufftExecR2C failed");
        const size_t total = (size_t)nbatch * pl.nf;
        psd_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sc.spec, nb, pl.S, pl.wgt_d, pl.k0, pl.nf, pl.base, pl.edge, pl.nyq_bin,
                                                                            cfg->db, psd_dev + (size_t)b0 * pl.nf, total);
        g_launch_count += 1;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return psd_fail(EEGLDM_ERR_CUDA, std::string("psd kernels: ") + cudaGetErrorString(e));
    return EEGLDM_OK;
}

// cropped = sample[:, :, crop_left : L - crop_right] straight into host memory (one strided copy; synchronises the stream)
int eegldm_crop_to_host(const float* x_dev, int64_t rows, int L, int crop_left, int crop_right, float* out_host, void* stream) {
    if (rows < 0 || L <= 0 || crop_left < 0 || crop_right < 0 || crop_left + crop_right >= L) return psd_fail(EEGLDM_ERR_SHAPE, "crop: bad shape");
    if (rows == 0) return EEGLDM_OK;
    if (!x_dev || !out_host) return psd_fail(EEGLDM_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t w = (size_t)(L - crop_left - crop_right) * sizeof(float);
    cudaError_t e = cudaMemcpy2DAsync(out_host, w, x_dev + crop_left, (size_t)L * sizeof(float), w, (size_t)rows, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return psd_fail(EEGLDM_ERR_CUDA, std::string("crop copy: ") + cudaGetErrorString(e));
    return EEGLDM_OK;
}

// numpy .npy, format version 1.0, little-endian fp32, C order
int eegldm_write_npy_f32(const char* path, const float* data_host, const int64_t* shape, int ndim) {
    if (!path || !shape || ndim < 0 || ndim > 8) return psd_fail(EEGLDM_ERR_INVALID, "npy: bad argument");
    size_t n = 1;
    std::string shp = "(";
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] < 0) return psd_fail(EEGLDM_ERR_SHAPE, "npy: negative dimension");
        n *= (size_t)shape[i];
        shp += std::to_string((long long)shape[i]);
        if (ndim == 1 || i + 1 < ndim) shp += ",";
        if (i + 1 < ndim) shp += " ";
    }
    shp += ")";
    if (n && !data_host) return psd_fail(EEGLDM_ERR_INVALID, "npy: null data");
    std::string hdr = "{'descr': '<f4', 'fortran_order': False, 'shape': " + shp + ", }";
    const size_t unpadded = 10 + hdr.size() + 1;                 // magic(6) + version(2) + len(2) + dict + newline
    hdr.append((64 - unpadded % 64) % 64, ' ');
    hdr.push_back('\n');
    if (hdr.size() > 65535) return psd_fail(EEGLDM_ERR_INVALID, "npy: header too long");
    FILE* f = std::fopen(path, "wb");
    if (!f) return psd_fail(EEGLDM_ERR_INVALID, std::string("npy: cannot open ") + path);
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    const unsigned char len[2] = {(unsigned char)(hdr.size() & 0xFF), (unsigned char)(hdr.size() >> 8)};
    bool ok = std::fwrite(magic, 1, 8, f) == 8 && std::fwrite(len, 1, 2, f) == 2 && std::fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    if (ok && n) ok = std::fwrite(data_host, sizeof(float), n, f) == n;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) return psd_fail(EEGLDM_ERR_INVALID, std::string("npy: write failed for ") + path);
    return EEGLDM_OK;
}

// sample_{first_index + i}.npy, each of shape [1, C, L] (what sample_trials.py:170 writes per seed), from a host batch [B][C][L]
int eegldm_save_windows_npy(const char* dir, const char* prefix, int64_t first_index, const float* data_host, int B, int C, int L) {
    if (!dir || !prefix || (B > 0 && !data_host) || B < 0 || C < 1 || L < 1) return psd_fail(EEGLDM_ERR_INVALID, "save_windows: bad argument");
    const int64_t shape[3] = {1, C, L};
    for (int i = 0; i < B; ++i) {
        const std::string path = std::string(dir) + "/" + prefix + std::to_string((long long)(first_index + i)) + ".npy";
        int r = eegldm_write_npy_f32(path.c_str(), data_host + (size_t)i * C * L, shape, 3);
        if (r) return r;
    }
    return EEGLDM_OK;
}

}  // extern "C"

static int psd_fail(int code, const std::string& m) {
    eegldm::set_last_error(m);   // eegldm_last_error() (engine.cu)
    return code;
}
