// PatchDiscriminator (monai-generative, as built at src/train_autoencoderkl.py:135-137 from config/config_aekl_eeg.yaml:30-40)
// and the adversarial half of the autoencoder training step (train_autoencoderkl.py:213-234):
//   generator term      adv_w * MSE(lrelu_0.05(D(recon)), 1)  -> gradient w.r.t. recon through D (BatchNorm in training mode)
//   discriminator step  loss_d = adv_w * 0.5 * (MSE(act(D(recon.detach())), 0) + MSE(act(D(x)), 1));  backward;  Adam(lr_d)
// Topology: initial_conv (in -> C, k3 s2 p1, bias, LeakyReLU 0.2) | layers l = 0..n-1 (C 2^l -> C 2^(l+1), s2 except the last,
// no bias, BatchNorm1d, LeakyReLU 0.2) | final_conv (-> out, k3 s1, bias).
//
// Data flow (channels-last [B][T][C] like the rest of the engine): a block's norm + activation is never materialised in the
// forward pass -- it is the prologue of the next conv (per-channel scale / shift + LeakyReLU), exactly as GroupNorm + SiLU are
// for the UNet.  Backward: the data gradient of a conv is the forward conv kernel on transformed weights (stride 2: the input's
// row pairs seen as one row of 2 Cin channels, launch_dgrad_weights), BatchNorm + LeakyReLU backward is two HBM-bound passes.
// The fake pass is run ONCE and used for both the generator term and the discriminator step (the reference's second forward on
// recon.detach() sees the same weights and the same input; its extra running-statistics update is applied).
#include <cmath>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "disc.h"
#include "eegldm.h"
#include "kernels.cuh"

using namespace eegldm;

namespace {
constexpr float BN_EPS = 1e-5f, BN_MOMENTUM = 0.1f, LEAKY = 0.2f;

int dfail(int code, const std::string& m) { set_last_error(m); return code; }
int dcuda(cudaError_t e, const char* what) { set_last_error(std::string(what) + ": " + cudaGetErrorString(e)); return EEGLDM_ERR_CUDA; }
#define DCU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return dcuda(e__, #expr); } while (0)

struct DParam { std::string name; std::vector<int64_t> shape; std::vector<float> data; bool loaded = false, buffer = false;
                size_t numel() const { size_t n = 1; for (auto s : shape) n *= (size_t)s; return n; } };
struct DLayer { std::string prefix; int cin, cout, stride, pad; bool has_bias, has_norm, has_act;
                size_t ow = 0, ob = 0, og = 0, obe = 0, orm = 0, orv = 0;     // offsets into P (parameters) / R (running statistics)
                // tensor-pipe form (f16x3): the conv runs as a stride-1 conv over cin_v = stride * cin channels (stride 2: row pairs)
                bool tc = false; int cin_v = 0, cod = 0; size_t o_imgf = 0, o_imgd = 0; };   // offsets (u16) into the image pool
}  // namespace

struct DiscPass {   // saved tensors of one forward pass (pointers into the workspace)
    const float* x = nullptr;
    std::vector<float*> h;        // conv outputs (pre-norm), one per block
    std::vector<float*> ss;       // [2][B][C] scale / shift of a block's norm (null: no norm)
    std::vector<float*> mr;       // [2][C] mean / rstd
    std::vector<int> T;           // output length per block
    std::vector<int> rp;          // copies of scale / shift per sample (2: the consumer reads row pairs)
    int B = 0, L = 0;
    bool valid = false;
};

struct eegldm_disc {
    eegldm_disc_cfg cfg{};
    std::vector<DParam> params;
    std::unordered_map<std::string, int> index;
    std::vector<DLayer> layers;
    bool finalized = false;
    size_t nP = 0, nR = 0;
    float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr, *R = nullptr;   // parameters (SIMT conv images), grads, Adam moments, running stats
    float* Wd = nullptr; size_t Wd_cap = 0;        // data-gradient weights of the layer being processed
    int math = EEGLDM_MATH_F16X3_TC;                // tensor-pipe convs where the shapes allow (eegldm_disc_set_math)
    uint16_t* img = nullptr; size_t img_n = 0;      // tcgen05 weight images (forward + data-gradient) of the tensor-pipe layers
    float* wv = nullptr; size_t wv_n = 0;           // scratch: virtual (row-pair / transposed) weights in the SIMT image, and their gradient
    bool images_dirty = true;
    double* sums = nullptr;                         // [2 * max C]
    float* ws = nullptr; size_t ws_cap = 0, ws_off = 0;   // workspace (bump allocator, reset per training step / forward)
    float* losses = nullptr;                        // device [4]: generator term, d_fake, d_real, spare
    int step = 0;
    long long batches_tracked = 0;
    DiscPass fake, real;
    size_t pass_end[2] = {0, 0}, slot_floats = 0;   // autograd boundary (eegldm_disc_forward_train): two workspace halves, one per recorded pass
    ~eegldm_disc() { for (void* p : {(void*)P, (void*)G, (void*)M, (void*)V, (void*)R, (void*)Wd, (void*)sums, (void*)ws, (void*)losses, (void*)img, (void*)wv}) if (p) cudaFree(p); }
    float* alloc(size_t n) { n = (n + 63) & ~size_t(63); float* p = ws ? ws + ws_off : nullptr; ws_off += n; return p; }
};

namespace {

int build_disc(eegldm_disc* d) {
    const auto& c = d->cfg;
    if (c.in_channels < 1 || c.out_channels < 1 || c.num_channels < 1 || c.num_layers_d < 1 || c.num_layers_d > 8)
        return dfail(EEGLDM_ERR_INVALID, "bad discriminator config");
    if (c.kernel_size != 3) return dfail(EEGLDM_ERR_INVALID, "PatchDiscriminator: kernel_size must be 3 (config_aekl_eeg.yaml:37)");
    if (c.padding != 1) return dfail(EEGLDM_ERR_INVALID, "PatchDiscriminator: padding must be 1 (config_aekl_eeg.yaml:40)");
    auto add = [&](const std::string& n, std::vector<int64_t> shape, bool buffer) {
        d->index[n] = (int)d->params.size();
        DParam p; p.name = n; p.shape = std::move(shape); p.buffer = buffer;
        d->params.push_back(std::move(p));
    };
    auto layer = [&](const std::string& prefix, int cin, int cout, int stride, int pad, bool bias, bool norm, bool act) {
        DLayer l{prefix, cin, cout, stride, pad, bias, norm, act};
        add(prefix + ".conv.weight", {cout, cin, 3}, false);
        if (bias) add(prefix + ".conv.bias", {cout}, false);
        if (norm) {
            add(prefix + ".adn.N.weight", {cout}, false); add(prefix + ".adn.N.bias", {cout}, false);
            add(prefix + ".adn.N.running_mean", {cout}, true); add(prefix + ".adn.N.running_var", {cout}, true);
            add(prefix + ".adn.N.num_batches_tracked", {}, true);
        }
        d->layers.push_back(l);
    };
    layer("initial_conv", c.in_channels, c.num_channels, 2, c.padding, true, false, true);
    int cin = c.num_channels, cout = 2 * c.num_channels;
    for (int l = 0; l < c.num_layers_d; ++l) {
        layer(std::to_string(l), cin, cout, l == c.num_layers_d - 1 ? 1 : 2, c.padding, false, true, true);
        cin = cout; cout *= 2;
    }
    layer("final_conv", cin, c.out_channels, 1, (c.kernel_size - 1) / 2, true, false, false);
    return EEGLDM_OK;
}

const std::vector<float>& pget(const eegldm_disc* d, const std::string& n) { return d->params[d->index.at(n)].data; }

int finalize_disc(eegldm_disc* d) {
    for (auto& p : d->params) if (!p.loaded) return dfail(EEGLDM_ERR_MISSING, "missing state_dict key: " + p.name);
    std::vector<float> flat, run;
    auto push = [](std::vector<float>& dst, const float* src, size_t n) {
        const size_t off = (dst.size() + 63) & ~size_t(63);
        dst.resize(off + n);
        std::memcpy(dst.data() + off, src, n * sizeof(float));
        return off;
    };
    for (auto& l : d->layers) {
        const auto& w = pget(d, l.prefix + ".conv.weight");
        std::vector<float> pk((size_t)l.cout * l.cin * 3);
        for (int co = 0; co < l.cout; ++co)
            for (int ci = 0; ci < l.cin; ++ci)
                for (int k = 0; k < 3; ++k) pk[((size_t)ci * 3 + k) * l.cout + co] = w[((size_t)co * l.cin + ci) * 3 + k];
        l.ow = push(flat, pk.data(), pk.size());
        if (l.has_bias) { const auto& b = pget(d, l.prefix + ".conv.bias"); l.ob = push(flat, b.data(), b.size()); }
        if (l.has_norm) {
            const auto& g = pget(d, l.prefix + ".adn.N.weight"); l.og = push(flat, g.data(), g.size());
            const auto& b = pget(d, l.prefix + ".adn.N.bias"); l.obe = push(flat, b.data(), b.size());
            const auto& rm = pget(d, l.prefix + ".adn.N.running_mean"); l.orm = push(run, rm.data(), rm.size());
            const auto& rv = pget(d, l.prefix + ".adn.N.running_var"); l.orv = push(run, rv.data(), rv.size());
            d->batches_tracked = (long long)pget(d, l.prefix + ".adn.N.num_batches_tracked")[0];
        }
    }
    d->nP = (flat.size() + 63) & ~size_t(63);
    flat.resize(d->nP, 0.f);
    d->nR = std::max<size_t>((run.size() + 63) & ~size_t(63), 64);
    run.resize(d->nR, 0.f);
    for (float** p : {&d->P, &d->G, &d->M, &d->V, &d->R}) if (*p) { cudaFree(*p); *p = nullptr; }
    DCU(cudaMalloc((void**)&d->P, d->nP * sizeof(float)));
    DCU(cudaMalloc((void**)&d->G, d->nP * sizeof(float)));
    DCU(cudaMalloc((void**)&d->M, d->nP * sizeof(float)));
    DCU(cudaMalloc((void**)&d->V, d->nP * sizeof(float)));
    DCU(cudaMalloc((void**)&d->R, d->nR * sizeof(float)));
    DCU(cudaMemcpy(d->P, flat.data(), d->nP * sizeof(float), cudaMemcpyHostToDevice));
    DCU(cudaMemcpy(d->R, run.data(), d->nR * sizeof(float), cudaMemcpyHostToDevice));
    DCU(cudaMemset(d->G, 0, d->nP * sizeof(float)));
    DCU(cudaMemset(d->M, 0, d->nP * sizeof(float)));
    DCU(cudaMemset(d->V, 0, d->nP * sizeof(float)));
    int maxc = 1;
    for (auto& l : d->layers) maxc = std::max(maxc, l.cout);
    if (!d->sums) DCU(cudaMalloc((void**)&d->sums, (size_t)2 * 4096 * sizeof(double)));
    if (maxc > 4096) return dfail(EEGLDM_ERR_INVALID, "discriminator wider than 4096 channels");
    if (!d->losses) DCU(cudaMalloc((void**)&d->losses, 4 * sizeof(float)));
    // tensor-pipe layers: stride-1 view with cin_v = stride * cin input channels; forward, data-gradient and weight-gradient
    // kernels all need channel counts in multiples of 128 (the length conditions are checked per call)
    size_t img_n = 0, wv_n = 64;
    for (auto& l : d->layers) {
        l.cin_v = l.stride * l.cin; l.cod = l.cin_v;
        l.tc = !l.has_bias && l.cin_v % 128 == 0 && l.cout % 128 == 0;
        if (!l.tc) continue;
        l.o_imgf = img_n; img_n += conv_tc_image_u16(l.cin_v, l.cout, 3);
        l.o_imgd = img_n; img_n += conv_tc_image_u16(l.cout, l.cod, 3);
        wv_n = std::max(wv_n, (size_t)l.cin_v * 3 * l.cout);
    }
    if (d->img) { cudaFree(d->img); d->img = nullptr; }
    if (d->wv) { cudaFree(d->wv); d->wv = nullptr; }
    d->img_n = img_n; d->wv_n = wv_n;
    if (img_n) DCU(cudaMalloc((void**)&d->img, img_n * sizeof(uint16_t)));
    DCU(cudaMalloc((void**)&d->wv, wv_n * sizeof(float)));
    d->images_dirty = true;
    d->step = 0;
    d->fake.valid = d->real.valid = false;
    d->finalized = true;
    return EEGLDM_OK;
}

int out_len(int Tin, const DLayer& l) { return (Tin + 2 * l.pad - 3) / l.stride + 1; }

// does layer l run on the tensor pipe for this input length?  (stride 2: the row-pair view needs an even length)
bool use_tc(const eegldm_disc* d, const DLayer& l, int Tin) {
    if (d->math != EEGLDM_MATH_F16X3_TC || !l.tc) return false;
    if (l.stride == 2 && (Tin & 1)) return false;
    const int Tout = out_len(Tin, l);
    return Tout == Tin / l.stride && Tout % WG_CHUNK == 0;
}

// rebuild the tcgen05 weight images from the current parameters (after finalize and after every optimiser step)
int ensure_images(eegldm_disc* d, cudaStream_t st) {
    if (!d->images_dirty) return EEGLDM_OK;
    for (auto& l : d->layers) {
        if (!l.tc) continue;
        const float* wf = d->P + l.ow;
        if (l.stride == 2) { DCU(launch_s2z_weights(d->P + l.ow, l.cin, l.cout, d->wv, st)); wf = d->wv; }
        DCU(launch_pack_conv_tc_dev(wf, l.cin_v, l.cout, 3, d->img + l.o_imgf, st));
        DCU(launch_dgrad_weights(d->P + l.ow, l.cin, l.cout, l.stride, d->wv, st));
        DCU(launch_pack_conv_tc_dev(d->wv, l.cout, l.cod, 3, d->img + l.o_imgd, st));
    }
    d->images_dirty = false;
    return EEGLDM_OK;
}

// one tcgen05 conv launch (fused producer): out[B][T][Cout] = conv3_same(act(scale * x + shift)), x [B][T][Cin]
cudaError_t tc_conv(const float* x, const float* scale, const float* shift, int act, const uint16_t* img, int B, int T, int Cin, int Cout,
                    float* out, cudaStream_t st) {
    TcConvParams q{};
    q.nseg = 1; q.Cout = Cout; q.Tout = T; q.nsegs16 = (int)((long long)B * T / 16);
    q.bn = conv_tc_bn(Cout, Cin / TC_BK * 3);
    q.direct = 1;
    q.seg[0] = TcSeg{nullptr, reinterpret_cast<const uint8_t*>(img), 3, Cin / TC_BK, x, nullptr, Cin, 0, scale, shift, act, RS_NONE, T};
    q.out = out;
    return launch_conv_tc(q, true, st);
}

// workspace floats of one forward pass (+ backward temporaries when training)
size_t pass_floats(const eegldm_disc* d, int B, int L, bool backward) {
    size_t n = 0, maxact = 0;
    int T = L;
    for (auto& l : d->layers) {
        const size_t in_n = (size_t)B * T * l.cin;
        T = out_len(T, l);
        const size_t out_n = (size_t)B * T * l.cout;
        n += ((out_n + 63) & ~size_t(63)) + ((size_t)2 * B * l.cout + 64) + ((size_t)2 * l.cout + 64);
        maxact = std::max(maxact, std::max(in_n, out_n));
    }
    if (backward) {
        n += 3 * ((maxact + 63) & ~size_t(63)) + 256;   // dh, da, materialised conv input
        size_t img = 0;                                // fp16 hi/lo operand images of the weight-gradient GEMM
        int Tl = L;
        for (auto& l : d->layers) {
            const int To = out_len(Tl, l);
            if (l.tc) img = std::max(img, wgrad_image_bytes((size_t)B * To, l.cin_v, 1) + wgrad_image_bytes((size_t)B * To, l.cout, 0));
            Tl = To;
        }
        n += img / 4 + 256;
    }
    return n + 1024;
}

int ensure_ws(eegldm_disc* d, size_t floats) {
    if (floats <= d->ws_cap) return EEGLDM_OK;
    if (d->ws) { cudaFree(d->ws); d->ws = nullptr; d->ws_cap = 0; }
    DCU(cudaMalloc((void**)&d->ws, floats * sizeof(float)));
    d->ws_cap = floats;
    return EEGLDM_OK;
}

// D(x): x [B][L][in] channels-last.  training: batch statistics (running statistics updated n_updates times), else running ones.
int disc_forward(eegldm_disc* d, const float* x, int B, int L, bool training, int n_updates, DiscPass& ps, cudaStream_t st) {
    ps.x = x; ps.B = B; ps.L = L; ps.h.clear(); ps.ss.clear(); ps.mr.clear(); ps.T.clear(); ps.rp.clear();
    const float* in = x;
    const float *scale = nullptr, *shift = nullptr;
    int act = 0, T = L;
    int r0 = ensure_images(d, st);
    if (r0) return r0;
    for (size_t li = 0; li < d->layers.size(); ++li) {
        auto& l = d->layers[li];
        const int Tout = out_len(T, l);
        if (Tout < 1) return dfail(EEGLDM_ERR_SHAPE, "signal too short for the discriminator");
        float* h = d->alloc((size_t)B * Tout * l.cout);
        if (use_tc(d, l, T)) {   // (stride 2: rows 2t, 2t+1 are one row of 2 cin channels; scale / shift were written twice per sample)
            DCU(tc_conv(in, scale, shift, act, d->img + l.o_imgf, B, Tout, l.cin_v, l.cout, h, st));
        } else {
            ConvParams p{};
            p.seg[0] = ConvSeg{in, nullptr, l.cin, 0, scale, shift, act, RS_NONE, T, d->P + l.ow, 3};
            p.nseg = 1; p.Cout = l.cout; p.Tout = Tout; p.Tc = T; p.stride = l.stride; p.pad_left = l.pad;
            p.bias = l.has_bias ? d->P + l.ob : nullptr; p.out = h; p.B = B;
            DCU(launch_conv_simt(p, st));
        }
        float *ss = nullptr, *mr = nullptr;
        // a tensor-pipe stride-2 consumer reads this block's output as rows of 2 cout channels: its prologue then needs the
        // per-channel scale / shift twice per sample ([B][2][C] is [2B][C])
        const int rp = (li + 1 < d->layers.size() && d->layers[li + 1].stride == 2 && use_tc(d, d->layers[li + 1], Tout)) ? 2 : 1;
        const int Brep = B * rp;
        if (l.has_norm) {
            ss = d->alloc((size_t)2 * Brep * l.cout);
            mr = d->alloc((size_t)2 * l.cout);
            if (training) {
                DCU(launch_bn_stats(h, (size_t)B * Tout, l.cout, d->P + l.og, d->P + l.obe, BN_EPS, Brep, d->sums, ss, ss + (size_t)Brep * l.cout, mr,
                                    mr + l.cout, d->R + l.orm, d->R + l.orv, BN_MOMENTUM, n_updates, st));
            } else {   // eval(): normalise with the running statistics -- a per-channel affine, computed on the host side of the stream
                std::vector<float> run(2 * (size_t)l.cout), par(2 * (size_t)l.cout), hs((size_t)2 * Brep * l.cout);
                DCU(cudaMemcpyAsync(run.data(), d->R + l.orm, l.cout * sizeof(float), cudaMemcpyDeviceToHost, st));
                DCU(cudaMemcpyAsync(run.data() + l.cout, d->R + l.orv, l.cout * sizeof(float), cudaMemcpyDeviceToHost, st));
                DCU(cudaMemcpyAsync(par.data(), d->P + l.og, l.cout * sizeof(float), cudaMemcpyDeviceToHost, st));
                DCU(cudaMemcpyAsync(par.data() + l.cout, d->P + l.obe, l.cout * sizeof(float), cudaMemcpyDeviceToHost, st));
                DCU(cudaStreamSynchronize(st));
                for (int c = 0; c < l.cout; ++c) {
                    const float sc = par[c] / std::sqrt(run[l.cout + c] + BN_EPS), sh = par[l.cout + c] - run[c] * sc;
                    for (int b = 0; b < Brep; ++b) { hs[(size_t)b * l.cout + c] = sc; hs[((size_t)Brep + b) * l.cout + c] = sh; }
                }
                DCU(cudaMemcpyAsync(ss, hs.data(), hs.size() * sizeof(float), cudaMemcpyHostToDevice, st));
                DCU(cudaStreamSynchronize(st));
            }
        }
        ps.h.push_back(h); ps.ss.push_back(ss); ps.mr.push_back(mr); ps.T.push_back(Tout); ps.rp.push_back(rp);
        in = h; T = Tout;
        scale = ss; shift = ss ? ss + (size_t)Brep * l.cout : nullptr;
        act = l.has_act ? 2 : 0;
    }
    ps.valid = true;
    return EEGLDM_OK;
}

// Backward sweep through a saved pass.  dlogits: gradient w.r.t. the final conv's output (overwritten freely).  wgrad: accumulate the
// parameter gradients into G.  dx (nullable): receives (accumulate: +=) the gradient w.r.t. the input signal.
int disc_backward(eegldm_disc* d, const DiscPass& ps, float* dlogits, bool wgrad, float* dx, int dx_accumulate, cudaStream_t st) {
    const int B = ps.B, nl = (int)d->layers.size();
    size_t maxact = 0;
    { int T = ps.L; for (int i = 0; i < nl; ++i) { maxact = std::max(maxact, (size_t)B * T * d->layers[i].cin); T = ps.T[i]; maxact = std::max(maxact, (size_t)B * T * d->layers[i].cout); } }
    float* buf_a = d->alloc(maxact);    // gradient w.r.t. a conv's input (post-activation)
    float* buf_h = d->alloc(maxact);    // gradient w.r.t. a conv's output (pre-norm)
    float* buf_in = wgrad ? d->alloc(maxact) : nullptr;   // materialised conv input for the weight gradient
    float* dh = dlogits;
    for (int i = nl - 1; i >= 0; --i) {
        const DLayer& l = d->layers[i];
        const int Tin = i ? ps.T[i - 1] : ps.L, Tout = ps.T[i];
        const float* in_raw = i ? ps.h[i - 1] : ps.x;            // the conv's input before the previous block's norm + activation
        const DLayer* prev = i ? &d->layers[i - 1] : nullptr;
        const bool tc = use_tc(d, l, Tin);
        const size_t ss_half = prev && prev->has_norm ? (size_t)B * ps.rp[i - 1] * prev->cout : 0;   // shift follows scale after this many floats
        if (wgrad && tc) {
            // weight gradient on the tensor pipe: operands split to fp16 hi/lo once (the conv input with its norm + activation applied,
            // rows of cin_v channels with a halo; dy plain), then one split-K GEMM per tap
            const size_t mark = d->ws_off;
            uint8_t* a_img = reinterpret_cast<uint8_t*>(d->alloc(wgrad_image_bytes((size_t)B * Tout, l.cin_v, 1) / 4 + 64));
            uint8_t* y_img = reinterpret_cast<uint8_t*>(d->alloc(wgrad_image_bytes((size_t)B * Tout, l.cout, 0) / 4 + 64));
            const float* sc = prev && prev->has_norm ? ps.ss[i - 1] : nullptr;   // [rp][C] at the front of the buffer = one row of cin_v channels
            DCU(launch_wgrad_split(in_raw, sc, sc ? sc + ss_half : nullptr, 0, prev && prev->has_act ? 2 : 0, B, Tout, l.cin_v, 1, a_img, st));
            DCU(launch_wgrad_split(dh, nullptr, nullptr, 0, 0, B, Tout, l.cout, 0, y_img, st));
            if (l.stride == 1) DCU(launch_wgrad_tc(a_img, y_img, d->G + l.ow, B, Tout, l.cin_v, l.cout, 3, 0, 3, st));
            else {   // row-pair view: taps (t-1, t) of the virtual weights, folded back onto the three real taps
                DCU(cudaMemsetAsync(d->wv, 0, (size_t)l.cin_v * 3 * l.cout * sizeof(float), st));
                DCU(launch_wgrad_tc(a_img, y_img, d->wv, B, Tout, l.cin_v, l.cout, 3, 0, 2, st));
                DCU(launch_s2z_grad_fold(d->wv, l.cin, l.cout, d->G + l.ow, st));
            }
            d->ws_off = mark;
        } else if (wgrad) {
            const float* a = in_raw;
            if (prev && (prev->has_norm || prev->has_act)) {
                DCU(launch_affine_lrelu(in_raw, prev->has_norm ? ps.ss[i - 1] : nullptr, prev->has_norm ? ps.ss[i - 1] + ss_half : nullptr,
                                        prev->cout, prev->has_act ? LEAKY : 1.f, buf_in, (size_t)B * Tin * l.cin, st));
                a = buf_in;
            }
            ConvGradParams g{};
            g.dy = dh; g.a = a; g.w = d->P + l.ow; g.dw = d->G + l.ow; g.db = l.has_bias ? d->G + l.ob : nullptr;
            g.Cin = l.cin; g.Cout = l.cout; g.taps = 3; g.stride = l.stride; g.pad = l.pad; g.ups = 0; g.Tin = Tin; g.Tc = Tin; g.Tout = Tout; g.B = B;
            DCU(launch_conv_bwd_weight(g, st));
        }
        const bool need_da = i > 0 || dx != nullptr;
        if (!need_da) break;
        // data gradient through the forward conv kernel on transformed weights
        const bool z2 = l.stride == 2;
        if (z2 && (Tin & 1)) return dfail(EEGLDM_ERR_SHAPE, "discriminator backward: odd length before a stride-2 conv");
        const int cod = z2 ? 2 * l.cin : l.cin;
        const size_t nwd = (size_t)l.cout * 3 * cod;
        if (nwd > d->Wd_cap) { if (d->Wd) cudaFree(d->Wd); d->Wd = nullptr; d->Wd_cap = 0; DCU(cudaMalloc((void**)&d->Wd, nwd * sizeof(float))); d->Wd_cap = nwd; }
        float* da = (i == 0 && !dx_accumulate) ? dx : buf_a;
        if (!z2 && Tout != Tin) return dfail(EEGLDM_ERR_SHAPE, "discriminator backward: stride-1 conv must keep the length");
        if (tc) DCU(tc_conv(dh, nullptr, nullptr, 0, d->img + l.o_imgd, B, Tout, l.cout, cod, da, st));
        else {
            DCU(launch_dgrad_weights(d->P + l.ow, l.cin, l.cout, l.stride, d->Wd, st));
            ConvParams p{};
            p.seg[0] = ConvSeg{dh, nullptr, l.cout, 0, nullptr, nullptr, 0, RS_NONE, Tout, d->Wd, 3};
            p.nseg = 1; p.Cout = cod; p.Tout = Tout; p.Tc = Tout; p.stride = 1; p.pad_left = 1; p.out = da; p.B = B;
            DCU(launch_conv_simt(p, st));
        }
        if (i == 0) {
            if (dx_accumulate) DCU(launch_axpy(buf_a, dx, 1.f, 1, (size_t)B * Tin * l.cin, st));
            break;
        }
        // through the previous block's activation (+ norm): da -> gradient w.r.t. that block's conv output
        const size_t n = (size_t)B * Tin * prev->cout;
        if (prev->has_norm)
            DCU(launch_bn_lrelu_bwd(da, in_raw, ps.mr[i - 1], ps.mr[i - 1] + prev->cout, d->P + prev->og, d->P + prev->obe, (size_t)B * Tin, prev->cout,
                                    prev->has_act ? LEAKY : 1.f, d->sums, buf_h, wgrad ? d->G + prev->og : nullptr, wgrad ? d->G + prev->obe : nullptr, st));
        else if (prev->has_act) DCU(launch_lrelu_bwd(da, in_raw, LEAKY, buf_h, n, st));
        else DCU(cudaMemcpyAsync(buf_h, da, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
        dh = buf_h;
    }
    return EEGLDM_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ hooks used by the autoencoder step
namespace eegldm {

int disc_prepare_step(eegldm_disc* d, int B, int L, cudaStream_t st) {
    if (!d || !d->finalized) return dfail(EEGLDM_ERR_MISSING, "eegldm_disc_finalize has not been called");
    if (d->cfg.in_channels != 1) return dfail(EEGLDM_ERR_INVALID, "adversarial step: the discriminator must take the 1-channel signal");
    const size_t need = 2 * pass_floats(d, B, L, false) + pass_floats(d, B, L, true) + (size_t)4 * B * (L / 8 + 8);
    int r = ensure_ws(d, need);
    if (r) return r;
    d->ws_off = 0;
    d->fake.valid = d->real.valid = false;
    DCU(cudaMemsetAsync(d->losses, 0, 4 * sizeof(float), st));
    return EEGLDM_OK;
}

// generator term: loss_dev[0] += MSE(act(D(recon)), 1) (unweighted, as the reference logs it); drecon += adv_weight * d(that)/d recon
int disc_generator_term(eegldm_disc* d, const float* recon, int B, int L, float adv_weight, int no_act, float* drecon, cudaStream_t st) {
    int r = disc_forward(d, recon, B, L, true, /*n_updates=*/2, d->fake, st);
    if (r) return r;
    const size_t n = (size_t)B * d->fake.T.back() * d->cfg.out_channels;
    float* dl = d->alloc(n);
    DCU(launch_adv_loss(d->fake.h.back(), n, 1.f, no_act ? 1.f : 0.05f, 1.f, d->losses + 0, adv_weight, dl, st));
    const size_t mark = d->ws_off;
    r = disc_backward(d, d->fake, dl, /*wgrad=*/false, drecon, /*accumulate=*/1, st);
    d->ws_off = mark;   // the sweep's temporaries are free again
    return r;
}

// discriminator step on the saved fake pass and a fresh real pass; losses[1] = d_fake, losses[2] = d_real (unweighted)
int disc_step(eegldm_disc* d, const float* x_real, int B, int L, float adv_weight, int no_act, float lr, float b1, float b2, float eps,
              cudaStream_t st) {
    if (!d->fake.valid) return dfail(EEGLDM_ERR_INVALID, "discriminator step without a generator term");
    DCU(cudaMemsetAsync(d->G, 0, d->nP * sizeof(float), st));
    const float slope = no_act ? 1.f : 0.05f;
    const size_t n = (size_t)B * d->fake.T.back() * d->cfg.out_channels;
    float* dl = d->alloc(n);
    size_t mark = d->ws_off;
    DCU(launch_adv_loss(d->fake.h.back(), n, 0.f, slope, 1.f, d->losses + 1, 0.5f * adv_weight, dl, st));
    int r = disc_backward(d, d->fake, dl, true, nullptr, 0, st);
    if (r) return r;
    d->ws_off = mark;
    r = disc_forward(d, x_real, B, L, true, 1, d->real, st);
    if (r) return r;
    mark = d->ws_off;
    DCU(launch_adv_loss(d->real.h.back(), n, 1.f, slope, 1.f, d->losses + 2, 0.5f * adv_weight, dl, st));
    r = disc_backward(d, d->real, dl, true, nullptr, 0, st);
    if (r) return r;
    d->ws_off = mark;
    if (d->ws_off > d->ws_cap) return dfail(EEGLDM_ERR_NOMEM, "discriminator workspace overflow");
    d->batches_tracked += 3;
    if (lr > 0.f) {
        d->step += 1;
        DCU(launch_adam(d->P, d->G, d->M, d->V, lr, b1, b2, eps, d->step, d->nP, st));
        d->images_dirty = true;
    }
    return EEGLDM_OK;
}

const float* disc_losses_dev(const eegldm_disc* d) { return d->losses; }

}  // namespace eegldm

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int eegldm_disc_create(const eegldm_disc_cfg* cfg, eegldm_disc** out) {
    if (!cfg || !out) return dfail(EEGLDM_ERR_INVALID, "null argument");
    auto* d = new (std::nothrow) eegldm_disc();
    if (!d) return dfail(EEGLDM_ERR_NOMEM, "out of host memory");
    d->cfg = *cfg;
    int r = build_disc(d);
    if (r) { delete d; return r; }
    *out = d;
    return EEGLDM_OK;
}
void eegldm_disc_destroy(eegldm_disc* d) { delete d; }
int eegldm_disc_num_params(const eegldm_disc* d) { return d ? (int)d->params.size() : 0; }
int eegldm_disc_param_info(const eegldm_disc* d, int i, const char** name, int64_t shape[4], int* ndim, int* is_buffer) {
    if (!d || i < 0 || i >= (int)d->params.size()) return dfail(EEGLDM_ERR_INVALID, "bad parameter index");
    const DParam& p = d->params[i];
    if (name) *name = p.name.c_str();
    if (ndim) *ndim = (int)p.shape.size();
    if (shape) for (size_t k = 0; k < p.shape.size() && k < 4; ++k) shape[k] = p.shape[k];
    if (is_buffer) *is_buffer = p.buffer ? 1 : 0;
    return EEGLDM_OK;
}
int eegldm_disc_load(eegldm_disc* d, const char* name, const float* host, const int64_t* shape, int ndim) {
    if (!d || !name || !host || (ndim > 0 && !shape)) return dfail(EEGLDM_ERR_INVALID, "null argument");
    std::string key(name);
    if (key.rfind("module.", 0) == 0) key = key.substr(7);
    auto it = d->index.find(key);
    if (it == d->index.end()) return dfail(EEGLDM_ERR_MISSING, "unexpected state_dict key: " + key);
    DParam& p = d->params[it->second];
    bool ok = (int)p.shape.size() == ndim;
    for (int i = 0; ok && i < ndim; ++i) ok = p.shape[i] == shape[i];
    if (!ok) return dfail(EEGLDM_ERR_SHAPE, "shape mismatch for " + key);
    p.data.assign(host, host + p.numel());
    p.loaded = true;
    d->finalized = false;
    return EEGLDM_OK;
}
int eegldm_disc_finalize(eegldm_disc* d) {
    if (!d) return dfail(EEGLDM_ERR_INVALID, "null handle");
    return finalize_disc(d);
}

int eegldm_disc_forward(eegldm_disc* d, const float* x_dev, float* logits_dev, int B, int L, int training, void* stream) {
    if (!d) return dfail(EEGLDM_ERR_INVALID, "null handle");
    if (!d->finalized) return dfail(EEGLDM_ERR_MISSING, "eegldm_disc_finalize has not been called");
    if (d->cfg.in_channels != 1 || d->cfg.out_channels != 1)
        return dfail(EEGLDM_ERR_INVALID, "discriminator forward: in / out channels must be 1 (config_aekl_eeg.yaml:34-35)");
    if (B < 0 || L < 8) return dfail(EEGLDM_ERR_SHAPE, "bad input shape");
    if (B == 0) return EEGLDM_OK;
    if (!x_dev || !logits_dev) return dfail(EEGLDM_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int r = ensure_ws(d, pass_floats(d, B, L, false));
    if (r) return r;
    d->ws_off = 0;
    DiscPass ps;
    r = disc_forward(d, x_dev, B, L, training != 0, 1, ps, st);
    if (r) return r;
    if (training) d->batches_tracked += 1;
    DCU(cudaMemcpyAsync(logits_dev, ps.h.back(), (size_t)B * ps.T.back() * sizeof(float), cudaMemcpyDeviceToDevice, st));
    d->fake.valid = d->real.valid = false;
    return EEGLDM_OK;
}

// PatchDiscriminator.forward(x)[-1] in training mode across an autograd boundary (the reference's own loop,
// train_autoencoderkl.py:213-234: logits = discriminator(x)[-1]; loss.backward(); optimizer_d.step()).  Up to two recorded passes
// (slot 0 / 1: the fake and the real batch of the discriminator part) live in the handle; each is consumed by ONE backward call.
int eegldm_disc_forward_train(eegldm_disc* d, const float* x_dev, float* logits_dev, int B, int L, int slot, void* stream) {
    if (!d) return dfail(EEGLDM_ERR_INVALID, "null handle");
    if (!d->finalized) return dfail(EEGLDM_ERR_MISSING, "eegldm_disc_finalize has not been called");
    if (d->cfg.in_channels != 1 || d->cfg.out_channels != 1)
        return dfail(EEGLDM_ERR_INVALID, "discriminator forward: in / out channels must be 1 (config_aekl_eeg.yaml:34-35)");
    if (slot != 0 && slot != 1) return dfail(EEGLDM_ERR_INVALID, "slot must be 0 or 1");
    if (B <= 0 || L < 8) return dfail(EEGLDM_ERR_SHAPE, "bad input shape");
    if (!x_dev || !logits_dev) return dfail(EEGLDM_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t half = (pass_floats(d, B, L, true) + (size_t)B * L + (size_t)B * (L / 8 + 8) + 1024 + 63) & ~size_t(63);   // 256-byte aligned halves
    if (half != d->slot_floats || 2 * half > d->ws_cap) {   // shape change: both recorded passes are gone
        float* old = d->ws;
        int r = ensure_ws(d, 2 * half);
        if (r) return r;
        (void)old;
        d->slot_floats = half;
        d->fake.valid = d->real.valid = false;
    }
    DiscPass& ps = slot ? d->real : d->fake;
    d->ws_off = (size_t)slot * half;
    float* xk = d->alloc((size_t)B * L);   // the pass keeps its own copy of the input (read again by the backward sweep)
    DCU(cudaMemcpyAsync(xk, x_dev, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToDevice, st));
    int r = disc_forward(d, xk, B, L, true, 1, ps, st);
    if (r) return r;
    d->pass_end[slot] = d->ws_off;
    d->batches_tracked += 1;
    DCU(cudaMemcpyAsync(logits_dev, ps.h.back(), (size_t)B * ps.T.back() * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return EEGLDM_OK;
}

// dlogits_dev [B, 1, L_out] -> dx_dev (nullable) [B, 1, L]; want_param_grads: the parameter gradients of THIS pass replace the handle's
// gradient buffer (eegldm_disc_export(h, 1, ...)).
int eegldm_disc_backward(eegldm_disc* d, int slot, const float* dlogits_dev, float* dx_dev, int want_param_grads, void* stream) {
    if (!d || !dlogits_dev) return dfail(EEGLDM_ERR_INVALID, "null argument");
    if (slot != 0 && slot != 1) return dfail(EEGLDM_ERR_INVALID, "slot must be 0 or 1");
    DiscPass& ps = slot ? d->real : d->fake;
    if (!ps.valid || !d->slot_floats) return dfail(EEGLDM_ERR_MISSING, "no recorded forward pass in this slot (eegldm_disc_forward_train)");
    cudaStream_t st = (cudaStream_t)stream;
    d->ws_off = d->pass_end[slot];
    const size_t n = (size_t)ps.B * ps.T.back() * d->cfg.out_channels;
    float* dl = d->alloc(n);
    DCU(cudaMemcpyAsync(dl, dlogits_dev, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (want_param_grads) DCU(cudaMemsetAsync(d->G, 0, d->nP * sizeof(float), st));
    int r = disc_backward(d, ps, dl, want_param_grads != 0, dx_dev, 0, st);
    ps.valid = false;
    if (r) return r;
    if (d->ws_off > (size_t)(slot + 1) * d->slot_floats) return dfail(EEGLDM_ERR_NOMEM, "discriminator workspace overflow");
    return EEGLDM_OK;
}

int eegldm_disc_set_math(eegldm_disc* d, int mode) {
    if (!d) return dfail(EEGLDM_ERR_INVALID, "null handle");
    if (mode != EEGLDM_MATH_FP32_SIMT && mode != EEGLDM_MATH_F16X3_TC) return dfail(EEGLDM_ERR_INVALID, "discriminator math: fp32 SIMT or f16x3");
    d->math = mode;
    return EEGLDM_OK;
}

int eegldm_disc_out_len(const eegldm_disc* d, int L) {
    if (!d) return -1;
    int T = L;
    for (auto& l : d->layers) T = out_len(T, l);
    return T;
}

// what: 0 = parameter / buffer value, 1 = gradient of the last discriminator step (parameters only); reference layout
int eegldm_disc_export(eegldm_disc* d, int what, const char* name, float* host_out) {
    if (!d || !name || !host_out) return dfail(EEGLDM_ERR_INVALID, "null argument");
    if (!d->finalized) return dfail(EEGLDM_ERR_MISSING, "eegldm_disc_finalize has not been called");
    if (what != 0 && what != 1) return dfail(EEGLDM_ERR_INVALID, "what must be 0 (value) or 1 (gradient)");
    const std::string key(name);
    DCU(cudaDeviceSynchronize());
    for (auto& l : d->layers) {
        const std::string& p = l.prefix;
        const float* base = what ? d->G : d->P;
        if (key == p + ".conv.weight") {
            std::vector<float> pk((size_t)l.cout * l.cin * 3);
            DCU(cudaMemcpy(pk.data(), base + l.ow, pk.size() * sizeof(float), cudaMemcpyDeviceToHost));
            for (int co = 0; co < l.cout; ++co)
                for (int ci = 0; ci < l.cin; ++ci)
                    for (int k = 0; k < 3; ++k) host_out[((size_t)co * l.cin + ci) * 3 + k] = pk[((size_t)ci * 3 + k) * l.cout + co];
            return EEGLDM_OK;
        }
        if (l.has_bias && key == p + ".conv.bias") { DCU(cudaMemcpy(host_out, base + l.ob, l.cout * sizeof(float), cudaMemcpyDeviceToHost)); return EEGLDM_OK; }
        if (l.has_norm) {
            if (key == p + ".adn.N.weight") { DCU(cudaMemcpy(host_out, base + l.og, l.cout * sizeof(float), cudaMemcpyDeviceToHost)); return EEGLDM_OK; }
            if (key == p + ".adn.N.bias") { DCU(cudaMemcpy(host_out, base + l.obe, l.cout * sizeof(float), cudaMemcpyDeviceToHost)); return EEGLDM_OK; }
            if (what == 0 && key == p + ".adn.N.running_mean") { DCU(cudaMemcpy(host_out, d->R + l.orm, l.cout * sizeof(float), cudaMemcpyDeviceToHost)); return EEGLDM_OK; }
            if (what == 0 && key == p + ".adn.N.running_var") { DCU(cudaMemcpy(host_out, d->R + l.orv, l.cout * sizeof(float), cudaMemcpyDeviceToHost)); return EEGLDM_OK; }
            if (what == 0 && key == p + ".adn.N.num_batches_tracked") { host_out[0] = (float)d->batches_tracked; return EEGLDM_OK; }
        }
    }
    return dfail(EEGLDM_ERR_MISSING, "unknown state_dict key: " + key);
}

}  // extern "C"
