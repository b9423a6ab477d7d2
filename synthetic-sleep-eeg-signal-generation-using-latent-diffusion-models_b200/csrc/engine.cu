// eegldm engine: host-side runtime of the B200-native latent-diffusion sampler and the C ABI
// declared in include/eegldm.h.
//
// What lives here (all of it host code; the device code is in kernels_simt.cu / conv_tc.cu):
//   * topology of the reference denoiser  UNetModel.__init__   src/models/unet.py:372-505
//     and of the KL autoencoder (monai-generative AutoencoderKL; in-tree ancestor src/models/ae_kl.py)
//   * state_dict ingestion (reference key grammar) and weight repacking
//   * the per-(B,T) launch plan: every activation is a slot of one arena in HBM, laid out
//     channels-last [B][T][C]; GroupNorm/SiLU/resample/concat/residual/time-embedding are folded
//     into the prologue/epilogue of the conv launches, so a ResBlock is 2 stat passes + 2 conv launches
//   * the DDIM loop (src/sample_trials.py:153-166): one CUDA graph per denoise step, the scheduler
//     update fused into the output conv's epilogue, per-step time-embedding rows precomputed
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "disc.h"
#include "eegldm.h"
#include "kernels.cuh"

namespace eegldm {
std::atomic<long long> g_launch_count{0};
}

using namespace eegldm;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
}  // namespace
namespace eegldm {
void set_last_error(const std::string& msg) { g_err = msg; }
}
namespace {
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return EEGLDM_ERR_CUDA;
}
#define CU(expr)                                                   \
    do {                                                           \
        cudaError_t e__ = (expr);                                  \
        if (e__ != cudaSuccess) return cuda_fail(e__, #expr);      \
    } while (0)

// ------------------------------------------------------------------------------------------------
// state_dict store
struct HostParam {
    std::string name;
    std::vector<int64_t> shape;
    std::vector<float> data;
    bool loaded = false;
    size_t numel() const {
        size_t n = 1;
        for (auto s : shape) n *= (size_t)s;
        return n;
    }
};

struct ParamSet {
    std::vector<HostParam> params;
    std::unordered_map<std::string, int> index;
    void add(const std::string& name, std::vector<int64_t> shape) {
        index[name] = (int)params.size();
        HostParam p;
        p.name = name;
        p.shape = std::move(shape);
        params.push_back(std::move(p));
    }
    void lin(const std::string& p, int i, int o) { add(p + ".weight", {o, i}); add(p + ".bias", {o}); }
    void conv(const std::string& p, int i, int o, int k) { add(p + ".weight", {o, i, k}); add(p + ".bias", {o}); }
    void gn(const std::string& p, int c) { add(p + ".weight", {c}); add(p + ".bias", {c}); }
    int load(const char* name, const float* host, const int64_t* shape, int ndim) {
        if (!name || !host || !shape) return fail(EEGLDM_ERR_INVALID, "null argument");
        std::string key(name);
        if (key.rfind("module.", 0) == 0) key = key.substr(7);  // DataParallel prefix (MSSIM_reconstruction.py:66-69)
        auto it = index.find(key);
        if (it == index.end()) return fail(EEGLDM_ERR_MISSING, "unexpected state_dict key: " + key);
        HostParam& p = params[it->second];
        bool ok = (int)p.shape.size() == ndim;
        for (int i = 0; ok && i < ndim; ++i) ok = p.shape[i] == shape[i];
        if (!ok) return fail(EEGLDM_ERR_SHAPE, "shape mismatch for " + key);
        p.data.assign(host, host + p.numel());
        p.loaded = true;
        return EEGLDM_OK;
    }
    const std::vector<float>& get(const std::string& name) const { return params[index.at(name)].data; }
    int check_all_loaded() const {
        for (auto& p : params)
            if (!p.loaded) return fail(EEGLDM_ERR_MISSING, "missing state_dict key: " + p.name);
        return EEGLDM_OK;
    }
};

// Packed device weights: staged on the host, uploaded once.
struct WeightPool {
    std::vector<float> stage;
    float* dev = nullptr;
    size_t push(const float* src, size_t n) {
        size_t off = (stage.size() + 63) & ~size_t(63);
        stage.resize(off + n);
        std::memcpy(stage.data() + off, src, n * sizeof(float));
        return off;
    }
    size_t push(const std::vector<float>& v) { return push(v.data(), v.size()); }
    size_t push_u16(const std::vector<uint16_t>& v) {   // raw bf16 images (tcgen05 weight tiles), 256-byte aligned
        const size_t nfl = (v.size() + 1) / 2;
        size_t off = (stage.size() + 63) & ~size_t(63);
        stage.resize(off + nfl);
        std::memcpy(stage.data() + off, v.data(), v.size() * sizeof(uint16_t));
        return off;
    }
    int upload() {
        if (dev) { cudaFree(dev); dev = nullptr; }
        if (stage.empty()) return EEGLDM_OK;
        CU(cudaMalloc(&dev, stage.size() * sizeof(float)));
        CU(cudaMemcpy(dev, stage.data(), stage.size() * sizeof(float), cudaMemcpyHostToDevice));
        return EEGLDM_OK;
    }
    const float* at(size_t off) const { return dev + off; }
    ~WeightPool() { if (dev) cudaFree(dev); }
};

// [Cout][Cin][k] (PyTorch)  ->  [(ci*k + kk)][Cout]  (SIMT conv kernel layout)
std::vector<float> pack_conv(const std::vector<float>& w, int Cout, int Cin, int k) {
    std::vector<float> out((size_t)Cout * Cin * k);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int kk = 0; kk < k; ++kk)
                out[((size_t)ci * k + kk) * Cout + co] = w[((size_t)co * Cin + ci) * k + kk];
    return out;
}

// ------------------------------------------------------------------------------------------------
// Plan-time arena: activations are offsets into one device buffer; slots are recycled as soon as
// their last consumer has been planned (launches are stream-ordered, so reuse is safe).
struct Planner {
    std::vector<std::pair<size_t, size_t>> free_;  // (off, n)
    size_t top = 0;
    size_t alloc(size_t n) {
        n = (n + 63) & ~size_t(63);
        int best = -1;
        for (int i = 0; i < (int)free_.size(); ++i)
            if (free_[i].second >= n && (best < 0 || free_[i].second < free_[best].second)) best = i;
        if (best >= 0) {
            size_t off = free_[best].first;
            if (free_[best].second == n) free_.erase(free_.begin() + best);
            else { free_[best].first += n; free_[best].second -= n; }
            return off;
        }
        size_t off = top;
        top += n;
        return off;
    }
    void release(size_t off, size_t n) {
        n = (n + 63) & ~size_t(63);
        free_.push_back({off, n});
        // coalesce neighbours
        std::sort(free_.begin(), free_.end());
        std::vector<std::pair<size_t, size_t>> m;
        for (auto& f : free_) {
            if (!m.empty() && m.back().first + m.back().second == f.first) m.back().second += f.second;
            else m.push_back(f);
        }
        if (!m.empty() && m.back().first + m.back().second == top) { top = m.back().first; m.pop_back(); }
        free_.swap(m);
    }
};

struct Buf {
    Planner* pl;
    size_t off, n;
    Buf(Planner* p, size_t n_) : pl(p), off(p->alloc(n_)), n(n_) {}
    ~Buf() { pl->release(off, n); }
};
struct Act {  // channels-last activation [B][T][C]
    std::shared_ptr<Buf> buf;
    const float* ext = nullptr;  // external (caller-owned) tensor instead of an arena slot
    int C = 0, T = 0;
    // GroupNorm statistics emitted by the producing conv's epilogue: [B][gn_nsplit][gn_G][3] records (count, mean, M2)
    std::shared_ptr<Buf> gn_part;
    int gn_nsplit = 0, gn_G = 0;
};

using OpFn = std::function<cudaError_t(cudaStream_t)>;

// per-launch bookkeeping for the live profile (eegldm_profile_*): algorithmic FLOPs / HBM bytes
enum OpKind : int { OP_CONV = 0, OP_GN = 1, OP_ATTN = 2, OP_OTHER = 3, OP_SPLIT = 4, OP_CONV_SIMT = 5, OP_NKIND = 6 };   // OP_CONV: tcgen05 convs only
struct OpMeta { int kind; double flops, bytes; };

struct Builder {
    Planner pl;
    float* base = nullptr;  // null during the sizing pass
    std::vector<OpFn> ops;
    std::vector<OpMeta> meta;
    int n_kernels = 0;
    int B = 0;
    int math = EEGLDM_MATH_FP32_SIMT;
    int* range_flag = nullptr;   // device word the f16x3 operand producers raise when a value is outside the fp16 range
    size_t peak = 0;
    Act act(int C, int T) {
        Act a;
        a.buf = std::make_shared<Buf>(&pl, (size_t)B * T * C);
        a.C = C; a.T = T;
        peak = std::max(peak, pl.top);
        return a;
    }
    std::shared_ptr<Buf> scratch(size_t n) {
        auto b = std::make_shared<Buf>(&pl, n);
        peak = std::max(peak, pl.top);
        return b;
    }
    float* ptr(const std::shared_ptr<Buf>& b) const { return base + b->off; }
    const float* ptr(const Act& a) const { return a.ext ? a.ext : base + a.buf->off; }
    float* wptr(const Act& a) const { return base + a.buf->off; }
    void add(OpFn f, int kernels, int kind = OP_OTHER, double flops = 0, double bytes = 0) {
        if (base) { ops.push_back(std::move(f)); meta.push_back({kind, flops, bytes}); }
        n_kernels += kernels;
    }
};

// GroupNorm statistics of (virtual concat of) x0,x1 -> per-(sample,channel) scale/shift
struct ScaleShift { std::shared_ptr<Buf> buf; const float* scale; const float* shift; };
ScaleShift plan_gn(Builder& bd, const Act& x0, const Act* x1, int G, const float* gamma, const float* beta, float eps) {
    const int C = x0.C + (x1 ? x1->C : 0);
    ScaleShift ss;
    ss.buf = bd.scratch((size_t)bd.B * C * 2);
    GnParams p{};
    p.src0 = bd.ptr(x0); p.C0 = x0.C;
    p.src1 = x1 ? bd.ptr(*x1) : nullptr; p.C1 = x1 ? x1->C : 0;
    p.T = x0.T; p.G = G; p.gamma = gamma; p.beta = beta; p.eps = eps; p.B = bd.B;
    p.scale = bd.ptr(ss.buf);
    p.shift = p.scale + (size_t)bd.B * C;
    ss.scale = p.scale; ss.shift = p.shift;
    // The producers' epilogues already wrote the statistics as (count, mean, M2) records at THEIR record width (4 or 8 channels,
    // plan_conv): no pass over the tensor when every group of this norm is a whole number of records -- also for a virtual concat
    // whose groups straddle the two sources (768 = 512 + 256 in groups of 24, 384 = 256 + 128 in groups of 12).
    {
        const int cpg = C / G;
        const bool have0 = x0.gn_part != nullptr, have1 = !x1 || x1->gn_part != nullptr;
        const int w0 = have0 ? x0.C / x0.gn_G : 0, w1 = x1 && have1 ? x1->C / x1->gn_G : 0;
        bool ok = have0 && have1 && G <= 256 && C % G == 0 && cpg % w0 == 0;
        if (ok && x1) ok = cpg % w1 == 0 && x0.C % w1 == 0;
        if (ok) {
            p.nsplit = x0.gn_nsplit; p.nsplit1 = x1 ? x1->gn_nsplit : 0;
            p.partial = bd.ptr(x0.gn_part); p.rec_G0 = x0.gn_G;
            if (x1) { p.partial1 = bd.ptr(x1->gn_part); p.rec_G1 = x1->gn_G; }
            bd.add([p](cudaStream_t st) { return launch_groupnorm_finalize(p, st); }, 1, OP_GN, 0.0,
                   4.0 * bd.B * (3.0 * p.nsplit * (x0.gn_G + (x1 ? x1->gn_G : 0)) + 2.0 * C));
            return ss;
        }
    }
    p.nsplit = groupnorm_nsplit(C, x0.T, G);
    auto part = bd.scratch((size_t)bd.B * p.nsplit * G * 3);
    p.partial = bd.ptr(part);
    bd.add([p](cudaStream_t st) { return launch_groupnorm(p, st); }, 2, OP_GN, 3.0 * bd.B * C * x0.T,
           4.0 * bd.B * C * (x0.T + 2.0));
    return ss;
}

int resampled_len(int T, int mode) { return mode == RS_AVGPOOL2 ? T / 2 : (mode == RS_NEAREST2 ? T * 2 : T); }

ConvSeg make_seg(Builder& bd, const Act& x0, const Act* x1, const ScaleShift* ss, int silu, int resample,
                 const float* w, int taps) {
    ConvSeg s{};
    s.src0 = bd.ptr(x0); s.C0 = x0.C;
    s.src1 = x1 ? bd.ptr(*x1) : nullptr; s.C1 = x1 ? x1->C : 0;
    s.scale = ss ? ss->scale : nullptr; s.shift = ss ? ss->shift : nullptr;
    s.silu = silu; s.resample = resample; s.Tin = x0.T; s.w = w; s.taps = taps;
    return s;
}

// ================================================================================================ UNet
struct ULayer {
    enum Kind { CONV_IN, RES, ATTN, DOWN, UP } kind;
    std::string prefix;
    int cin = 0, cout = 0, mode = RS_NONE;  // RES
    int ch = 0, heads = 1;                  // ATTN / DOWN / UP
    int use_conv = 0;
    // device weights (valid after finalize)
    const float *g1 = nullptr, *be1 = nullptr, *w1 = nullptr, *b1 = nullptr;
    const float *g2 = nullptr, *be2 = nullptr, *w2 = nullptr, *b2 = nullptr, *wskip = nullptr;
    int emb_off = 0;
    const float *wqkv = nullptr, *bqkv = nullptr, *wproj = nullptr, *bproj = nullptr;
    size_t o_g1, o_be1, o_w1, o_b1, o_g2, o_be2, o_w2, o_b2, o_wskip, o_wqkv, o_bqkv, o_wproj, o_bproj;
    // tcgen05 weight images (null when the layer's channel counts are not tensor-pipe eligible)
    const uint8_t *t_w1 = nullptr, *t_w2 = nullptr, *t_wskip = nullptr, *t_wqkv = nullptr, *t_wproj = nullptr;
    const uint8_t* t_w1p = nullptr;   // up-sampling ResBlock: polyphase image of in_layers' conv (pack_conv_tc_poly)
    size_t ot_w1 = 0, ot_w2 = 0, ot_wskip = 0, ot_wqkv = 0, ot_wproj = 0, ot_w1p = 0;
    bool h_w1 = false, h_w2 = false, h_wskip = false, h_wqkv = false, h_wproj = false, h_w1p = false;
};

struct GraphEntry {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int n_kernels = 0;
};

}  // namespace

struct UnetTrain;                        // latent-diffusion training state (unet_train.inc)
void unet_train_free(UnetTrain* t);

struct eegldm_unet {
    eegldm_unet_cfg cfg{};
    ParamSet ps;
    std::vector<std::vector<ULayer>> input_blocks, output_blocks;
    std::vector<ULayer> middle;
    int final_ch = 0, ted = 0, emb_total = 0;
    bool finalized = false;
    eegldm_math math = EEGLDM_MATH_FP32_SIMT;
    WeightPool pool;
    const float *te0_w = nullptr, *te0_b = nullptr, *te2_w = nullptr, *te2_b = nullptr;
    const float *emb_w = nullptr, *emb_b = nullptr;  // concatenated emb_layers.1 of every ResBlock
    const float *out_g = nullptr, *out_be = nullptr, *out_w = nullptr, *out_b = nullptr;
    // runtime buffers
    float* arena = nullptr; size_t arena_cap = 0;
    float* temb_fwd = nullptr; size_t temb_fwd_cap = 0;    // forward(): [nt][emb_total]
    float* tscratch = nullptr; size_t tscratch_cap = 0;    // time MLP scratch
    float* temb_step = nullptr;                            // ddim: current row [emb_total]
    float* temb_table = nullptr; size_t temb_table_cap = 0;  // ddim: [n_steps][emb_total]
    float* coef_table = nullptr; size_t coef_table_cap = 0;  // ddim: [n_steps][2]
    float* coef_cur = nullptr;                             // [2]
    int* step_ctr = nullptr;
    int* range_flag = nullptr;                             // f16x3: raised by operand producers on |x| >= 65504 / NaN
    cudaStream_t cap_stream = nullptr;                     // private stream used only for graph capture
    cudaStream_t cap_stream2 = nullptr;                    // second capture stream: the other batch half of a two-lane step
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    float* xbuf = nullptr; size_t xbuf_cap = 0;            // ddim state [B][T][z]
    float* xtmp = nullptr; size_t xtmp_cap = 0;            // NCL<->NLC staging
    float* host_in = nullptr; size_t host_in_cap = 0;      // eegldm_ddim_sample_host: device copies of the host buffers
    float* host_out = nullptr; size_t host_out_cap = 0;
    std::vector<float> table_key;                          // identifies the cached temb/coef tables
    std::map<std::pair<int, int>, GraphEntry> graphs;      // (B,T) -> one denoise step
    UnetTrain* train = nullptr;                            // created by the first eegldm_unet_train_step
    ~eegldm_unet() {
        drop_graphs();
        if (train) unet_train_free(train);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        if (cap_stream2) cudaStreamDestroy(cap_stream2);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        for (void* p : {(void*)arena, (void*)temb_fwd, (void*)tscratch, (void*)temb_step, (void*)temb_table,
                        (void*)coef_table, (void*)coef_cur, (void*)step_ctr, (void*)range_flag, (void*)xbuf, (void*)xtmp, (void*)host_in,
                        (void*)host_out})
            if (p) cudaFree(p);
    }
    void drop_graphs() {
        for (auto& g : graphs) {
            if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
            if (g.second.graph) cudaGraphDestroy(g.second.graph);
        }
        graphs.clear();
    }
};

namespace {

bool g_conv_qkv_fused = true;  // the qkv conv writes attention operand images directly (f16x3; eegldm_set_conv_tuning bit 1): no qkv_split pass
bool g_attn_direct = false;    // the tcgen05 attention splits fp32 q, k, v itself (eegldm_set_conv_tuning bit 4): measured no faster
bool g_attn_u_fused = true;    // the tcgen05 attention writes proj_out's operand image instead of fp32 (eegldm_set_conv_tuning bit 3)
bool g_conv_direct = true;     // tensor-pipe convs produce their activation operands in-kernel (no act_split pre-pass)
bool g_conv_gn_fine = true;        // epilogue GroupNorm records 4 / 8 channels wide (eegldm_set_conv_tuning bit 8 = the consumer's group width instead)
bool g_conv_gn_tile = true;        // epilogue GroupNorm records per 128-row tile where tiles never straddle samples (eegldm_set_conv_tuning bit 10 = off)
bool g_conv_poly = true;           // polyphase form of the up-sampling ResBlocks' first conv (eegldm_set_conv_tuning bit 9 = off)
bool g_conv_direct_wide = false;   // fused producer also for 1x1 convs with more than two N tiles (qkv): eegldm_set_conv_tuning bit 7
bool g_conv_gn_fused = true;   // tensor-pipe convs emit the GroupNorm statistics of their output (eegldm_set_conv_tuning)
bool g_graphs_enabled = true;
// Denoise-step graph of eegldm_ddim_sample: 2 = the batch is planned as two independent halves captured on two streams
// (fork / join inside the graph), so the HBM-bound passes of one half (GroupNorm, activation split) run underneath the
// tensor-bound convolutions of the other half instead of in series with them; 1 = one chain (default: the step runs
// at the board's power cap, ~1.63 GHz of 1.965, so overlapping the two kinds of work buys nothing -- measured 312 vs 320
// windows/s at B=1024, profiles/r01_summary.md).
int g_sample_lanes = 1;

template <class T>
int ensure(T*& p, size_t& cap, size_t n) {
    if (n <= cap) return EEGLDM_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    CU(cudaMalloc((void**)&p, n * sizeof(T)));
    cap = n;
    return EEGLDM_OK;
}

int n_heads_for(const eegldm_unet_cfg& c, int ch, bool upsample) {
    if (c.num_head_channels != -1) return c.num_head_channels > 0 ? ch / c.num_head_channels : 0;
    return (upsample && c.num_heads_upsample != -1) ? c.num_heads_upsample : c.num_heads;
}

// UNetModel.__init__ topology, unet.py:382-505
int build_unet_topology(eegldm_unet* h) {
    const auto& c = h->cfg;
    if (c.use_scale_shift_norm) return fail(EEGLDM_ERR_INVALID, "use_scale_shift_norm is not supported (no reference config enables it)");
    if (c.n_channel_mult < 1 || c.n_channel_mult > 8 || c.n_attention_resolutions < 0 || c.n_attention_resolutions > 8)
        return fail(EEGLDM_ERR_INVALID, "bad channel_mult / attention_resolutions length");
    if (c.model_channels % 32 != 0) return fail(EEGLDM_ERR_INVALID, "model_channels must be a multiple of 32 (GroupNorm32)");
    if (c.in_channels < 1 || c.out_channels < 1 || c.num_res_blocks < 1) return fail(EEGLDM_ERR_INVALID, "bad channel / block counts");
    if (c.num_head_channels == -1 && c.num_heads < 1) return fail(EEGLDM_ERR_INVALID, "num_heads must be >= 1");
    const int mc = c.model_channels;
    auto has_att = [&](int ds) {
        for (int i = 0; i < c.n_attention_resolutions; ++i)
            if (c.attention_resolutions[i] == ds) return true;
        return false;
    };
    auto res = [&](const std::string& p, int cin, int cout, int mode) {
        ULayer l; l.kind = ULayer::RES; l.prefix = p; l.cin = cin; l.cout = cout; l.mode = mode; return l;
    };
    auto attn = [&](const std::string& p, int ch, bool upsample = false) {
        ULayer l; l.kind = ULayer::ATTN; l.prefix = p; l.ch = ch; l.heads = n_heads_for(c, ch, upsample); return l;
    };
    h->ted = mc * 4;
    {
        ULayer l; l.kind = ULayer::CONV_IN; l.prefix = "input_blocks.0.0"; l.cin = c.in_channels; l.cout = mc;
        h->input_blocks.push_back({l});
    }
    std::vector<int> chans{mc};
    int ch = mc, ds = 1;
    for (int level = 0; level < c.n_channel_mult; ++level) {
        const int m = c.channel_mult[level];
        for (int r = 0; r < c.num_res_blocks; ++r) {
            const int idx = (int)h->input_blocks.size();
            std::vector<ULayer> layers{res("input_blocks." + std::to_string(idx) + ".0", ch, m * mc, RS_NONE)};
            ch = m * mc;
            if (has_att(ds)) layers.push_back(attn("input_blocks." + std::to_string(idx) + ".1", ch));
            h->input_blocks.push_back(layers);
            chans.push_back(ch);
        }
        if (level != c.n_channel_mult - 1) {
            const int idx = (int)h->input_blocks.size();
            const std::string p = "input_blocks." + std::to_string(idx) + ".0";
            if (c.resblock_updown) h->input_blocks.push_back({res(p, ch, ch, RS_AVGPOOL2)});
            else { ULayer l; l.kind = ULayer::DOWN; l.prefix = p; l.ch = ch; l.use_conv = c.conv_resample; h->input_blocks.push_back({l}); }
            chans.push_back(ch);
            ds *= 2;
        }
    }
    h->middle = {res("middle_block.0", ch, ch, RS_NONE), attn("middle_block.1", ch), res("middle_block.2", ch, ch, RS_NONE)};
    for (int level = c.n_channel_mult - 1; level >= 0; --level) {
        const int m = c.channel_mult[level];
        for (int i = 0; i <= c.num_res_blocks; ++i) {
            const int ich = chans.back(); chans.pop_back();
            const int idx = (int)h->output_blocks.size();
            const std::string bp = "output_blocks." + std::to_string(idx) + ".";
            std::vector<ULayer> layers{res(bp + "0", ch + ich, mc * m, RS_NONE)};
            ch = mc * m;
            if (has_att(ds)) layers.push_back(attn(bp + std::to_string(layers.size()), ch, true));
            if (level && i == c.num_res_blocks) {
                const std::string p = bp + std::to_string(layers.size());
                if (c.resblock_updown) layers.push_back(res(p, ch, ch, RS_NEAREST2));
                else { ULayer l; l.kind = ULayer::UP; l.prefix = p; l.ch = ch; l.use_conv = c.conv_resample; layers.push_back(l); }
                ds /= 2;
            }
            h->output_blocks.push_back(layers);
        }
    }
    h->final_ch = ch;
    // state_dict entries in the reference's registration order
    ParamSet& ps = h->ps;
    ps.lin("time_embed.0", mc, h->ted);
    ps.lin("time_embed.2", h->ted, h->ted);
    auto reg = [&](const ULayer& l) -> int {
        const std::string& p = l.prefix;
        switch (l.kind) {
            case ULayer::CONV_IN: ps.conv(p, l.cin, l.cout, 3); break;
            case ULayer::RES:
                if (l.cin % 32 || l.cout % 32) return fail(EEGLDM_ERR_INVALID, "ResBlock channels must be multiples of 32");
                ps.gn(p + ".in_layers.0", l.cin); ps.conv(p + ".in_layers.2", l.cin, l.cout, 3);
                ps.lin(p + ".emb_layers.1", h->ted, l.cout);
                ps.gn(p + ".out_layers.0", l.cout); ps.conv(p + ".out_layers.3", l.cout, l.cout, 3);
                if (l.cin != l.cout) ps.conv(p + ".skip_connection", l.cin, l.cout, 1);
                break;
            case ULayer::ATTN:
                if (l.heads < 1 || l.ch % l.heads) return fail(EEGLDM_ERR_INVALID, "attention channels not divisible by heads");
                ps.gn(p + ".norm", l.ch); ps.conv(p + ".qkv", l.ch, 3 * l.ch, 1); ps.conv(p + ".proj_out", l.ch, l.ch, 1);
                break;
            case ULayer::DOWN: if (l.use_conv) ps.conv(p + ".op", l.ch, l.ch, 3); break;
            case ULayer::UP: if (l.use_conv) ps.conv(p + ".conv", l.ch, l.ch, 3); break;
        }
        return EEGLDM_OK;
    };
    for (auto& b : h->input_blocks) for (auto& l : b) { int r = reg(l); if (r) return r; }
    for (auto& l : h->middle) { int r = reg(l); if (r) return r; }
    for (auto& b : h->output_blocks) for (auto& l : b) { int r = reg(l); if (r) return r; }
    ps.gn("out.0", h->final_ch);
    ps.conv("out.2", mc, c.out_channels, 3);
    if (h->final_ch != mc) return fail(EEGLDM_ERR_INVALID, "channel_mult[0] must be 1 (out conv takes model_channels)");
    return EEGLDM_OK;
}

template <class F>
void for_each_layer(eegldm_unet* h, F f) {
    for (auto& b : h->input_blocks) for (auto& l : b) f(l);
    for (auto& l : h->middle) f(l);
    for (auto& b : h->output_blocks) for (auto& l : b) f(l);
}

int finalize_unet(eegldm_unet* h) {
    int r = h->ps.check_all_loaded();
    if (r) return r;
    h->drop_graphs();
    h->table_key.clear();
    WeightPool& wp = h->pool;
    wp.stage.clear();
    const ParamSet& ps = h->ps;
    const size_t o_te0w = wp.push(ps.get("time_embed.0.weight")), o_te0b = wp.push(ps.get("time_embed.0.bias"));
    const size_t o_te2w = wp.push(ps.get("time_embed.2.weight")), o_te2b = wp.push(ps.get("time_embed.2.bias"));
    // concatenated emb_layers.1 (ResBlock time projections): one batched linear per step
    std::vector<float> embw, embb;
    for_each_layer(h, [&](ULayer& l) {
        if (l.kind != ULayer::RES) return;
        l.emb_off = (int)embb.size();
        const auto& w = ps.get(l.prefix + ".emb_layers.1.weight");
        const auto& b = ps.get(l.prefix + ".emb_layers.1.bias");
        embw.insert(embw.end(), w.begin(), w.end());
        embb.insert(embb.end(), b.begin(), b.end());
    });
    h->emb_total = (int)embb.size();
    const size_t o_embw = wp.push(embw), o_embb = wp.push(embb);
    // tcgen05 images: bf16 hi/lo split, packed as shared-memory stage images (conv_tc.cu)
    auto tc_pack = [&](const std::string& name, int cout, int cin, int k, int stages, size_t& off, bool& has) {
        has = h->math != EEGLDM_MATH_FP32_SIMT && conv_tc_eligible(cin, 0, cout, 16, k, 1);
        if (has && h->math == EEGLDM_MATH_F16X3_TC) {
            // f16x3 operand range: a weight with |w| >= 65504 (or NaN) cannot be split; that layer stays on the fp32 SIMT kernel
            for (float v : ps.get(name)) if (!(std::fabs(v) < 65504.f)) { has = false; break; }
        }
        if (!has) return;
        std::vector<uint16_t> img;
        pack_conv_tc(ps.get(name).data(), cout, cin, k, h->math == EEGLDM_MATH_F16X3_TC, img);
        off = wp.push_u16(img);
    };
    for_each_layer(h, [&](ULayer& l) {
        const std::string& p = l.prefix;
        switch (l.kind) {
            case ULayer::RES: {
                // weight stages per tile = sum over K segments of (Cin/32)*taps; conv2 and its skip segment share one tile shape
                const int st1 = l.cin / TC_BK * 3, st2 = l.cout / TC_BK * 3 + (l.cin != l.cout ? l.cin / TC_BK : 0);
                tc_pack(p + ".in_layers.2.weight", l.cout, l.cin, 3, st1, l.ot_w1, l.h_w1);
                if (l.h_w1 && l.mode == RS_NEAREST2 && l.cout % 256 == 0) {   // up-sampling ResBlock: polyphase image as well
                    std::vector<uint16_t> img;
                    pack_conv_tc_poly(ps.get(p + ".in_layers.2.weight").data(), l.cout, l.cin, h->math == EEGLDM_MATH_F16X3_TC, img);
                    l.ot_w1p = wp.push_u16(img);
                    l.h_w1p = true;
                }
                tc_pack(p + ".out_layers.3.weight", l.cout, l.cout, 3, st2, l.ot_w2, l.h_w2);
                if (l.cin != l.cout) tc_pack(p + ".skip_connection.weight", l.cout, l.cin, 1, st2, l.ot_wskip, l.h_wskip);
                break;
            }
            case ULayer::ATTN:
                tc_pack(p + ".qkv.weight", 3 * l.ch, l.ch, 1, l.ch / TC_BK, l.ot_wqkv, l.h_wqkv);
                tc_pack(p + ".proj_out.weight", l.ch, l.ch, 1, l.ch / TC_BK, l.ot_wproj, l.h_wproj);
                break;
            case ULayer::UP:
                if (l.use_conv) tc_pack(p + ".conv.weight", l.ch, l.ch, 3, l.ch / TC_BK * 3, l.ot_w1, l.h_w1);
                break;
            default: break;
        }
    });
    for_each_layer(h, [&](ULayer& l) {
        const std::string& p = l.prefix;
        switch (l.kind) {
            case ULayer::CONV_IN:
                l.o_w1 = wp.push(pack_conv(ps.get(p + ".weight"), l.cout, l.cin, 3));
                l.o_b1 = wp.push(ps.get(p + ".bias"));
                break;
            case ULayer::RES: {
                l.o_g1 = wp.push(ps.get(p + ".in_layers.0.weight")); l.o_be1 = wp.push(ps.get(p + ".in_layers.0.bias"));
                l.o_w1 = wp.push(pack_conv(ps.get(p + ".in_layers.2.weight"), l.cout, l.cin, 3));
                l.o_b1 = wp.push(ps.get(p + ".in_layers.2.bias"));
                l.o_g2 = wp.push(ps.get(p + ".out_layers.0.weight")); l.o_be2 = wp.push(ps.get(p + ".out_layers.0.bias"));
                l.o_w2 = wp.push(pack_conv(ps.get(p + ".out_layers.3.weight"), l.cout, l.cout, 3));
                std::vector<float> b2 = ps.get(p + ".out_layers.3.bias");
                if (l.cin != l.cout) {
                    l.o_wskip = wp.push(pack_conv(ps.get(p + ".skip_connection.weight"), l.cout, l.cin, 1));
                    const auto& bs = ps.get(p + ".skip_connection.bias");
                    for (int i = 0; i < l.cout; ++i) b2[i] += bs[i];
                }
                l.o_b2 = wp.push(b2);
                break;
            }
            case ULayer::ATTN:
                l.o_g1 = wp.push(ps.get(p + ".norm.weight")); l.o_be1 = wp.push(ps.get(p + ".norm.bias"));
                l.o_wqkv = wp.push(pack_conv(ps.get(p + ".qkv.weight"), 3 * l.ch, l.ch, 1));
                l.o_bqkv = wp.push(ps.get(p + ".qkv.bias"));
                l.o_wproj = wp.push(pack_conv(ps.get(p + ".proj_out.weight"), l.ch, l.ch, 1));
                l.o_bproj = wp.push(ps.get(p + ".proj_out.bias"));
                break;
            case ULayer::DOWN:
                if (l.use_conv) { l.o_w1 = wp.push(pack_conv(ps.get(p + ".op.weight"), l.ch, l.ch, 3)); l.o_b1 = wp.push(ps.get(p + ".op.bias")); }
                break;
            case ULayer::UP:
                if (l.use_conv) { l.o_w1 = wp.push(pack_conv(ps.get(p + ".conv.weight"), l.ch, l.ch, 3)); l.o_b1 = wp.push(ps.get(p + ".conv.bias")); }
                break;
        }
    });
    const size_t o_og = wp.push(ps.get("out.0.weight")), o_obe = wp.push(ps.get("out.0.bias"));
    const size_t o_ow = wp.push(pack_conv(ps.get("out.2.weight"), h->cfg.out_channels, h->cfg.model_channels, 3));
    const size_t o_ob = wp.push(ps.get("out.2.bias"));
    r = wp.upload();
    if (r) return r;
    h->te0_w = wp.at(o_te0w); h->te0_b = wp.at(o_te0b); h->te2_w = wp.at(o_te2w); h->te2_b = wp.at(o_te2b);
    h->emb_w = wp.at(o_embw); h->emb_b = wp.at(o_embb);
    h->out_g = wp.at(o_og); h->out_be = wp.at(o_obe); h->out_w = wp.at(o_ow); h->out_b = wp.at(o_ob);
    for_each_layer(h, [&](ULayer& l) {
        auto tcp = [&](bool has, size_t off) { return has ? reinterpret_cast<const uint8_t*>(wp.at(off)) : nullptr; };
        l.t_w1 = tcp(l.h_w1, l.ot_w1); l.t_w2 = tcp(l.h_w2, l.ot_w2); l.t_wskip = tcp(l.h_wskip, l.ot_wskip);
        l.t_w1p = tcp(l.h_w1p, l.ot_w1p);
        l.t_wqkv = tcp(l.h_wqkv, l.ot_wqkv); l.t_wproj = tcp(l.h_wproj, l.ot_wproj);
        switch (l.kind) {
            case ULayer::CONV_IN: l.w1 = wp.at(l.o_w1); l.b1 = wp.at(l.o_b1); break;
            case ULayer::RES:
                l.g1 = wp.at(l.o_g1); l.be1 = wp.at(l.o_be1); l.w1 = wp.at(l.o_w1); l.b1 = wp.at(l.o_b1);
                l.g2 = wp.at(l.o_g2); l.be2 = wp.at(l.o_be2); l.w2 = wp.at(l.o_w2); l.b2 = wp.at(l.o_b2);
                l.wskip = l.cin != l.cout ? wp.at(l.o_wskip) : nullptr;
                break;
            case ULayer::ATTN:
                l.g1 = wp.at(l.o_g1); l.be1 = wp.at(l.o_be1); l.wqkv = wp.at(l.o_wqkv); l.bqkv = wp.at(l.o_bqkv);
                l.wproj = wp.at(l.o_wproj); l.bproj = wp.at(l.o_bproj);
                break;
            case ULayer::DOWN: case ULayer::UP:
                if (l.use_conv) { l.w1 = wp.at(l.o_w1); l.b1 = wp.at(l.o_b1); }
                break;
        }
    });
    if (!h->temb_step) CU(cudaMalloc((void**)&h->temb_step, (size_t)std::max(h->emb_total, 1) * sizeof(float)));
    else { cudaFree(h->temb_step); h->temb_step = nullptr; CU(cudaMalloc((void**)&h->temb_step, (size_t)std::max(h->emb_total, 1) * sizeof(float))); }
    if (!h->coef_cur) CU(cudaMalloc((void**)&h->coef_cur, 2 * sizeof(float)));
    if (!h->step_ctr) CU(cudaMalloc((void**)&h->step_ctr, sizeof(int)));
    if (!h->range_flag) CU(cudaMalloc((void**)&h->range_flag, sizeof(int)));
    CU(cudaMemset(h->range_flag, 0, sizeof(int)));
    h->finalized = true;
    return EEGLDM_OK;
}

// ---- launch plan of one UNetModel.forward (unet.py:512-563) on channels-last tensors -----------
struct UNetIO {
    const float* x;       // [B][T][in_ch]
    float* out;           // [B][T][out_ch]
    const float* temb;    // [nt][emb_total]
    int temb_stride;      // 0 (shared timestep) or emb_total
    const float* ddim_x;  // non-null: out = coef[0]*ddim_x + coef[1]*model_output
    const float* ddim_coef;
};

// tcgen05 plumbing between the two convs of a ResBlock: conv1's activation pre-pass also emits the raw twin of x
// that conv2's skip_connection segment consumes, so x is read once.
struct TcShare {
    std::shared_ptr<Buf> raw;   // U image of the raw (resampled) block input
    bool want_raw = false;      // conv1: produce it;   conv2: consume it for segment 1
};

// out_act / gn_G: when the tensor-pipe path is taken and the output's next consumer is a GroupNorm(gn_G), the conv's
// epilogue also emits the statistics (attached to *out_act; plan_gn then skips its pass over the tensor).
// qkv: the output is written as attention operand images instead of fp32 (tensor-pipe f16x3 path only; the caller checks).
struct QkvOut { uint8_t* dst; int H, ch; };
// premade_u0: segment 0's operand image already exists (written by the attention kernel's epilogue): no pre-pass, no producer.
void plan_conv(Builder& bd, ConvParams p, const uint8_t* tw0 = nullptr, const uint8_t* tw1 = nullptr, TcShare* share = nullptr,
               Act* out_act = nullptr, int gn_G = 0, const QkvOut* qkv = nullptr, const uint8_t* premade_u0 = nullptr,
               const uint8_t* tw0_poly = nullptr) {
    p.B = bd.B;
    double flops = 0, bytes = 4.0 * p.B * (double)p.Tout * p.Cout;   // output write
    for (int s = 0; s < p.nseg; ++s) {
        const double cin = p.seg[s].C0 + p.seg[s].C1;
        flops += 2.0 * cin * p.Cout * p.seg[s].taps * (double)p.Tout * p.B;
        if (s == 0 || p.seg[s].src0 != p.seg[0].src0) bytes += 4.0 * p.B * (double)p.seg[s].Tin * cin;  // input read (once)
        bytes += 4.0 * cin * p.Cout * p.seg[s].taps;                                                     // weights
    }
    if (p.res) bytes += 4.0 * p.B * (double)p.res_Tin * p.Cout;
    // tensor-pipe path when the math mode asks for it and the shape is eligible; otherwise fp32 SIMT
    bool tc = bd.math != EEGLDM_MATH_FP32_SIMT && tw0 && (p.nseg == 1 || tw1) && !p.ddim_x && p.stride == 1 && p.Tc == p.Tout &&
              p.pad_left == p.seg[0].taps / 2 && (p.nseg == 1 || p.seg[1].taps == 1);
    for (int s = 0; tc && s < p.nseg; ++s)
        tc = conv_tc_eligible(p.seg[s].C0, p.seg[s].C1, p.Cout, p.Tout, p.seg[s].taps, 1);
    if (tc) {
        const bool x3 = bd.math == EEGLDM_MATH_F16X3_TC;
        TcConvParams q{};
        q.nseg = p.nseg; q.Cout = p.Cout; q.Tout = p.Tout; q.nsegs16 = (int)((long long)p.B * p.Tout / 16);
        int stages = 0;
        for (int s = 0; s < p.nseg; ++s) stages += (p.seg[s].C0 + p.seg[s].C1) / TC_BK * p.seg[s].taps;
        q.bn = conv_tc_bn(p.Cout, stages);   // tile width is a launch-time choice: the weight image is width-agnostic
        // "nearest x2 -> 3-tap conv" (in_layers of an up-sampling ResBlock) in polyphase form: the conv runs on the low-resolution
        // input with [even | odd] output phases as 2 Cout columns and two taps of MMAs per phase instead of three (TcConvParams.poly)
        const bool poly = tw0_poly && g_conv_poly && g_conv_gn_fine && p.nseg == 1 && p.seg[0].taps == 3 &&
                          p.seg[0].resample == RS_NEAREST2 && 2 * p.seg[0].Tin == p.Tout && p.seg[0].Tin % 16 == 0 && !p.res && !qkv &&
                          !premade_u0 && !(share && share->want_raw) && p.Cout % 256 == 0 && q.bn == 256;
        if (poly) {
            q.poly = 1; q.Cout = 2 * p.Cout; q.Tout = p.seg[0].Tin; q.nsegs16 = (int)((long long)p.B * q.Tout / 16);
            p.seg[0].resample = RS_NONE;   // (p is this function's copy: the tensor path below sees the low-resolution input as it is)
            tw0 = tw0_poly;
        }
        std::shared_ptr<Buf> ubuf[2];
        // fused producer: the conv kernel reads the fp32 sources itself (no act_split pass, no U tensors); AvgPool inputs
        // (the two down-sampling ResBlocks) keep the pre-pass
        // and 1x1 convs with many N tiles (qkv: 6) would redo the transform per N tile with only one tap of MMAs to hide it
        bool direct = g_conv_direct && !premade_u0 && (g_conv_direct_wide || !(p.seg[0].taps == 1 && p.Cout / q.bn > 2));
        for (int s = 0; direct && s < p.nseg; ++s) direct = p.seg[s].resample == RS_NONE || p.seg[s].resample == RS_NEAREST2;
        q.direct = direct ? 1 : 0;
        for (int s = 0; s < p.nseg; ++s) {
            // pre-pass: GroupNorm apply + SiLU + resample + 16-bit split -> tile images (one pass per conv input)
            const ConvSeg& a = p.seg[s];
            const int cin = a.C0 + a.C1;
            if (direct) {
                q.seg[s] = TcSeg{nullptr, s == 0 ? tw0 : tw1, a.taps, cin / TC_BK, a.src0, a.src1, a.C0, a.C1, a.scale, a.shift, a.silu,
                                 a.resample, a.Tin};
                continue;
            }
            const size_t ub = act_split_bytes(q.nsegs16, cin);
            if (s == 0 && premade_u0) {
                q.seg[s] = TcSeg{premade_u0, tw0, a.taps, cin / TC_BK};
                continue;
            }
            if (s == 1 && share && share->raw) {   // raw twin already written by conv1's pre-pass
                q.seg[s] = TcSeg{reinterpret_cast<const uint8_t*>(bd.ptr(share->raw)), tw1, a.taps, cin / TC_BK};
                continue;
            }
            ubuf[s] = bd.scratch((ub + 3) / 4);
            uint8_t* uraw = nullptr;
            if (s == 0 && share && share->want_raw) {
                share->raw = bd.scratch((ub + 3) / 4);
                uraw = reinterpret_cast<uint8_t*>(bd.ptr(share->raw));
            }
            ActSplitParams sp{a.src0, a.src1, a.C0, a.C1, a.scale, a.shift, a.silu, a.resample, a.Tin, q.Tout, q.nsegs16, cin / TC_BK,
                              reinterpret_cast<uint8_t*>(bd.ptr(ubuf[s])), uraw, bd.range_flag};
            bd.add([sp, x3](cudaStream_t st) { return launch_act_split(sp, x3, st); }, 1, OP_SPLIT, 0.0,
                   4.0 * p.B * (double)a.Tin * cin + (double)ub * (x3 ? 1.0 : 0.5) * (uraw ? 2.0 : 1.0));
            q.seg[s] = TcSeg{sp.U, s == 0 ? tw0 : tw1, a.taps, cin / TC_BK};
        }
        q.bias = p.bias; q.temb = p.temb; q.temb_stride = p.temb_stride; q.res = p.res; q.res_mode = p.res_mode; q.res_Tin = p.res_Tin;
        q.out = p.out;
        q.range_flag = bd.range_flag;
        if (qkv) { q.qkv16 = qkv->dst; q.qkv_H = qkv->H; q.qkv_ch = qkv->ch; }
        // record width: 4 channels (8 from 512 output channels) instead of the consumer's own group width, so that the same records
        // also serve the concat norms of the up path, whose groups are 12 / 24 channels wide and straddle the two sources (plan_gn)
        if (out_act && gn_G > 0 && g_conv_gn_fused && g_conv_gn_fine) {
            const int rec_G = p.Cout / (p.Cout >= 512 ? 8 : 4);
            if (rec_G % gn_G == 0 && conv_tc_gn_ok(p.Cout, rec_G)) gn_G = rec_G;
        }
        if (out_act && g_conv_gn_fused && conv_tc_gn_ok(p.Cout, gn_G)) {
            // one record per 128-row tile where a tile never straddles samples (kernel-side length % 128 == 0), else per segment
            q.gn_tile = g_conv_gn_tile && g_conv_tc_epi8 && q.Tout % 128 == 0 && p.Cout / gn_G != 32 ? 1 : 0;
            out_act->gn_nsplit = q.gn_tile ? p.Tout / 128 : p.Tout / 16; out_act->gn_G = gn_G;
            out_act->gn_part = bd.scratch((size_t)bd.B * out_act->gn_nsplit * gn_G * 3);
            q.gn_partial = bd.ptr(out_act->gn_part); q.gn_cpg = p.Cout / gn_G;
        }
        bd.add([q, x3](cudaStream_t st) { return launch_conv_tc(q, x3, st); }, 1, OP_CONV, flops, bytes);
        return;
    }
    // profile family: the narrow convs (1-channel in / out convs of the UNet, the 2-2-4 autoencoder) are HBM / latency-bound
    // SIMT launches and are kept apart from the GEMM-shaped ones the roofline is quoted on
    // the UNet's input conv (1 -> 128 channels) writes the GroupNorm records of its output like a tensor-pipe conv's epilogue does
    if (out_act && gn_G > 0 && g_conv_gn_fused && g_conv_gn_fine && conv_narrow_in_gn_ok(p) && (p.Cout / 4) % gn_G == 0) {
        p.gn_rec_tile = g_conv_gn_tile && p.Tout % 128 == 0 ? 1 : 0;
        out_act->gn_nsplit = p.gn_rec_tile ? p.Tout / 128 : p.Tout / 16; out_act->gn_G = p.Cout / 4;
        out_act->gn_part = bd.scratch((size_t)bd.B * out_act->gn_nsplit * out_act->gn_G * 3);
        p.gn_rec = bd.ptr(out_act->gn_part);
    }
    const bool narrow = p.seg[0].C0 + p.seg[0].C1 < 32 || p.Cout < 32;
    bd.add([p](cudaStream_t st) { return launch_conv_simt(p, st); }, 1, narrow ? OP_CONV_SIMT : OP_CONV, flops, bytes);
}

Act plan_res(Builder& bd, const ULayer& l, const Act& x0, const Act* x1, const UNetIO& io) {
    const int Tc = resampled_len(x0.T, l.mode);
    Act h1 = bd.act(l.cout, Tc);
    TcShare share;
    // the skip_connection's operand is produced by conv1's pre-pass when both convs take the tensor-pipe path
    share.want_raw = bd.math != EEGLDM_MATH_FP32_SIMT && l.cin != l.cout && l.t_w1 && l.t_w2 && l.t_wskip &&
                     conv_tc_eligible(x0.C, x1 ? x1->C : 0, l.cout, Tc, 3, 1);
    {
        ScaleShift ss = plan_gn(bd, x0, x1, 32, l.g1, l.be1, 1e-6f);
        ConvParams p{};
        p.seg[0] = make_seg(bd, x0, x1, &ss, 1, l.mode, l.w1, 3);
        p.nseg = 1; p.Cout = l.cout; p.Tout = Tc; p.Tc = Tc; p.stride = 1; p.pad_left = 1;
        p.bias = l.b1; p.temb = io.temb + l.emb_off; p.temb_stride = io.temb_stride;
        p.out = bd.wptr(h1);
        plan_conv(bd, p, l.t_w1, nullptr, &share, &h1, 32, nullptr, nullptr, l.t_w1p);
    }
    Act y = bd.act(l.cout, Tc);
    {
        ScaleShift ss = plan_gn(bd, h1, nullptr, 32, l.g2, l.be2, 1e-6f);
        ConvParams p{};
        p.seg[0] = make_seg(bd, h1, nullptr, &ss, 1, RS_NONE, l.w2, 3);
        p.nseg = 1; p.Cout = l.cout; p.Tout = Tc; p.Tc = Tc; p.stride = 1; p.pad_left = 1;
        p.bias = l.b2;
        if (l.cin != l.cout) {  // skip_connection = Conv1d k=1 on the (resampled) raw input, unet.py:295-302
            p.seg[1] = make_seg(bd, x0, x1, nullptr, 0, l.mode, l.wskip, 1);
            p.nseg = 2;
        } else {                // Identity skip (cin == cout implies a single source)
            p.res = bd.ptr(x0); p.res_mode = l.mode; p.res_Tin = x0.T;
        }
        p.out = bd.wptr(y);
        share.want_raw = false;
        plan_conv(bd, p, l.t_w2, l.t_wskip, &share, &y, 32);   // every consumer of a block output normalises with 32 groups
    }
    return y;
}

Act plan_attn(Builder& bd, const ULayer& l, const Act& x) {
    const int hch = l.ch / l.heads;
    const bool attn_tc = bd.math != EEGLDM_MATH_FP32_SIMT && attn_tc_eligible(x.T, hch);
    // f16x3: the qkv conv's epilogue writes the attention kernel's fp16 hi/lo operand images directly (no fp32 qkv tensor)
    const bool fuse_qkv = attn_tc && g_conv_qkv_fused && bd.math == EEGLDM_MATH_F16X3_TC && l.t_wqkv &&
                          conv_tc_eligible(x.C, 0, 3 * l.ch, x.T, 1, 1);
    // the attention kernel reads the fp32 qkv tensor itself and splits q, k, v in its producer warps (no qkv_split pass)
    const bool attn_direct = attn_tc && !fuse_qkv && g_attn_direct && attn_direct_eligible(x.T, hch);
    std::shared_ptr<Buf> q16;
    if (attn_tc && !attn_direct) q16 = bd.scratch((attn_qkv16_bytes(bd.B, x.T, l.heads, hch) + 3) / 4);
    Act qkv;
    if (!fuse_qkv) qkv = bd.act(3 * l.ch, x.T);
    {
        ScaleShift ss = plan_gn(bd, x, nullptr, 32, l.g1, l.be1, 1e-6f);
        ConvParams p{};
        p.seg[0] = make_seg(bd, x, nullptr, &ss, 0, RS_NONE, l.wqkv, 1);
        p.nseg = 1; p.Cout = 3 * l.ch; p.Tout = x.T; p.Tc = x.T; p.stride = 1; p.pad_left = 0;
        p.bias = l.bqkv; p.out = fuse_qkv ? nullptr : bd.wptr(qkv);
        QkvOut qo{fuse_qkv ? reinterpret_cast<uint8_t*>(bd.ptr(q16)) : nullptr, l.heads, hch};
        plan_conv(bd, p, l.t_wqkv, nullptr, nullptr, nullptr, 0, fuse_qkv ? &qo : nullptr);
    }
    // tensor-pipe attention followed by a tensor-pipe proj_out: the attention epilogue writes proj_out's fp16 hi/lo operand
    // image directly (no fp32 attention output, no pre-pass / producer for the 1x1 conv)
    // (sequences longer than one 256-key score tile run as key blocks merged into an fp32 output: no operand-image epilogue there)
    const bool attn_split = attn_tc && attn_tc_key_block(x.T) != x.T;
    const bool fuse_proj = attn_tc && !attn_split && g_attn_u_fused && l.t_wproj && x.T % 16 == 0 && conv_tc_eligible(l.ch, 0, l.ch, x.T, 1, 1);
    std::shared_ptr<Buf> au;
    Act a;
    if (fuse_proj) au = bd.scratch((act_split_bytes((int)((long long)bd.B * x.T / 16), l.ch) + 3) / 4);
    else a = bd.act(l.ch, x.T);
    if (attn_tc) {
        // tensor-pipe attention: q,k,v as fp16 hi/lo operand images, then S = QK^T -> softmax -> PV in one kernel
        const bool x3 = bd.math == EEGLDM_MATH_F16X3_TC;
        uint8_t* qdst = attn_direct ? nullptr : reinterpret_cast<uint8_t*>(bd.ptr(q16));
        const int B = bd.B, T = x.T, H = l.heads;
        if (!fuse_qkv && !attn_direct) {
            const float* qsrc = bd.ptr(qkv);
            int* rf = x3 ? bd.range_flag : nullptr;
            bd.add([=](cudaStream_t st) { return launch_qkv_split(qsrc, qdst, B, T, H, hch, st, rf); }, 1, OP_SPLIT, 0.0,
                   8.0 * B * (double)T * 3 * l.ch);
        }
        AttnTcParams tp{qdst, fuse_proj ? nullptr : bd.wptr(a), T, H, hch, B, 1.4426950408889634f / sqrtf((float)hch),
                        fuse_proj ? reinterpret_cast<uint8_t*>(bd.ptr(au)) : nullptr, attn_direct ? bd.ptr(qkv) : nullptr};
        std::shared_ptr<Buf> part;
        if (attn_split) {   // partial outputs + (max, sum) per key block
            part = bd.scratch((attn_tc_scratch_bytes(B, T, H, hch) + 3) / 4);
            tp.part_out = bd.ptr(part);
        }
        bd.add([tp, x3](cudaStream_t st) { return launch_attention_tc(tp, x3, st); }, attn_split ? 2 : 1, OP_ATTN,
               4.0 * B * (double)T * T * l.ch, 4.0 * B * (double)T * l.ch * 4.0);
    } else {
        AttnParams ap{};
        ap.qkv = bd.ptr(qkv); ap.out = bd.wptr(a); ap.T = x.T; ap.H = l.heads; ap.ch = l.ch / l.heads; ap.B = bd.B;
        bd.add([ap](cudaStream_t st) { return launch_attention_simt(ap, st); }, 1, OP_ATTN,
               4.0 * bd.B * (double)x.T * x.T * l.ch, 4.0 * bd.B * (double)x.T * l.ch * 4.0);
    }
    qkv.buf.reset();
    q16.reset();
    Act y = bd.act(l.ch, x.T);
    {
        ConvParams p{};
        if (fuse_proj) { Act ph; ph.ext = bd.ptr(x); ph.C = l.ch; ph.T = x.T; p.seg[0] = make_seg(bd, ph, nullptr, nullptr, 0, RS_NONE, l.wproj, 1); }
        else p.seg[0] = make_seg(bd, a, nullptr, nullptr, 0, RS_NONE, l.wproj, 1);   // (fused: the source pointer is a placeholder)
        p.nseg = 1; p.Cout = l.ch; p.Tout = x.T; p.Tc = x.T; p.stride = 1; p.pad_left = 0;
        p.bias = l.bproj; p.res = bd.ptr(x); p.res_mode = RS_NONE; p.res_Tin = x.T;
        p.out = bd.wptr(y);
        plan_conv(bd, p, l.t_wproj, nullptr, nullptr, &y, 32, nullptr, fuse_proj ? reinterpret_cast<const uint8_t*>(bd.ptr(au)) : nullptr);
    }
    return y;
}

Act plan_layer(Builder& bd, const ULayer& l, const Act& x0, const Act* x1, const UNetIO& io) {
    switch (l.kind) {
        case ULayer::CONV_IN: {
            Act y = bd.act(l.cout, x0.T);
            ConvParams p{};
            p.seg[0] = make_seg(bd, x0, nullptr, nullptr, 0, RS_NONE, l.w1, 3);
            p.nseg = 1; p.Cout = l.cout; p.Tout = x0.T; p.Tc = x0.T; p.stride = 1; p.pad_left = 1;
            p.bias = l.b1; p.out = bd.wptr(y);
            plan_conv(bd, p, nullptr, nullptr, nullptr, &y, 32);
            return y;
        }
        case ULayer::RES: return plan_res(bd, l, x0, x1, io);
        case ULayer::ATTN: return plan_attn(bd, l, x0);
        case ULayer::DOWN: {  // Downsample, unet.py:177-199
            if (l.use_conv) {
                const int Tout = (x0.T + 2 - 3) / 2 + 1;
                Act y = bd.act(l.ch, Tout);
                ConvParams p{};
                p.seg[0] = make_seg(bd, x0, nullptr, nullptr, 0, RS_NONE, l.w1, 3);
                p.nseg = 1; p.Cout = l.ch; p.Tout = Tout; p.Tc = x0.T; p.stride = 2; p.pad_left = 1;
                p.bias = l.b1; p.out = bd.wptr(y);
                plan_conv(bd, p);
                return y;
            }
            Act y = bd.act(l.ch, x0.T / 2);
            const float* src = bd.ptr(x0); float* dst = bd.wptr(y);
            const int B = bd.B, C = l.ch, Tin = x0.T;
            bd.add([=](cudaStream_t st) { return launch_resample(src, dst, B, Tin, C, RS_AVGPOOL2, st); }, 1);
            return y;
        }
        case ULayer::UP: {  // Upsample, unet.py:202-224
            if (l.use_conv) {
                Act y = bd.act(l.ch, x0.T * 2);
                ConvParams p{};
                p.seg[0] = make_seg(bd, x0, nullptr, nullptr, 0, RS_NEAREST2, l.w1, 3);
                p.nseg = 1; p.Cout = l.ch; p.Tout = x0.T * 2; p.Tc = x0.T * 2; p.stride = 1; p.pad_left = 1;
                p.bias = l.b1; p.out = bd.wptr(y);
                plan_conv(bd, p, l.t_w1);
                return y;
            }
            Act y = bd.act(l.ch, x0.T * 2);
            const float* src = bd.ptr(x0); float* dst = bd.wptr(y);
            const int B = bd.B, C = l.ch, Tin = x0.T;
            bd.add([=](cudaStream_t st) { return launch_resample(src, dst, B, Tin, C, RS_NEAREST2, st); }, 1);
            return y;
        }
    }
    return Act{};
}

int plan_unet_body(eegldm_unet* h, Builder& bd, int T, const UNetIO& io) {
    Act x; x.ext = io.x; x.C = h->cfg.in_channels; x.T = T;
    std::vector<Act> hs;
    Act cur = x;
    for (auto& blk : h->input_blocks) {
        for (auto& l : blk) cur = plan_layer(bd, l, cur, nullptr, io);
        hs.push_back(cur);
    }
    for (auto& l : h->middle) cur = plan_layer(bd, l, cur, nullptr, io);
    for (auto& blk : h->output_blocks) {
        Act skip = hs.back(); hs.pop_back();
        if (skip.T != cur.T) return fail(EEGLDM_ERR_SHAPE, "skip length mismatch: T must be divisible by 2^(levels-1)");
        bool first = true;
        for (auto& l : blk) {
            cur = plan_layer(bd, l, cur, first ? &skip : nullptr, io);
            if (first) skip.buf.reset();
            first = false;
        }
    }
    // out = Conv3(SiLU(GN(h)))  unet.py:501-505  (+ fused scheduler step)
    ScaleShift ss = plan_gn(bd, cur, nullptr, 32, h->out_g, h->out_be, 1e-6f);
    ConvParams p{};
    p.seg[0] = make_seg(bd, cur, nullptr, &ss, 1, RS_NONE, h->out_w, 3);
    p.nseg = 1; p.Cout = h->cfg.out_channels; p.Tout = T; p.Tc = T; p.stride = 1; p.pad_left = 1;
    p.bias = h->out_b; p.out = io.out; p.ddim_x = io.ddim_x; p.ddim_coef = io.ddim_coef;
    plan_conv(bd, p);
    return EEGLDM_OK;
}

// two passes: size the arena, then emit launches with real pointers
int build_unet_plan(eegldm_unet* h, int B, int T, const UNetIO& io, Builder& out) {
    const int levels = h->cfg.n_channel_mult;
    if (T <= 0 || (T % (1 << (levels - 1))) != 0)
        return fail(EEGLDM_ERR_SHAPE, "T must be a positive multiple of 2^(levels-1)");
    Builder sizing; sizing.B = B; sizing.math = h->math;
    int r = plan_unet_body(h, sizing, T, io);
    if (r) return r;
    if (sizing.peak > h->arena_cap) {
        h->drop_graphs();
        r = ensure(h->arena, h->arena_cap, sizing.peak);
        if (r) return r;
    }
    out.B = B; out.base = h->arena; out.math = h->math; out.range_flag = h->range_flag;
    return plan_unet_body(h, out, T, io);
}

// Two independent plans over the batch halves [0, B0) and [B0, B) with disjoint arena regions (B0 = ceil(B/2)).
// io describes the whole batch; per-sample tensors of the second half are offset by B0 rows.
int build_unet_plan_lanes(eegldm_unet* h, int B, int T, const UNetIO& io, Builder& lane0, Builder& lane1) {
    const int levels = h->cfg.n_channel_mult;
    if (T <= 0 || (T % (1 << (levels - 1))) != 0)
        return fail(EEGLDM_ERR_SHAPE, "T must be a positive multiple of 2^(levels-1)");
    const int B0 = (B + 1) / 2, B1 = B - B0;
    Builder sizing; sizing.B = B0; sizing.math = h->math;
    int r = plan_unet_body(h, sizing, T, io);
    if (r) return r;
    const size_t region = (sizing.peak + 63) & ~size_t(63);
    if (2 * region > h->arena_cap) {
        h->drop_graphs();
        r = ensure(h->arena, h->arena_cap, 2 * region);
        if (r) return r;
    }
    lane0.B = B0; lane0.base = h->arena; lane0.math = h->math; lane0.range_flag = h->range_flag;
    r = plan_unet_body(h, lane0, T, io);
    if (r) return r;
    UNetIO io1 = io;
    const size_t xo = (size_t)B0 * T * h->cfg.in_channels, oo = (size_t)B0 * T * h->cfg.out_channels;
    io1.x = io.x + xo; io1.out = io.out + oo;
    if (io.ddim_x) io1.ddim_x = io.ddim_x + oo;
    if (io.temb_stride) io1.temb = io.temb + (size_t)B0 * io.temb_stride;
    lane1.B = B1; lane1.base = h->arena + region; lane1.math = h->math; lane1.range_flag = h->range_flag;
    return plan_unet_body(h, lane1, T, io1);
}

struct ProfRec { OpMeta m; cudaEvent_t e0, e1; };
bool g_profile = false;
std::vector<ProfRec> g_prof;

int run_ops(const Builder& bd, cudaStream_t st) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    const bool prof = g_profile && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
    for (size_t i = 0; i < bd.ops.size(); ++i) {
        ProfRec r{};
        if (prof) {
            r.m = bd.meta[i];
            CU(cudaEventCreate(&r.e0)); CU(cudaEventCreate(&r.e1));
            CU(cudaEventRecord(r.e0, st));
        }
        cudaError_t e = bd.ops[i](st);
        if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
        if (prof) { CU(cudaEventRecord(r.e1, st)); g_prof.push_back(r); }
    }
    return EEGLDM_OK;
}

// one launch made outside a Builder plan (scheduler step bookkeeping, boundary scaling), recorded in the live profile as "other"
template <class F>
int run_profiled(int kind, double bytes, cudaStream_t st, F launch) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    const bool prof = g_profile && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
    ProfRec r{};
    if (prof) {
        r.m = OpMeta{kind, 0.0, bytes};
        CU(cudaEventCreate(&r.e0)); CU(cudaEventCreate(&r.e1));
        CU(cudaEventRecord(r.e0, st));
    }
    cudaError_t e = launch();
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    if (prof) { CU(cudaEventRecord(r.e1, st)); g_prof.push_back(r); }
    return EEGLDM_OK;
}

// timestep_embedding (unet.py:12-36) -- double precision on the host, rounded to fp32 at the same
// points the reference rounds (freqs, the product t*freq, the cos/sin result).
void timestep_embedding_host(const float* t, int nt, int dim, float* out) {
    const int half = dim / 2;
    for (int i = 0; i < nt; ++i) {
        for (int j = 0; j < half; ++j) {
            const float fr = (float)std::exp((double)(-(float)std::log(10000.0) * (float)j / (float)half));
            const float arg = t[i] * fr;
            out[(size_t)i * dim + j] = (float)std::cos((double)arg);
            out[(size_t)i * dim + half + j] = (float)std::sin((double)arg);
        }
        if (dim % 2) out[(size_t)i * dim + dim - 1] = 0.f;
    }
}

// time MLP + every ResBlock's emb projection: temb_out[nt][emb_total]   (unet.py:372-377, 277-285)
// t_host: host timesteps (embedding computed on the host), or null with t_dev: device timesteps (embedding kernel, no host
// round trip -- the reference's training loop passes a CUDA tensor, training.py:430)
int run_time_mlp(eegldm_unet* h, const float* t_host, int nt, float* temb_out, cudaStream_t st, const float* t_dev = nullptr) {
    const int mc = h->cfg.model_channels, ted = h->ted;
    int r = ensure(h->tscratch, h->tscratch_cap, (size_t)nt * (mc + 2 * ted));
    if (r) return r;
    float* d_te = h->tscratch;
    float* d_h0 = d_te + (size_t)nt * mc;
    float* d_emb = d_h0 + (size_t)nt * ted;
    if (t_host) {
        std::vector<float> te((size_t)nt * mc);
        timestep_embedding_host(t_host, nt, mc, te.data());
        CU(cudaMemcpyAsync(d_te, te.data(), te.size() * sizeof(float), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    } else CU(launch_timestep_embedding(t_dev, nt, mc, d_te, st));
    CU(launch_linear(d_te, h->te0_w, h->te0_b, d_h0, nt, mc, ted, 0, st));
    CU(launch_linear(d_h0, h->te2_w, h->te2_b, d_emb, nt, ted, ted, 1, st));
    CU(launch_linear(d_emb, h->emb_w, h->emb_b, temb_out, nt, ted, h->emb_total, 1, st));
    return EEGLDM_OK;
}

// ------------------------------------------------------------------------------------------------
// scheduler tables (host)   generative/networks/schedulers/{scheduler,ddim}.py [upstream]
int sched_alphas(const eegldm_sched_cfg* c, std::vector<float>& ac) {
    if (!c || c->num_train_timesteps < 1) return fail(EEGLDM_ERR_INVALID, "bad scheduler config");
    const int n = c->num_train_timesteps;
    ac.resize(n);
    // torch.linspace(fp32): start + i*step for the first half, end - (n-1-i)*step for the second
    auto linspace = [n](float a, float b, int i) -> float {
        if (n == 1) return a;
        const float step = (b - a) / (float)(n - 1);
        return i < n / 2 ? a + step * (float)i : b - step * (float)(n - 1 - i);
    };
    float prod = 1.f;
    for (int i = 0; i < n; ++i) {
        float beta;
        if (c->schedule == 0) beta = linspace(c->beta_start, c->beta_end, i);
        else if (c->schedule == 1) {
            const float s = linspace(std::sqrt(c->beta_start), std::sqrt(c->beta_end), i);
            beta = s * s;
        } else return fail(EEGLDM_ERR_INVALID, "schedule must be 0 (linear_beta) or 1 (scaled_linear_beta)");
        prod *= (1.f - beta);
        ac[i] = prod;
    }
    return EEGLDM_OK;
}

int sched_ddim_tables(const eegldm_sched_cfg* c, int n_steps, std::vector<int64_t>& ts, std::vector<float>& coef) {
    std::vector<float> ac;
    int r = sched_alphas(c, ac);
    if (r) return r;
    const int n = c->num_train_timesteps;
    if (n_steps < 1 || n_steps > n) return fail(EEGLDM_ERR_INVALID, "n_steps must be in [1, num_train_timesteps]");
    const int ratio = n / n_steps;
    ts.resize(n_steps); coef.resize(2 * (size_t)n_steps);
    for (int i = 0; i < n_steps; ++i) {
        const int t = (n_steps - 1 - i) * ratio + c->steps_offset;
        if (t < 0 || t >= n) return fail(EEGLDM_ERR_INVALID, "timestep out of range (steps_offset)");
        ts[i] = t;
        const int prev = t - ratio;
        const double a_t = ac[t];
        const double a_p = prev >= 0 ? (double)ac[prev] : (c->set_alpha_to_one ? 1.0 : (double)ac[0]);
        const double sa = std::sqrt(a_t), sb = std::sqrt(1.0 - a_t), sap = std::sqrt(a_p), sbp = std::sqrt(1.0 - a_p);
        double cx, cm;
        if (c->prediction_type == 1) { cx = sap * sa + sbp * sb; cm = -sap * sb + sbp * sa; }       // v-prediction
        else if (c->prediction_type == 0) { cx = sap / sa; cm = -sap * sb / sa + sbp; }             // epsilon
        else return fail(EEGLDM_ERR_INVALID, "prediction_type must be 0 (epsilon) or 1 (v_prediction)");
        coef[2 * i] = (float)cx; coef[2 * i + 1] = (float)cm;
    }
    return EEGLDM_OK;
}

}  // namespace

// ================================================================================================ AEKL
namespace {
struct ALayer {
    enum Kind { CONV, RES, DOWN, UP, NORM } kind;
    std::string prefix;
    int cin = 0, cout = 0, k = 3;
    const float *g1 = nullptr, *be1 = nullptr, *w1 = nullptr, *b1 = nullptr, *g2 = nullptr, *be2 = nullptr, *w2 = nullptr,
                *b2 = nullptr, *wskip = nullptr;
    size_t o_g1 = 0, o_be1 = 0, o_w1 = 0, o_b1 = 0, o_g2 = 0, o_be2 = 0, o_w2 = 0, o_b2 = 0, o_wskip = 0;
};
}  // namespace

// Training state of the autoencoder (eegldm_aekl_train_step): flat parameter / gradient / Adam-moment buffers in the
// engine's packed layouts, plus the map back to the reference's state_dict entries.
namespace {
struct TrainOff { size_t w1 = 0, b1 = 0, g1 = 0, be1 = 0, w2 = 0, b2 = 0, g2 = 0, be2 = 0, ws = 0, bs = 0; };
struct TrainEntry { std::string name; size_t off; std::vector<int64_t> shape; bool conv; };
struct AeklPending;                      // a recorded forward pass waiting for eegldm_aekl_backward (defined with TrainRun)
void aekl_pending_free(AeklPending* p);
struct AeklTrain {
    std::vector<TrainOff> enc, dec;
    TrainOff qmu, qls, pq;
    std::vector<TrainEntry> entries;
    size_t n = 0;
    float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr, *losses = nullptr, *arena = nullptr;
    size_t arena_cap = 0;
    int step = 0;
    bool dirty = false;   // P has moved away from the host state_dict / inference weights
    AeklPending* pending = nullptr;
    ~AeklTrain() {
        if (pending) aekl_pending_free(pending);
        for (float* p : {P, G, M, V, losses, arena}) if (p) cudaFree(p);
    }
};
}  // namespace

struct eegldm_aekl {
    AeklTrain* train = nullptr;
    eegldm_aekl_cfg cfg{};
    ParamSet ps;
    std::vector<ALayer> enc, dec;
    ALayer q_mu, q_ls, post_q;
    WeightPool pool;
    bool finalized = false;
    float* arena = nullptr; size_t arena_cap = 0;
    float* io_tmp = nullptr; size_t io_tmp_cap = 0;
    int down_factor() const { return 1 << (cfg.n_levels - 1); }
    ~eegldm_aekl() { if (arena) cudaFree(arena); if (io_tmp) cudaFree(io_tmp); delete train; }
};

namespace {

int build_aekl_topology(eegldm_aekl* h) {
    const auto& c = h->cfg;
    if (c.n_levels < 1 || c.n_levels > 8) return fail(EEGLDM_ERR_INVALID, "bad n_levels");
    if (c.in_channels < 1 || c.out_channels < 1 || c.latent_channels < 1 || c.norm_num_groups < 1)
        return fail(EEGLDM_ERR_INVALID, "bad autoencoder config");
    for (int i = 0; i < c.n_levels; ++i)
        if (c.num_channels[i] < 1 || c.num_channels[i] % c.norm_num_groups || c.num_res_blocks[i] < 0)
            return fail(EEGLDM_ERR_INVALID, "num_channels must be positive multiples of norm_num_groups");
    auto mk = [](ALayer::Kind k, const std::string& p, int cin, int cout, int ks) {
        ALayer l; l.kind = k; l.prefix = p; l.cin = cin; l.cout = cout; l.k = ks; return l;
    };
    const int L = c.n_levels, z = c.latent_channels;
    auto& enc = h->enc; auto& dec = h->dec;
    enc.push_back(mk(ALayer::CONV, "encoder.blocks.0", c.in_channels, c.num_channels[0], 3));
    int out_ch = c.num_channels[0];
    for (int i = 0; i < L; ++i) {
        int in_ch = out_ch; out_ch = c.num_channels[i];
        for (int r = 0; r < c.num_res_blocks[i]; ++r) {
            enc.push_back(mk(ALayer::RES, "encoder.blocks." + std::to_string(enc.size()), in_ch, out_ch, 3));
            in_ch = out_ch;
        }
        if (i != L - 1) enc.push_back(mk(ALayer::DOWN, "encoder.blocks." + std::to_string(enc.size()), in_ch, in_ch, 3));
        out_ch = in_ch;
    }
    enc.push_back(mk(ALayer::NORM, "encoder.blocks." + std::to_string(enc.size()), out_ch, out_ch, 0));
    enc.push_back(mk(ALayer::CONV, "encoder.blocks." + std::to_string(enc.size()), out_ch, z, 3));

    std::vector<int> rc(c.num_channels, c.num_channels + L), rr(c.num_res_blocks, c.num_res_blocks + L);
    std::reverse(rc.begin(), rc.end()); std::reverse(rr.begin(), rr.end());
    dec.push_back(mk(ALayer::CONV, "decoder.blocks.0", z, rc[0], 3));
    out_ch = rc[0];
    int in_ch = out_ch;
    for (int i = 0; i < L; ++i) {
        in_ch = out_ch; out_ch = rc[i];
        for (int r = 0; r < rr[i]; ++r) {
            dec.push_back(mk(ALayer::RES, "decoder.blocks." + std::to_string(dec.size()), in_ch, out_ch, 3));
            in_ch = out_ch;
        }
        if (i != L - 1) dec.push_back(mk(ALayer::UP, "decoder.blocks." + std::to_string(dec.size()), in_ch, in_ch, 3));
        out_ch = in_ch;
    }
    dec.push_back(mk(ALayer::NORM, "decoder.blocks." + std::to_string(dec.size()), in_ch, in_ch, 0));
    dec.push_back(mk(ALayer::CONV, "decoder.blocks." + std::to_string(dec.size()), in_ch, c.out_channels, 3));
    h->q_mu = mk(ALayer::CONV, "quant_conv_mu", z, z, 1);
    h->q_ls = mk(ALayer::CONV, "quant_conv_log_sigma", z, z, 1);
    h->post_q = mk(ALayer::CONV, "post_quant_conv", z, z, 1);

    ParamSet& ps = h->ps;
    auto conv = [&](const std::string& p, int i, int o, int k) { ps.add(p + ".conv.weight", {o, i, k}); ps.add(p + ".conv.bias", {o}); };
    auto reg = [&](const ALayer& l) {
        const std::string& p = l.prefix;
        switch (l.kind) {
            case ALayer::CONV: conv(p, l.cin, l.cout, l.k); break;
            case ALayer::RES:
                ps.gn(p + ".norm1", l.cin); conv(p + ".conv1", l.cin, l.cout, 3);
                ps.gn(p + ".norm2", l.cout); conv(p + ".conv2", l.cout, l.cout, 3);
                if (l.cin != l.cout) conv(p + ".nin_shortcut", l.cin, l.cout, 1);
                break;
            case ALayer::DOWN: case ALayer::UP: conv(p + ".conv", l.cin, l.cin, 3); break;
            case ALayer::NORM: ps.gn(p, l.cin); break;
        }
    };
    for (auto& l : enc) reg(l);
    for (auto& l : dec) reg(l);
    reg(h->q_mu); reg(h->q_ls); reg(h->post_q);
    return EEGLDM_OK;
}

int finalize_aekl(eegldm_aekl* h) {
    int r = h->ps.check_all_loaded();
    if (r) return r;
    WeightPool& wp = h->pool;
    wp.stage.clear();
    const ParamSet& ps = h->ps;
    auto stage = [&](ALayer& l) {
        const std::string& p = l.prefix;
        switch (l.kind) {
            case ALayer::CONV:
                l.o_w1 = wp.push(pack_conv(ps.get(p + ".conv.weight"), l.cout, l.cin, l.k)); l.o_b1 = wp.push(ps.get(p + ".conv.bias"));
                break;
            case ALayer::RES: {
                l.o_g1 = wp.push(ps.get(p + ".norm1.weight")); l.o_be1 = wp.push(ps.get(p + ".norm1.bias"));
                l.o_w1 = wp.push(pack_conv(ps.get(p + ".conv1.conv.weight"), l.cout, l.cin, 3)); l.o_b1 = wp.push(ps.get(p + ".conv1.conv.bias"));
                l.o_g2 = wp.push(ps.get(p + ".norm2.weight")); l.o_be2 = wp.push(ps.get(p + ".norm2.bias"));
                l.o_w2 = wp.push(pack_conv(ps.get(p + ".conv2.conv.weight"), l.cout, l.cout, 3));
                std::vector<float> b2 = ps.get(p + ".conv2.conv.bias");
                if (l.cin != l.cout) {
                    l.o_wskip = wp.push(pack_conv(ps.get(p + ".nin_shortcut.conv.weight"), l.cout, l.cin, 1));
                    const auto& bs = ps.get(p + ".nin_shortcut.conv.bias");
                    for (int i = 0; i < l.cout; ++i) b2[i] += bs[i];
                }
                l.o_b2 = wp.push(b2);
                break;
            }
            case ALayer::DOWN: case ALayer::UP:
                l.o_w1 = wp.push(pack_conv(ps.get(p + ".conv.conv.weight"), l.cin, l.cin, 3)); l.o_b1 = wp.push(ps.get(p + ".conv.conv.bias"));
                break;
            case ALayer::NORM:
                l.o_g1 = wp.push(ps.get(p + ".weight")); l.o_be1 = wp.push(ps.get(p + ".bias"));
                break;
        }
    };
    for (auto& l : h->enc) stage(l);
    for (auto& l : h->dec) stage(l);
    stage(h->q_mu); stage(h->q_ls); stage(h->post_q);
    r = wp.upload();
    if (r) return r;
    auto fix = [&](ALayer& l) {
        l.g1 = wp.at(l.o_g1); l.be1 = wp.at(l.o_be1); l.w1 = wp.at(l.o_w1); l.b1 = wp.at(l.o_b1);
        l.g2 = wp.at(l.o_g2); l.be2 = wp.at(l.o_be2); l.w2 = wp.at(l.o_w2); l.b2 = wp.at(l.o_b2);
        l.wskip = (l.kind == ALayer::RES && l.cin != l.cout) ? wp.at(l.o_wskip) : nullptr;
    };
    for (auto& l : h->enc) fix(l);
    for (auto& l : h->dec) fix(l);
    fix(h->q_mu); fix(h->q_ls); fix(h->post_q);
    h->finalized = true;
    return EEGLDM_OK;
}

Act plan_aekl_conv(Builder& bd, const ALayer& l, const Act& x, const ScaleShift* ss, int resample, int stride, int pad_left,
                   float* ext_out = nullptr) {
    const int Tc = resampled_len(x.T, resample);
    const int Tout = stride == 1 ? Tc : (Tc + 1 - 3) / 2 + 1;  // stride 2: pad right 1, k3, padding 0
    Act y;
    if (ext_out) { y.ext = ext_out; y.C = l.cout; y.T = Tout; }
    else y = bd.act(l.cout, Tout);
    ConvParams p{};
    p.seg[0] = make_seg(bd, x, nullptr, ss, 0, resample, l.w1, l.k);
    p.nseg = 1; p.Cout = l.cout; p.Tout = Tout; p.Tc = Tc; p.stride = stride; p.pad_left = pad_left;
    p.bias = l.b1; p.out = ext_out ? ext_out : bd.wptr(y);
    plan_conv(bd, p);
    return y;
}

// generic block runner (Encoder.forward / Decoder.forward upstream; ae_kl.py:66-80 ResBlock)
Act plan_aekl_blocks(Builder& bd, const std::vector<ALayer>& blocks, Act cur, int G, float* final_out) {
    ScaleShift pending{};  // the final GroupNorm is folded into the last conv's prologue (no SiLU)
    bool has_pending = false;
    for (size_t i = 0; i < blocks.size(); ++i) {
        const ALayer& l = blocks[i];
        const bool last = i + 1 == blocks.size();
        switch (l.kind) {
            case ALayer::CONV:
                cur = plan_aekl_conv(bd, l, cur, has_pending ? &pending : nullptr, RS_NONE, 1, l.k / 2, last ? final_out : nullptr);
                has_pending = false; pending = ScaleShift{};
                break;
            case ALayer::RES: {
                Act h1 = bd.act(l.cout, cur.T);
                {
                    ScaleShift ss = plan_gn(bd, cur, nullptr, G, l.g1, l.be1, 1e-6f);
                    ConvParams p{};
                    p.seg[0] = make_seg(bd, cur, nullptr, &ss, 1, RS_NONE, l.w1, 3);
                    p.nseg = 1; p.Cout = l.cout; p.Tout = cur.T; p.Tc = cur.T; p.stride = 1; p.pad_left = 1;
                    p.bias = l.b1; p.out = bd.wptr(h1);
                    plan_conv(bd, p);
                }
                Act y = bd.act(l.cout, cur.T);
                {
                    ScaleShift ss = plan_gn(bd, h1, nullptr, G, l.g2, l.be2, 1e-6f);
                    ConvParams p{};
                    p.seg[0] = make_seg(bd, h1, nullptr, &ss, 1, RS_NONE, l.w2, 3);
                    p.nseg = 1; p.Cout = l.cout; p.Tout = cur.T; p.Tc = cur.T; p.stride = 1; p.pad_left = 1;
                    p.bias = l.b2;
                    if (l.cin != l.cout) { p.seg[1] = make_seg(bd, cur, nullptr, nullptr, 0, RS_NONE, l.wskip, 1); p.nseg = 2; }
                    else { p.res = bd.ptr(cur); p.res_mode = RS_NONE; p.res_Tin = cur.T; }
                    p.out = bd.wptr(y);
                    plan_conv(bd, p);
                }
                cur = y;
                break;
            }
            case ALayer::DOWN: cur = plan_aekl_conv(bd, l, cur, nullptr, RS_NONE, 2, 0); break;     // ae_kl.py:41-45
            case ALayer::UP: cur = plan_aekl_conv(bd, l, cur, nullptr, RS_NEAREST2, 1, 1); break;   // ae_kl.py:27-30
            case ALayer::NORM:
                pending = plan_gn(bd, cur, nullptr, G, l.g1, l.be1, 1e-6f);
                has_pending = true;
                break;
        }
    }
    return cur;
}

struct AeklEncodeIO { const float* x; float* mu; float* sigma; };

int plan_aekl_encode(eegldm_aekl* h, Builder& bd, int L, const AeklEncodeIO& io) {
    Act x; x.ext = io.x; x.C = h->cfg.in_channels; x.T = L;
    Act hh = plan_aekl_blocks(bd, h->enc, x, h->cfg.norm_num_groups, nullptr);
    plan_aekl_conv(bd, h->q_mu, hh, nullptr, RS_NONE, 1, 0, io.mu);
    Act ls = plan_aekl_conv(bd, h->q_ls, hh, nullptr, RS_NONE, 1, 0);
    const float* lsp = bd.ptr(ls); float* sg = io.sigma;
    const size_t n = (size_t)bd.B * ls.T * ls.C;
    bd.add([=](cudaStream_t st) { return launch_kl_sigma(lsp, sg, n, st); }, 1);
    return EEGLDM_OK;
}

int plan_aekl_decode(eegldm_aekl* h, Builder& bd, int T, const float* z, float* out) {
    Act x; x.ext = z; x.C = h->cfg.latent_channels; x.T = T;
    Act p = plan_aekl_conv(bd, h->post_q, x, nullptr, RS_NONE, 1, 0);
    plan_aekl_blocks(bd, h->dec, p, h->cfg.norm_num_groups, out);
    return EEGLDM_OK;
}

template <class PlanFn>
int build_aekl_plan(eegldm_aekl* h, int B, PlanFn fn, Builder& out) {
    Builder sizing; sizing.B = B;
    int r = fn(sizing);
    if (r) return r;
    r = ensure(h->arena, h->arena_cap, std::max<size_t>(sizing.peak, 64));
    if (r) return r;
    out.B = B; out.base = h->arena;
    return fn(out);
}

// boundary layout helpers: reference NCL <-> engine NLC (identity when C == 1)
int to_nlc(const float* src_ncl, float*& tmp, size_t& cap, size_t tmp_off, int B, int C, int T, cudaStream_t st,
           const float** out) {
    if (C == 1) { *out = src_ncl; return EEGLDM_OK; }
    (void)cap;
    CU(launch_transpose_ncl_to_nlc(src_ncl, tmp + tmp_off, B, C, T, st));
    *out = tmp + tmp_off;
    return EEGLDM_OK;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

const char* eegldm_last_error(void) { return g_err.c_str(); }
const char* eegldm_version(void) { return "eegldm 0.1 (sm_100a)"; }
int64_t eegldm_launch_count(void) { return (int64_t)g_launch_count.load(); }
int eegldm_set_graphs(int enabled) { g_graphs_enabled = enabled != 0; return EEGLDM_OK; }
int eegldm_set_sample_lanes(int lanes) {
    if (lanes != 1 && lanes != 2) return fail(EEGLDM_ERR_INVALID, "lanes must be 1 or 2");
    g_sample_lanes = lanes;
    return EEGLDM_OK;
}

int eegldm_set_conv_cluster(int ctas) {
    if (ctas != 1 && ctas != 2 && ctas != 4) return fail(EEGLDM_ERR_INVALID, "cluster size must be 1, 2 or 4");
    g_conv_tc_cluster = ctas;
    return EEGLDM_OK;
}

int eegldm_set_conv_tuning(int pair, int bn256_min_stages, int fuse_epilogues) {
    if (pair < 0 || pair > 3) return fail(EEGLDM_ERR_INVALID, "pair must be a bit mask in 0..3 (bit 0: 256-wide launches, bit 1: 128-wide launches)");
    g_conv_gn_fused = (fuse_epilogues & 1) != 0;
    g_conv_qkv_fused = (fuse_epilogues & 2) != 0;
    g_conv_direct = (fuse_epilogues & 4) != 0;
    g_attn_u_fused = (fuse_epilogues & 8) != 0;
    g_attn_direct = (fuse_epilogues & 16) != 0;
    g_conv_direct_wide = (fuse_epilogues & 128) != 0;
    g_conv_gn_fine = (fuse_epilogues & 256) == 0;
    g_conv_poly = (fuse_epilogues & 512) == 0;
    g_conv_gn_tile = (fuse_epilogues & 1024) == 0;
    g_conv_tc_epi8 = (fuse_epilogues & 64) ? 0 : 1;  // bit 6 switches the two-warpgroup conv epilogue OFF (A/B timing)
    g_conv_tc_cat = (fuse_epilogues & 32) ? 0 : 1;   // bit 5 switches the concatenated hi|lo MMA of the N = 128 tiles OFF (A/B timing)
    if (bn256_min_stages < 1) return fail(EEGLDM_ERR_INVALID, "bn256_min_stages must be >= 1");
    g_conv_tc_pair = pair;
    g_conv_tc_bn256_stages = bn256_min_stages;
    return EEGLDM_OK;
}

int eegldm_profile_enable(int on) {
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_profile = on != 0;
    return EEGLDM_OK;
}
int eegldm_profile_record(int i, int* kind, double* ms, double* flops, double* bytes) {
    if (i < 0 || i >= (int)g_prof.size()) return fail(EEGLDM_ERR_INVALID, "record index out of range");
    ProfRec& r = g_prof[i];
    CU(cudaEventSynchronize(r.e1));
    float dt = 0.f;
    CU(cudaEventElapsedTime(&dt, r.e0, r.e1));
    if (kind) *kind = r.m.kind; if (ms) *ms = dt; if (flops) *flops = r.m.flops; if (bytes) *bytes = r.m.bytes;
    return EEGLDM_OK;
}
int eegldm_profile_read(int kind, double* ms, double* flops, double* bytes, int64_t* launches) {
    if (kind < 0 || kind >= OP_NKIND) return fail(EEGLDM_ERR_INVALID, "kind must be 0 (conv), 1 (groupnorm), 2 (attention), 3 (other), 4 (activation split) or 5 (narrow fp32 conv)");
    double t = 0, f = 0, b = 0; int64_t n = 0;
    for (auto& r : g_prof) {
        if (r.m.kind != kind) continue;
        CU(cudaEventSynchronize(r.e1));
        float dt = 0.f;
        CU(cudaEventElapsedTime(&dt, r.e0, r.e1));
        t += dt; f += r.m.flops; b += r.m.bytes; ++n;
    }
    if (ms) *ms = t; if (flops) *flops = f; if (bytes) *bytes = b; if (launches) *launches = n;
    return EEGLDM_OK;
}

int eegldm_unet_create(const eegldm_unet_cfg* cfg, eegldm_unet** out) {
    if (!cfg || !out) return fail(EEGLDM_ERR_INVALID, "null argument");
    auto* h = new (std::nothrow) eegldm_unet();
    if (!h) return fail(EEGLDM_ERR_NOMEM, "out of host memory");
    h->cfg = *cfg;
    int r = build_unet_topology(h);
    if (r) { delete h; return r; }
    *out = h;
    return EEGLDM_OK;
}
void eegldm_unet_destroy(eegldm_unet* h) { delete h; }
int eegldm_unet_num_params(const eegldm_unet* h) { return h ? (int)h->ps.params.size() : 0; }
int eegldm_unet_param_info(const eegldm_unet* h, int i, const char** name, int64_t shape[4], int* ndim) {
    if (!h || i < 0 || i >= (int)h->ps.params.size()) return fail(EEGLDM_ERR_INVALID, "bad parameter index");
    const HostParam& p = h->ps.params[i];
    if (name) *name = p.name.c_str();
    if (ndim) *ndim = (int)p.shape.size();
    if (shape) for (size_t k = 0; k < p.shape.size() && k < 4; ++k) shape[k] = p.shape[k];
    return EEGLDM_OK;
}
int eegldm_unet_load(eegldm_unet* h, const char* name, const float* host, const int64_t* shape, int ndim) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    h->finalized = false;
    if (h->train) { unet_train_free(h->train); h->train = nullptr; }   // new weights: optimiser state starts over
    return h->ps.load(name, host, shape, ndim);
}
int eegldm_unet_finalize(eegldm_unet* h) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    return finalize_unet(h);
}
int eegldm_unet_set_math(eegldm_unet* h, eegldm_math mode) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    if (mode != EEGLDM_MATH_FP32_SIMT && mode != EEGLDM_MATH_F16X3_TC && mode != EEGLDM_MATH_BF16_TC)
        return fail(EEGLDM_ERR_INVALID, "unknown math mode");
    if (mode == h->math) return EEGLDM_OK;
    h->drop_graphs();
    h->math = mode;
    if (h->finalized) return finalize_unet(h);   // weight images depend on the math mode
    return EEGLDM_OK;
}

static int unet_forward_impl(eegldm_unet* h, const float* x_dev, const float* timesteps_host, const float* timesteps_dev, int nt,
                             float* out_dev, int B, int T, void* stream) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    if (!h->finalized) return fail(EEGLDM_ERR_MISSING, "eegldm_unet_finalize has not been called");
    if (B < 0 || (nt != 1 && nt != B)) return fail(EEGLDM_ERR_SHAPE, "timesteps must have 1 or B entries");
    if (B == 0) return EEGLDM_OK;  // empty batch: nothing to do (pointers may be null)
    if (!x_dev || (!timesteps_host && !timesteps_dev) || !out_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int zin = h->cfg.in_channels, zout = h->cfg.out_channels;
    int r = ensure(h->temb_fwd, h->temb_fwd_cap, (size_t)nt * h->emb_total);
    if (r) return r;
    r = run_time_mlp(h, timesteps_host, nt, h->temb_fwd, st, timesteps_dev);
    if (r) return r;
    const size_t nin = (size_t)B * T * zin, nout = (size_t)B * T * zout;
    r = ensure(h->xtmp, h->xtmp_cap, nin + nout);
    if (r) return r;
    UNetIO io{};
    r = to_nlc(x_dev, h->xtmp, h->xtmp_cap, 0, B, zin, T, st, &io.x);
    if (r) return r;
    io.out = zout == 1 ? out_dev : h->xtmp + nin;
    io.temb = h->temb_fwd; io.temb_stride = nt == 1 ? 0 : h->emb_total;
    Builder bd;
    r = build_unet_plan(h, B, T, io, bd);
    if (r) return r;
    r = run_ops(bd, st);
    if (r) return r;
    if (zout != 1) CU(launch_transpose_nlc_to_ncl(io.out, out_dev, B, zout, T, st));
    return EEGLDM_OK;
}

int eegldm_unet_forward(eegldm_unet* h, const float* x_dev, const float* timesteps_host, int nt, float* out_dev, int B,
                        int T, void* stream) {
    if (!timesteps_host && B > 0) return fail(EEGLDM_ERR_INVALID, "null argument");
    return unet_forward_impl(h, x_dev, timesteps_host, nullptr, nt, out_dev, B, T, stream);
}
int eegldm_unet_forward_devt(eegldm_unet* h, const float* x_dev, const float* timesteps_dev, int nt, float* out_dev, int B,
                             int T, void* stream) {
    if (!timesteps_dev && B > 0) return fail(EEGLDM_ERR_INVALID, "null argument");
    return unet_forward_impl(h, x_dev, nullptr, timesteps_dev, nt, out_dev, B, T, stream);
}

int eegldm_unet_range_status(eegldm_unet* h, int* overflow_out) {
    if (!h || !overflow_out) return fail(EEGLDM_ERR_INVALID, "null argument");
    *overflow_out = 0;
    if (!h->range_flag) return EEGLDM_OK;
    int v = 0;
    CU(cudaMemcpy(&v, h->range_flag, sizeof(int), cudaMemcpyDeviceToHost));   // synchronises with the work that may raise it
    if (v) CU(cudaMemset(h->range_flag, 0, sizeof(int)));
    *overflow_out = v != 0;
    return EEGLDM_OK;
}

// ---------------------------------------------------------------------------------------------- AEKL
int eegldm_aekl_create(const eegldm_aekl_cfg* cfg, eegldm_aekl** out) {
    if (!cfg || !out) return fail(EEGLDM_ERR_INVALID, "null argument");
    auto* h = new (std::nothrow) eegldm_aekl();
    if (!h) return fail(EEGLDM_ERR_NOMEM, "out of host memory");
    h->cfg = *cfg;
    int r = build_aekl_topology(h);
    if (r) { delete h; return r; }
    *out = h;
    return EEGLDM_OK;
}
void eegldm_aekl_destroy(eegldm_aekl* h) { delete h; }
int eegldm_aekl_num_params(const eegldm_aekl* h) { return h ? (int)h->ps.params.size() : 0; }
int eegldm_aekl_param_info(const eegldm_aekl* h, int i, const char** name, int64_t shape[4], int* ndim) {
    if (!h || i < 0 || i >= (int)h->ps.params.size()) return fail(EEGLDM_ERR_INVALID, "bad parameter index");
    const HostParam& p = h->ps.params[i];
    if (name) *name = p.name.c_str();
    if (ndim) *ndim = (int)p.shape.size();
    if (shape) for (size_t k = 0; k < p.shape.size() && k < 4; ++k) shape[k] = p.shape[k];
    return EEGLDM_OK;
}
int eegldm_aekl_load(eegldm_aekl* h, const char* name, const float* host, const int64_t* shape, int ndim) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    h->finalized = false;
    if (h->train) { delete h->train; h->train = nullptr; }   // new weights: optimiser state starts over
    return h->ps.load(name, host, shape, ndim);
}
int eegldm_aekl_finalize(eegldm_aekl* h) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    return finalize_aekl(h);
}

int eegldm_aekl_encode(eegldm_aekl* h, const float* x_dev, float* z_mu_dev, float* z_sigma_dev, int B, int L, void* stream) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    if (!h->finalized) return fail(EEGLDM_ERR_MISSING, "eegldm_aekl_finalize has not been called");
    const int f = h->down_factor();
    if (B < 0 || L <= 0 || L % f) return fail(EEGLDM_ERR_SHAPE, "L must be a positive multiple of 2^(levels-1)");
    if (B == 0) return EEGLDM_OK;
    if (!x_dev || !z_mu_dev || !z_sigma_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int cin = h->cfg.in_channels, z = h->cfg.latent_channels, T = L / f;
    const size_t nin = (size_t)B * L * cin, nz = (size_t)B * T * z;
    int r = ensure(h->io_tmp, h->io_tmp_cap, nin + 2 * nz);
    if (r) return r;
    AeklEncodeIO io{};
    r = to_nlc(x_dev, h->io_tmp, h->io_tmp_cap, 0, B, cin, L, st, &io.x);
    if (r) return r;
    io.mu = z == 1 ? z_mu_dev : h->io_tmp + nin;
    io.sigma = z == 1 ? z_sigma_dev : h->io_tmp + nin + nz;
    Builder bd;
    r = build_aekl_plan(h, B, [&](Builder& b) { return plan_aekl_encode(h, b, L, io); }, bd);
    if (r) return r;
    r = run_ops(bd, st);
    if (r) return r;
    if (z != 1) {
        CU(launch_transpose_nlc_to_ncl(io.mu, z_mu_dev, B, z, T, st));
        CU(launch_transpose_nlc_to_ncl(io.sigma, z_sigma_dev, B, z, T, st));
    }
    return EEGLDM_OK;
}

int eegldm_aekl_decode(eegldm_aekl* h, const float* z_dev, float* out_dev, int B, int T, void* stream) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    if (!h->finalized) return fail(EEGLDM_ERR_MISSING, "eegldm_aekl_finalize has not been called");
    if (B < 0 || T <= 0) return fail(EEGLDM_ERR_SHAPE, "bad latent shape");
    if (B == 0) return EEGLDM_OK;
    if (!z_dev || !out_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int z = h->cfg.latent_channels, co = h->cfg.out_channels, L = T * h->down_factor();
    const size_t nz = (size_t)B * T * z, nout = (size_t)B * L * co;
    int r = ensure(h->io_tmp, h->io_tmp_cap, nz + nout);
    if (r) return r;
    const float* zin;
    r = to_nlc(z_dev, h->io_tmp, h->io_tmp_cap, 0, B, z, T, st, &zin);
    if (r) return r;
    float* out = co == 1 ? out_dev : h->io_tmp + nz;
    Builder bd;
    r = build_aekl_plan(h, B, [&](Builder& b) { return plan_aekl_decode(h, b, T, zin, out); }, bd);
    if (r) return r;
    r = run_ops(bd, st);
    if (r) return r;
    if (co != 1) CU(launch_transpose_nlc_to_ncl(out, out_dev, B, co, L, st));
    return EEGLDM_OK;
}

int eegldm_aekl_forward(eegldm_aekl* h, const float* x_dev, const float* eps_dev, float* recon_dev, float* z_mu_dev,
                        float* z_sigma_dev, int B, int L, void* stream) {
    if (!h || !eps_dev || !recon_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    int r = eegldm_aekl_encode(h, x_dev, z_mu_dev, z_sigma_dev, B, L, stream);
    if (r || B == 0) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = L / h->down_factor();
    const size_t nz = (size_t)B * T * h->cfg.latent_channels;
    float* zbuf = nullptr;
    CU(cudaMallocAsync((void**)&zbuf, nz * sizeof(float), st));
    cudaError_t e = launch_axpy_sampling(z_mu_dev, z_sigma_dev, eps_dev, zbuf, nz, st);  // NCL elementwise
    if (e == cudaSuccess) {
        r = eegldm_aekl_decode(h, zbuf, recon_dev, B, T, stream);
    } else r = cuda_fail(e, "sampling kernel");
    cudaFreeAsync(zbuf, st);
    return r;
}

// ---------------------------------------------------------------------------------------------- scheduler
int eegldm_sched_alphas_cumprod(const eegldm_sched_cfg* cfg, float* out) {
    if (!out) return fail(EEGLDM_ERR_INVALID, "null argument");
    std::vector<float> ac;
    int r = sched_alphas(cfg, ac);
    if (r) return r;
    std::memcpy(out, ac.data(), ac.size() * sizeof(float));
    return EEGLDM_OK;
}
int eegldm_sched_ddim_tables(const eegldm_sched_cfg* cfg, int n_steps, int64_t* timesteps, float* coef) {
    std::vector<int64_t> ts; std::vector<float> cf;
    int r = sched_ddim_tables(cfg, n_steps, ts, cf);
    if (r) return r;
    if (timesteps) std::memcpy(timesteps, ts.data(), ts.size() * sizeof(int64_t));
    if (coef) std::memcpy(coef, cf.data(), cf.size() * sizeof(float));
    return EEGLDM_OK;
}
int eegldm_timestep_embedding(const float* timesteps, int nt, int dim, float* out) {
    if (!timesteps || !out || nt < 0 || dim < 1) return fail(EEGLDM_ERR_INVALID, "bad argument");
    timestep_embedding_host(timesteps, nt, dim, out);
    return EEGLDM_OK;
}

// ---------------------------------------------------------------------------------------------- sampling
int eegldm_ddim_sample(eegldm_unet* u, eegldm_aekl* a, const eegldm_sched_cfg* sc, const float* noise_dev, float scale_factor,
                       int n_steps, float* out_dev, int B, int T, void* stream) {
    if (!u || !sc) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (B > 0 && (!noise_dev || !out_dev)) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (!u->finalized) return fail(EEGLDM_ERR_MISSING, "eegldm_unet_finalize has not been called");
    if (a && !a->finalized) return fail(EEGLDM_ERR_MISSING, "eegldm_aekl_finalize has not been called");
    if (u->cfg.in_channels != u->cfg.out_channels) return fail(EEGLDM_ERR_INVALID, "sampling needs in_channels == out_channels");
    if (a && a->cfg.latent_channels != u->cfg.in_channels) return fail(EEGLDM_ERR_INVALID, "latent_channels mismatch");
    if (B < 0) return fail(EEGLDM_ERR_SHAPE, "negative batch");
    if (scale_factor == 0.f) return fail(EEGLDM_ERR_INVALID, "scale_factor must be non-zero");
    if (B == 0) return EEGLDM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int z = u->cfg.in_channels;
    const size_t nz = (size_t)B * T * z;

    // per-step tables (timestep embedding rows through the time MLP + DDIM coefficients), cached
    std::vector<int64_t> ts; std::vector<float> coef;
    int r = sched_ddim_tables(sc, n_steps, ts, coef);
    if (r) return r;
    std::vector<float> key{(float)sc->num_train_timesteps, sc->beta_start, sc->beta_end, (float)sc->schedule,
                           (float)sc->prediction_type, (float)sc->set_alpha_to_one, (float)sc->steps_offset, (float)n_steps};
    if (key != u->table_key) {
        // The captured step graphs bake the table pointers into their step_advance node: size both tables for the longest
        // schedule this scheduler can ask for (num_train_timesteps rows), and drop the graphs if they still have to move.
        const size_t rows = (size_t)std::max(n_steps, sc->num_train_timesteps);
        if (rows * u->emb_total > u->temb_table_cap || 2 * rows > u->coef_table_cap) u->drop_graphs();
        r = ensure(u->temb_table, u->temb_table_cap, rows * u->emb_total);
        if (r) return r;
        r = ensure(u->coef_table, u->coef_table_cap, 2 * rows);
        if (r) return r;
        std::vector<float> tf(ts.begin(), ts.end());
        r = run_time_mlp(u, tf.data(), n_steps, u->temb_table, st);
        if (r) return r;
        CU(cudaMemcpyAsync(u->coef_table, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));  // coef (host vector) must outlive the copy; tables are built once
        u->table_key = key;
    }
    if (nz > u->xbuf_cap) u->drop_graphs();
    r = ensure(u->xbuf, u->xbuf_cap, nz);
    if (r) return r;
    r = ensure(u->xtmp, u->xtmp_cap, 2 * nz);
    if (r) return r;
    // x_T = noise (sample_trials.py:151), engine layout
    if (z == 1) CU(cudaMemcpyAsync(u->xbuf, noise_dev, nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else CU(launch_transpose_ncl_to_nlc(noise_dev, u->xbuf, B, z, T, st));
    CU(cudaMemsetAsync(u->step_ctr, 0, sizeof(int), st));

    UNetIO io{};
    io.x = u->xbuf; io.out = u->xbuf; io.temb = u->temb_step; io.temb_stride = 0;
    io.ddim_x = u->xbuf; io.ddim_coef = u->coef_cur;
    auto emit_step = [&](const Builder& bd, cudaStream_t s) -> int {
        int rr = run_profiled(OP_OTHER, 8.0 * u->emb_total, s, [&] {
            return launch_step_advance(u->temb_table, u->emb_total, u->temb_step, u->coef_table, u->coef_cur, u->step_ctr, s); });
        if (rr) return rr;
        return run_ops(bd, s);
    };
    if (g_graphs_enabled) {
        auto key2 = std::make_pair(B, T);
        auto it = u->graphs.find(key2);
        if (it == u->graphs.end()) {
            const bool two = g_sample_lanes == 2 && B >= 2;
            Builder bd, bd1;
            r = two ? build_unet_plan_lanes(u, B, T, io, bd, bd1) : build_unet_plan(u, B, T, io, bd);   // may grow the arena (and drop stale graphs)
            if (r) return r;
            GraphEntry ge;
            const long long before = g_launch_count.load();
            // capture on a private stream: the caller's stream may be the legacy default stream, which cannot capture
            if (!u->cap_stream) CU(cudaStreamCreateWithFlags(&u->cap_stream, cudaStreamNonBlocking));
            if (two && !u->cap_stream2) {
                CU(cudaStreamCreateWithFlags(&u->cap_stream2, cudaStreamNonBlocking));
                CU(cudaEventCreateWithFlags(&u->ev_fork, cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&u->ev_join, cudaEventDisableTiming));
            }
            CU(cudaStreamBeginCapture(u->cap_stream, cudaStreamCaptureModeThreadLocal));
            cudaError_t e = launch_step_advance(u->temb_table, u->emb_total, u->temb_step, u->coef_table, u->coef_cur, u->step_ctr,
                                                u->cap_stream);
            if (e != cudaSuccess) r = cuda_fail(e, "step_advance");
            bool forked = false;
            if (!r && two) {   // fork: the second half's chain depends on step_advance only
                e = cudaEventRecord(u->ev_fork, u->cap_stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(u->cap_stream2, u->ev_fork, 0);
                if (e != cudaSuccess) r = cuda_fail(e, "graph fork");
                else forked = true;
            }
            if (!r) r = run_ops(bd, u->cap_stream);
            if (forked) {
                if (!r) r = run_ops(bd1, u->cap_stream2);
                e = cudaEventRecord(u->ev_join, u->cap_stream2);   // join (also on the error path: a capture cannot end forked)
                if (e == cudaSuccess) e = cudaStreamWaitEvent(u->cap_stream, u->ev_join, 0);
                if (e != cudaSuccess && !r) r = cuda_fail(e, "graph join");
            }
            e = cudaStreamEndCapture(u->cap_stream, &ge.graph);
            if (r) { if (ge.graph) cudaGraphDestroy(ge.graph); return r; }
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
            ge.n_kernels = (int)(g_launch_count.load() - before);
            g_launch_count.store(before);  // capture did not launch anything
            e = cudaGraphInstantiate(&ge.exec, ge.graph, 0);
            if (e != cudaSuccess) { cudaGraphDestroy(ge.graph); return cuda_fail(e, "cudaGraphInstantiate"); }
            it = u->graphs.emplace(key2, ge).first;
        }
        for (int s = 0; s < n_steps; ++s) {
            CU(cudaGraphLaunch(it->second.exec, st));
            g_launch_count += it->second.n_kernels;
        }
    } else {
        Builder bd;
        r = build_unet_plan(u, B, T, io, bd);
        if (r) return r;
        for (int s = 0; s < n_steps; ++s) { r = emit_step(bd, st); if (r) return r; }
    }
    if (!a) {
        if (z == 1) CU(cudaMemcpyAsync(out_dev, u->xbuf, nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
        else CU(launch_transpose_nlc_to_ncl(u->xbuf, out_dev, B, z, T, st));
        return EEGLDM_OK;
    }
    // decode_stage_2_outputs(latent / scale_factor)   sample_trials.py:166
    float* zs = u->xtmp;  // NCL staging for the decoder's boundary
    if (z == 1) { r = run_profiled(OP_OTHER, 8.0 * nz, st, [&] { return launch_scale(u->xbuf, zs, 1.0f / scale_factor, nz, st); }); if (r) return r; }
    else {
        CU(launch_transpose_nlc_to_ncl(u->xbuf, zs + nz, B, z, T, st));
        CU(launch_scale(zs + nz, zs, 1.0f / scale_factor, nz, st));
    }
    return eegldm_aekl_decode(a, zs, out_dev, B, T, stream);
}

int eegldm_ddim_sample_host(eegldm_unet* u, eegldm_aekl* a, const eegldm_sched_cfg* sc, const float* noise_host,
                            float scale_factor, int n_steps, float* out_host, int B, int T, void* stream) {
    if (!u || !noise_host || !out_host) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (B <= 0) return B == 0 ? EEGLDM_OK : fail(EEGLDM_ERR_SHAPE, "negative batch");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nz = (size_t)B * T * u->cfg.in_channels;
    const size_t nout = a ? (size_t)B * T * a->down_factor() * a->cfg.out_channels : nz;
    // device staging owned by the handle (grown on demand, reused): the stream-ordered allocator's trim / re-map at every
    // synchronisation showed up as occasional 0.2 - 1 s stalls of this call
    int r = ensure(u->host_in, u->host_in_cap, nz);
    if (r) return r;
    r = ensure(u->host_out, u->host_out_cap, nout);
    if (r) return r;
    cudaError_t e = cudaMemcpyAsync(u->host_in, noise_host, nz * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) r = cuda_fail(e, "H2D copy");
    if (!r) r = eegldm_ddim_sample(u, a, sc, u->host_in, scale_factor, n_steps, u->host_out, B, T, stream);
    if (!r) {
        e = cudaMemcpyAsync(out_host, u->host_out, nout * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) r = cuda_fail(e, "D2H copy");
    }
    e = cudaStreamSynchronize(st);
    if (!r && e != cudaSuccess) r = cuda_fail(e, "cudaStreamSynchronize");
    if (!r && u->math == EEGLDM_MATH_F16X3_TC) {
        int over = 0;
        r = eegldm_unet_range_status(u, &over);
        if (!r && over)
            r = fail(EEGLDM_ERR_INVALID, "f16x3: an activation left the fp16 operand range (|x| >= 65504 or NaN); the result is invalid -- "
                                         "use EEGLDM_MATH_FP32_SIMT for this model");
    }
    return r;
}

// ---------------------------------------------------------------------------------------------- test hook
// One fused convolution launch in a chosen math mode (tests/test_gpu_conv.py compares the tcgen05 kernel with
// the SIMT kernel and with torch.nn.functional.conv1d layer by layer).
int eegldm_test_conv(const float* x_dev, const float* scale_dev, const float* shift_dev, int silu, int resample,
                     const float* w_host, const float* bias_host, const float* res_dev, int B, int Tin, int Cin, int Cout, int k,
                     int math, float* out_dev, void* stream) {
    if (!x_dev || !w_host || !out_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (k != 1 && k != 3) return fail(EEGLDM_ERR_INVALID, "k must be 1 or 3");
    cudaStream_t st = (cudaStream_t)stream;
    const int Tc = resampled_len(Tin, resample);
    std::vector<float> w(w_host, w_host + (size_t)Cout * Cin * k);
    WeightPool wp;
    const size_t o_w = wp.push(pack_conv(w, Cout, Cin, k));
    size_t o_b = 0, o_t = 0;
    if (bias_host) o_b = wp.push(bias_host, Cout);
    const bool tc = math != EEGLDM_MATH_FP32_SIMT;
    // the polyphase form plan_conv picks for "nearest x2 -> 3-tap conv" (up-sampling ResBlocks)
    const bool poly = tc && g_conv_poly && resample == RS_NEAREST2 && k == 3 && !res_dev && Cout % 256 == 0 && Tin % 16 == 0 &&
                      conv_tc_bn(Cout, Cin / TC_BK * k) == 256;
    if (tc) {
        if (!conv_tc_eligible(Cin, 0, Cout, Tc, k, 1)) return fail(EEGLDM_ERR_SHAPE, "shape not eligible for the tcgen05 path");
        std::vector<uint16_t> img;
        if (poly) pack_conv_tc_poly(w.data(), Cout, Cin, math == EEGLDM_MATH_F16X3_TC, img);
        else pack_conv_tc(w.data(), Cout, Cin, k, math == EEGLDM_MATH_F16X3_TC, img);
        o_t = wp.push_u16(img);
    }
    int r = wp.upload();
    if (r) return r;
    ConvParams p{};
    p.seg[0] = ConvSeg{x_dev, nullptr, Cin, 0, scale_dev, shift_dev, silu, resample, Tin, wp.at(o_w), k};
    p.nseg = 1; p.Cout = Cout; p.Tout = Tc; p.Tc = Tc; p.stride = 1; p.pad_left = k / 2;
    p.bias = bias_host ? wp.at(o_b) : nullptr;
    p.res = res_dev; p.res_mode = RS_NONE; p.res_Tin = Tc;
    p.out = out_dev; p.B = B;
    cudaError_t ce;
    uint8_t* U = nullptr;
    if (tc) {
        const bool x3 = math == EEGLDM_MATH_F16X3_TC;
        TcConvParams q{};
        q.nseg = 1; q.Cout = Cout; q.Tout = Tc; q.nsegs16 = (int)((long long)B * Tc / 16);
        q.bn = conv_tc_bn(Cout, Cin / TC_BK * k);
        if (poly) { q.poly = 1; q.Cout = 2 * Cout; q.Tout = Tin; q.nsegs16 = (int)((long long)B * Tin / 16); resample = RS_NONE; }
        if (g_conv_direct && (resample == RS_NONE || resample == RS_NEAREST2)) {   // fused producer: no pre-pass
            q.direct = 1;
            q.seg[0] = TcSeg{nullptr, reinterpret_cast<const uint8_t*>(wp.at(o_t)), k, Cin / TC_BK, x_dev, nullptr, Cin, 0, scale_dev, shift_dev,
                             silu, resample, Tin};
            ce = cudaSuccess;
        } else {
            CU(cudaMalloc((void**)&U, act_split_bytes(q.nsegs16, Cin)));
            ActSplitParams sp{x_dev, nullptr, Cin, 0, scale_dev, shift_dev, silu, resample, Tin, q.Tout, q.nsegs16, Cin / TC_BK, U, nullptr};
            ce = launch_act_split(sp, x3, st);
            q.seg[0] = TcSeg{U, reinterpret_cast<const uint8_t*>(wp.at(o_t)), k, Cin / TC_BK};
        }
        q.bias = p.bias; q.res = res_dev; q.res_mode = RS_NONE; q.res_Tin = Tc; q.out = out_dev;
        if (ce == cudaSuccess) ce = launch_conv_tc(q, x3, st);
    } else ce = launch_conv_simt(p, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);   // the temporary weight pool / U are freed on return
    if (U) cudaFree(U);
    if (ce != cudaSuccess) return cuda_fail(ce, "conv launch");
    return EEGLDM_OK;
}

// tcgen05 convolution (no prologue, bias only) whose epilogue also emits the GroupNorm(G) statistics of its output;
// gn_finalize turns them into mean / rstd [B][G] (tests/test_gpu_conv.py compares both with torch on the same output).
int eegldm_test_conv_gn(const float* x_dev, const float* w_host, const float* bias_host, int B, int T, int Cin, int Cout, int k,
                        int G, float* out_dev, float* mean_dev, float* rstd_dev, void* stream) {
    if (!x_dev || !w_host || !out_dev || !mean_dev || !rstd_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (!conv_tc_eligible(Cin, 0, Cout, T, k, 1) || !conv_tc_gn_ok(Cout, G)) return fail(EEGLDM_ERR_SHAPE, "shape not eligible");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<uint16_t> img;
    pack_conv_tc(w_host, Cout, Cin, k, true, img);
    WeightPool wp;
    const size_t o_t = wp.push_u16(img);
    const size_t o_b = bias_host ? wp.push(bias_host, Cout) : 0;
    const size_t o_one = wp.push(std::vector<float>((size_t)Cout, 1.0f)), o_zero = wp.push(std::vector<float>((size_t)Cout, 0.0f));
    int r = wp.upload();
    if (r) return r;
    TcConvParams q{};
    q.nseg = 1; q.Cout = Cout; q.Tout = T; q.nsegs16 = (int)((long long)B * T / 16);
    q.bn = conv_tc_bn(Cout, Cin / TC_BK * k);
    uint8_t* U = nullptr;
    float *part = nullptr, *ss = nullptr;
    const int nsplit = T / 16;
    cudaError_t ce = cudaMalloc((void**)&U, act_split_bytes(q.nsegs16, Cin));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&part, (size_t)B * nsplit * G * 3 * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&ss, (size_t)B * Cout * 2 * sizeof(float));
    if (ce == cudaSuccess) {
        ActSplitParams sp{x_dev, nullptr, Cin, 0, nullptr, nullptr, 0, RS_NONE, T, T, q.nsegs16, Cin / TC_BK, U, nullptr};
        ce = launch_act_split(sp, true, st);
    }
    q.seg[0] = TcSeg{U, reinterpret_cast<const uint8_t*>(wp.at(o_t)), k, Cin / TC_BK};
    q.bias = bias_host ? wp.at(o_b) : nullptr; q.out = out_dev; q.gn_partial = part; q.gn_cpg = Cout / G;
    if (ce == cudaSuccess) ce = launch_conv_tc(q, true, st);
    GnParams g{};
    g.C0 = Cout; g.T = T; g.G = G; g.gamma = wp.at(o_one); g.beta = wp.at(o_zero); g.eps = 1e-6f; g.B = B;
    g.scale = ss; g.shift = ss ? ss + (size_t)B * Cout : nullptr; g.partial = part; g.nsplit = nsplit;
    g.mean_out = mean_dev; g.rstd_out = rstd_dev;
    if (ce == cudaSuccess) ce = launch_groupnorm_finalize(g, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(U); cudaFree(part); cudaFree(ss);
    if (ce != cudaSuccess) return cuda_fail(ce, "conv + GroupNorm statistics");
    return EEGLDM_OK;
}

// qkv 1x1 convolution writing attention operand images from its epilogue, followed by the tcgen05 attention kernel
// (the fused path of plan_attn): x [B][T][C] fp32, w [3C][C][1], out [B][T][C]; C = H*ch.
int eegldm_test_qkv_attention(const float* x_dev, const float* w_host, const float* bias_host, int B, int T, int H, int ch,
                              float* out_dev, void* stream) {
    if (!x_dev || !w_host || !out_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    const int Cc = H * ch;
    if (!conv_tc_eligible(Cc, 0, 3 * Cc, T, 1, 1) || !attn_tc_eligible(T, ch)) return fail(EEGLDM_ERR_SHAPE, "shape not eligible");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<uint16_t> img;
    pack_conv_tc(w_host, 3 * Cc, Cc, 1, true, img);
    WeightPool wp;
    const size_t o_t = wp.push_u16(img);
    const size_t o_b = bias_host ? wp.push(bias_host, 3 * Cc) : 0;
    int r = wp.upload();
    if (r) return r;
    TcConvParams q{};
    q.nseg = 1; q.Cout = 3 * Cc; q.Tout = T; q.nsegs16 = (int)((long long)B * T / 16);
    q.bn = conv_tc_bn(3 * Cc, Cc / TC_BK);
    uint8_t *U = nullptr, *q16 = nullptr;
    cudaError_t ce = cudaMalloc((void**)&U, act_split_bytes(q.nsegs16, Cc));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&q16, attn_qkv16_bytes(B, T, H, ch));
    if (ce == cudaSuccess) {
        ActSplitParams sp{x_dev, nullptr, Cc, 0, nullptr, nullptr, 0, RS_NONE, T, T, q.nsegs16, Cc / TC_BK, U, nullptr};
        ce = launch_act_split(sp, true, st);
    }
    q.seg[0] = TcSeg{U, reinterpret_cast<const uint8_t*>(wp.at(o_t)), 1, Cc / TC_BK};
    q.bias = bias_host ? wp.at(o_b) : nullptr; q.qkv16 = q16; q.qkv_H = H; q.qkv_ch = ch;
    if (ce == cudaSuccess) ce = launch_conv_tc(q, true, st);
    AttnTcParams tp{q16, out_dev, T, H, ch, B, 1.4426950408889634f / sqrtf((float)ch)};
    if (ce == cudaSuccess) ce = launch_attention_tc(tp, true, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(U); cudaFree(q16);
    if (ce != cudaSuccess) return cuda_fail(ce, "qkv conv + attention");
    return EEGLDM_OK;
}

// Timing of one tcgen05 convolution launch on synthetic data (tools/conv_bench.py): `reps` back-to-back launches
// bracketed by CUDA events on `stream`; debug: 0 = the real kernel, 1 = no operand copies, 2 = no MMAs (experiments that
// separate the tensor-pipe time from the operand-staging time; their outputs are garbage).  Inputs are pseudo-random.
__global__ void bench_fill_kernel(float* x, size_t n, uint32_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        x[i] = (float)(h & 0xFFFF) * (2.0f / 65535.0f) - 1.0f;
    }
}
static int bench_conv_impl(int B, int T, int Cin, int Cout, int k, int with_res, int math, int debug, int reps, float* ms_out,
                           double* timeline_out, void* stream) {
    if (!ms_out || reps <= 0) return fail(EEGLDM_ERR_INVALID, "bad argument");
    if (math == EEGLDM_MATH_FP32_SIMT || !conv_tc_eligible(Cin, 0, Cout, T, k, 1))
        return fail(EEGLDM_ERR_SHAPE, "shape / math not eligible for the tcgen05 path");
    cudaStream_t st = (cudaStream_t)stream;
    const bool x3 = math == EEGLDM_MATH_F16X3_TC;
    std::vector<float> w((size_t)Cout * Cin * k);
    uint32_t h = 12345u;
    for (auto& v : w) { h = h * 1664525u + 1013904223u; v = ((float)(h >> 8) / 8388608.0f - 1.0f) / sqrtf((float)Cin * k); }
    std::vector<uint16_t> img;
    pack_conv_tc(w.data(), Cout, Cin, k, x3, img);
    WeightPool wp;
    const size_t o_t = wp.push_u16(img);
    int r = wp.upload();
    if (r) return r;
    TcConvParams q{};
    q.nseg = 1; q.Cout = Cout; q.Tout = T; q.nsegs16 = (int)((long long)B * T / 16);
    q.bn = conv_tc_bn(Cout, Cin / TC_BK * k);
    // with_res == 2: the qkv conv of an AttentionBlock (one head of Cout/3 channels) writing attention operand images (pre-pass form)
    const bool qkv_mode = with_res == 2 && x3 && Cout % 3 == 0 && attn_tc_eligible(T, Cout / 3);
    if (with_res == 2 && !qkv_mode) return fail(EEGLDM_ERR_SHAPE, "qkv bench mode needs f16x3 and an attention-eligible shape");
    float *x = nullptr, *out = nullptr, *res = nullptr;
    uint8_t* U = nullptr;
    const size_t nx = (size_t)B * T * Cin, no = (size_t)B * T * Cout;
    cudaError_t ce = cudaMalloc((void**)&x, nx * 4);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&out, no * 4);
    uint8_t* q16 = nullptr;
    if (ce == cudaSuccess && with_res == 1) ce = cudaMalloc((void**)&res, no * 4);
    if (ce == cudaSuccess && qkv_mode) ce = cudaMalloc((void**)&q16, attn_qkv16_bytes(B, T, 1, Cout / 3));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&U, act_split_bytes(q.nsegs16, Cin));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ce == cudaSuccess) {
        bench_fill_kernel<<<1024, 256, 0, st>>>(x, nx, 1u);
        if (res) bench_fill_kernel<<<1024, 256, 0, st>>>(res, no, 2u);
        ActSplitParams sp{x, nullptr, Cin, 0, nullptr, nullptr, 0, RS_NONE, T, T, q.nsegs16, Cin / TC_BK, U, nullptr};
        ce = launch_act_split(sp, x3, st);
    }
    q.seg[0] = TcSeg{U, reinterpret_cast<const uint8_t*>(wp.at(o_t)), k, Cin / TC_BK};
    float* ss = nullptr;
    if (qkv_mode) { q.qkv16 = q16; q.qkv_H = 1; q.qkv_ch = Cout / 3; }
    if (g_conv_direct && !qkv_mode) {   // fused producer with a GroupNorm affine + SiLU prologue (the ResBlock conv1 / conv2 shape)
        if (ce == cudaSuccess) ce = cudaMalloc((void**)&ss, (size_t)B * Cin * 2 * 4);
        if (ce == cudaSuccess) { bench_fill_kernel<<<64, 256, 0, st>>>(ss, (size_t)B * Cin * 2, 3u); ce = cudaGetLastError(); }
        q.direct = 1;
        q.seg[0] = TcSeg{nullptr, reinterpret_cast<const uint8_t*>(wp.at(o_t)), k, Cin / TC_BK, x, nullptr, Cin, 0,
                         (debug & 128) ? nullptr : ss, (debug & 128) ? nullptr : ss + (size_t)B * Cin, (debug & 128) ? 0 : 1, RS_NONE, T};
    }
    q.res = res; q.res_mode = RS_NONE; q.res_Tin = T; q.out = out; q.debug = debug;
    if (ce == cudaSuccess) ce = launch_conv_tc(q, x3, st);   // warm-up
    if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
    if (ce == cudaSuccess) ce = cudaEventRecord(e0, st);
    for (int i = 0; i < reps && ce == cudaSuccess; ++i) ce = launch_conv_tc(q, x3, st);
    if (ce == cudaSuccess) ce = cudaEventRecord(e1, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    float ms = 0.f;
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / reps;
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (timeline_out && ce == cudaSuccess) {
        // one more launch with the per-CTA cycle counters switched on; averaged over the CTAs that ran
        const int max_ctas = 1024;
        unsigned long long* tl = nullptr;
        ce = cudaMalloc((void**)&tl, (size_t)max_ctas * TC_TL_N * sizeof(unsigned long long));
        if (ce == cudaSuccess) ce = cudaMemsetAsync(tl, 0, (size_t)max_ctas * TC_TL_N * sizeof(unsigned long long), st);
        q.timeline = tl;
        if (ce == cudaSuccess) ce = launch_conv_tc(q, x3, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        std::vector<unsigned long long> hst((size_t)max_ctas * TC_TL_N);
        if (ce == cudaSuccess) ce = cudaMemcpy(hst.data(), tl, hst.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        for (int j = 0; j < TC_TL_N; ++j) timeline_out[j] = 0.0;
        int n = 0;
        for (int c = 0; c < max_ctas; ++c)
            if (hst[(size_t)c * TC_TL_N + TC_TL_TOTAL]) {
                ++n;
                for (int j = 0; j < TC_TL_N; ++j) timeline_out[j] += (double)hst[(size_t)c * TC_TL_N + j];
            }
        for (int j = 0; j < TC_TL_N; ++j) timeline_out[j] /= std::max(n, 1);
        timeline_out[TC_TL_N - 1] = n;
        cudaFree(tl);
    }
    cudaFree(x); cudaFree(out); cudaFree(res); cudaFree(U); cudaFree(ss); cudaFree(q16);
    if (ce != cudaSuccess) return cuda_fail(ce, "conv bench");
    return EEGLDM_OK;
}
int eegldm_bench_conv(int B, int T, int Cin, int Cout, int k, int with_res, int math, int debug, int reps, float* ms_out,
                      void* stream) {
    return bench_conv_impl(B, T, Cin, Cout, k, with_res, math, debug, reps, ms_out, nullptr, stream);
}
int eegldm_bench_conv_timeline(int B, int T, int Cin, int Cout, int k, int with_res, int math, int debug, int reps, float* ms_out,
                               double* timeline_out, void* stream) {
    if (!timeline_out) return fail(EEGLDM_ERR_INVALID, "null argument");
    return bench_conv_impl(B, T, Cin, Cout, k, with_res, math, debug, reps, ms_out, timeline_out, stream);
}

// Timing of the tcgen05 attention kernel on synthetic operand images (tools/attn_timeline.py): `reps` launches bracketed by CUDA
// events, then one launch with per-CTA cycle stamps; timeline_out[8] = averages over the CTAs {1: S phase (start -> scores complete),
// 2: softmax, 3: PV + epilogue, 4: total, 7: CTAs}.
int eegldm_bench_attention(int B, int T, int H, int ch, int reps, float* ms_out, double* timeline_out, void* stream) {
    if (!ms_out || reps <= 0) return fail(EEGLDM_ERR_INVALID, "bad argument");
    if (!attn_tc_eligible(T, ch)) return fail(EEGLDM_ERR_SHAPE, "shape not eligible for the tcgen05 attention");
    cudaStream_t st = (cudaStream_t)stream;
    float* qkv = nullptr; uint8_t *q16 = nullptr, *ou = nullptr; unsigned long long* tl = nullptr;
    const size_t nq = (size_t)B * T * H * 3 * ch;
    const size_t nctas = (size_t)B * H * ((T + 127) / 128) * (T / std::max(attn_tc_key_block(T), 1));
    cudaError_t ce = cudaMalloc((void**)&qkv, nq * 4);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&q16, attn_qkv16_bytes(B, T, H, ch));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&ou, act_split_bytes((int)((long long)B * T / 16), H * ch));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&tl, nctas * 8 * sizeof(unsigned long long));
    if (ce == cudaSuccess) { bench_fill_kernel<<<1024, 256, 0, st>>>(qkv, nq, 7u); ce = launch_qkv_split(qkv, q16, B, T, H, ch, st); }
    AttnTcParams tp{q16, nullptr, T, H, ch, B, 1.4426950408889634f / sqrtf((float)ch), ou, nullptr, nullptr};
    float *part = nullptr, *out32 = nullptr;
    if (const size_t sb = attn_tc_scratch_bytes(B, T, H, ch)) {   // T > 256: key blocks merged into an fp32 output
        if (ce == cudaSuccess) ce = cudaMalloc((void**)&part, sb);
        if (ce == cudaSuccess) ce = cudaMalloc((void**)&out32, (size_t)B * T * H * ch * sizeof(float));
        tp.out = out32; tp.out_u = nullptr; tp.part_out = part;
    }
    if (ce == cudaSuccess) ce = launch_attention_tc(tp, true, st);   // warm-up
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
    if (ce == cudaSuccess) ce = cudaEventRecord(e0, st);
    for (int i = 0; i < reps && ce == cudaSuccess; ++i) ce = launch_attention_tc(tp, true, st);
    if (ce == cudaSuccess) ce = cudaEventRecord(e1, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    float ms = 0.f;
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / reps;
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (timeline_out && ce == cudaSuccess) {
        ce = cudaMemsetAsync(tl, 0, nctas * 8 * sizeof(unsigned long long), st);
        tp.timeline = tl;
        if (ce == cudaSuccess) ce = launch_attention_tc(tp, true, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        std::vector<unsigned long long> hst(nctas * 8);
        if (ce == cudaSuccess) ce = cudaMemcpy(hst.data(), tl, hst.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        for (int j = 0; j < 8; ++j) timeline_out[j] = 0.0;
        for (size_t c = 0; c < nctas; ++c)
            for (int j = 1; j < 5; ++j) timeline_out[j] += (double)hst[c * 8 + j];
        for (int j = 1; j < 5; ++j) timeline_out[j] /= (double)nctas;
        timeline_out[7] = (double)nctas;
    }
    cudaFree(qkv); cudaFree(q16); cudaFree(ou); cudaFree(tl); cudaFree(part); cudaFree(out32);
    if (ce != cudaSuccess) return cuda_fail(ce, "attention bench");
    return EEGLDM_OK;
}

// One attention launch (QKVAttentionLegacy.forward) on channels-last qkv [B][T][H*3*ch] -> out [B][T][H*ch].
int eegldm_test_attention(const float* qkv_dev, int B, int T, int H, int ch, int math, float* out_dev, void* stream) {
    if (!qkv_dev || !out_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce;
    uint8_t* q16 = nullptr;
    float* part = nullptr;
    if (math != EEGLDM_MATH_FP32_SIMT) {
        if (!attn_tc_eligible(T, ch)) return fail(EEGLDM_ERR_SHAPE, "shape not eligible for the tcgen05 attention");
        if (g_attn_direct && attn_direct_eligible(T, ch)) {   // q, k, v split inside the kernel
            AttnTcParams tp{nullptr, out_dev, T, H, ch, B, 1.4426950408889634f / sqrtf((float)ch), nullptr, qkv_dev};
            ce = launch_attention_tc(tp, math == EEGLDM_MATH_F16X3_TC, st);
        } else {
            CU(cudaMalloc((void**)&q16, attn_qkv16_bytes(B, T, H, ch)));
            ce = launch_qkv_split(qkv_dev, q16, B, T, H, ch, st);
            AttnTcParams tp{q16, out_dev, T, H, ch, B, 1.4426950408889634f / sqrtf((float)ch)};
            const size_t sb = attn_tc_scratch_bytes(B, T, H, ch);   // T > 256: key blocks + merge
            if (ce == cudaSuccess && sb) ce = cudaMalloc((void**)&part, sb);
            tp.part_out = part;
            if (ce == cudaSuccess) ce = launch_attention_tc(tp, math == EEGLDM_MATH_F16X3_TC, st);
        }
    } else {
        AttnParams ap{qkv_dev, out_dev, T, H, ch, B};
        ce = launch_attention_simt(ap, st);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (q16) cudaFree(q16);
    if (part) cudaFree(part);
    if (ce != cudaSuccess) return cuda_fail(ce, "attention launch");
    return EEGLDM_OK;
}

}  // extern "C"

// ================================================================================================ AEKL training step
namespace {

std::vector<float> unpack_conv(const float* packed, int Cout, int Cin, int k) {   // inverse of pack_conv
    std::vector<float> out((size_t)Cout * Cin * k);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int kk = 0; kk < k; ++kk) out[((size_t)co * Cin + ci) * k + kk] = packed[((size_t)ci * k + kk) * Cout + co];
    return out;
}

int aekl_train_init(eegldm_aekl* h) {
    int r = h->ps.check_all_loaded();
    if (r) return r;
    delete h->train;
    h->train = new AeklTrain();
    AeklTrain& t = *h->train;
    std::vector<float> flat;
    const ParamSet& ps = h->ps;
    auto push = [&](const std::string& name, bool conv, int cout = 0, int cin = 0, int k = 0) -> size_t {
        const HostParam& hp = ps.params[ps.index.at(name)];
        const size_t off = (flat.size() + 63) & ~size_t(63);
        flat.resize(off);
        if (conv) { auto pk = pack_conv(hp.data, cout, cin, k); flat.insert(flat.end(), pk.begin(), pk.end()); }
        else flat.insert(flat.end(), hp.data.begin(), hp.data.end());
        t.entries.push_back({name, off, hp.shape, conv});
        return off;
    };
    auto stage = [&](const ALayer& l) {
        TrainOff o;
        const std::string& p = l.prefix;
        switch (l.kind) {
            case ALayer::CONV: o.w1 = push(p + ".conv.weight", true, l.cout, l.cin, l.k); o.b1 = push(p + ".conv.bias", false); break;
            case ALayer::RES:
                o.g1 = push(p + ".norm1.weight", false); o.be1 = push(p + ".norm1.bias", false);
                o.w1 = push(p + ".conv1.conv.weight", true, l.cout, l.cin, 3); o.b1 = push(p + ".conv1.conv.bias", false);
                o.g2 = push(p + ".norm2.weight", false); o.be2 = push(p + ".norm2.bias", false);
                o.w2 = push(p + ".conv2.conv.weight", true, l.cout, l.cout, 3); o.b2 = push(p + ".conv2.conv.bias", false);
                if (l.cin != l.cout) { o.ws = push(p + ".nin_shortcut.conv.weight", true, l.cout, l.cin, 1); o.bs = push(p + ".nin_shortcut.conv.bias", false); }
                break;
            case ALayer::DOWN: case ALayer::UP:
                o.w1 = push(p + ".conv.conv.weight", true, l.cin, l.cin, 3); o.b1 = push(p + ".conv.conv.bias", false); break;
            case ALayer::NORM: o.g1 = push(p + ".weight", false); o.be1 = push(p + ".bias", false); break;
        }
        return o;
    };
    for (auto& l : h->enc) t.enc.push_back(stage(l));
    for (auto& l : h->dec) t.dec.push_back(stage(l));
    t.qmu = stage(h->q_mu); t.qls = stage(h->q_ls); t.pq = stage(h->post_q);
    t.n = (flat.size() + 63) & ~size_t(63);
    flat.resize(t.n, 0.f);
    CU(cudaMalloc((void**)&t.P, t.n * sizeof(float)));
    CU(cudaMalloc((void**)&t.G, t.n * sizeof(float)));
    CU(cudaMalloc((void**)&t.M, t.n * sizeof(float)));
    CU(cudaMalloc((void**)&t.V, t.n * sizeof(float)));
    CU(cudaMalloc((void**)&t.losses, 4 * sizeof(float)));
    CU(cudaMemcpy(t.P, flat.data(), t.n * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemset(t.G, 0, t.n * sizeof(float)));
    CU(cudaMemset(t.M, 0, t.n * sizeof(float)));
    CU(cudaMemset(t.V, 0, t.n * sizeof(float)));
    return EEGLDM_OK;
}

// copy the trained parameters back into the host state_dict and the inference weight pool
int aekl_train_sync(eegldm_aekl* h) {
    AeklTrain* t = h->train;
    if (!t || !t->dirty) return EEGLDM_OK;
    std::vector<float> flat(t->n);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(flat.data(), t->P, t->n * sizeof(float), cudaMemcpyDeviceToHost));
    for (auto& en : t->entries) {
        HostParam& hp = h->ps.params[h->ps.index.at(en.name)];
        if (en.conv) hp.data = unpack_conv(flat.data() + en.off, (int)en.shape[0], (int)en.shape[1], (int)en.shape[2]);
        else hp.data.assign(flat.begin() + en.off, flat.begin() + en.off + hp.numel());
    }
    t->dirty = false;
    return finalize_aekl(h);
}

struct TT { float* p = nullptr; int C = 0, T = 0; };

struct TrainRun {
    eegldm_aekl* h; AeklTrain* t; int B; cudaStream_t st; bool dry;
    float* base = nullptr; size_t off = 0;
    cudaError_t err = cudaSuccess;
    enum Kind { CONV, NORM, LATENT };
    struct Op {
        Kind kind; TT in, out, res; bool has_res = false, in_needs_grad = true;
        size_t ow = 0, ob = 0; int taps = 0, stride = 1, pad = 0, ups = 0;
        float *mean = nullptr, *rstd = nullptr; size_t og = 0, obe = 0; int G = 1, silu = 0;
        TT mu, lv, sigma; const float* eps = nullptr;
    };
    std::vector<Op> tape;
    std::unordered_map<const float*, std::pair<float*, bool>> grads;
    const float *dmu_ext = nullptr, *dsigma_ext = nullptr;   // gradients at the z_mu / z_sigma outputs (eegldm_aekl_backward)

    float* alloc(size_t n) { n = (n + 63) & ~size_t(63); float* p = dry ? nullptr : base + off; off += n; return p; }
    TT tensor(int C, int T) { TT x; x.C = C; x.T = T; x.p = alloc((size_t)B * T * C); return x; }
    void ck(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }
    size_t numel(const TT& x) const { return (size_t)B * x.T * x.C; }

    TT conv(const TT& in, size_t ow, size_t ob, int Cout, int taps, int stride, int pad, int ups, const TT* res, bool in_needs_grad = true) {
        const int Tc = ups ? in.T * 2 : in.T;
        const int Tout = stride == 1 ? Tc : (Tc + 1 - 3) / 2 + 1;
        TT out = tensor(Cout, Tout);
        if (!dry) {
            ConvParams p{};
            p.seg[0] = ConvSeg{in.p, nullptr, in.C, 0, nullptr, nullptr, 0, ups ? RS_NEAREST2 : RS_NONE, in.T, t->P + ow, taps};
            p.nseg = 1; p.Cout = Cout; p.Tout = Tout; p.Tc = Tc; p.stride = stride; p.pad_left = pad; p.bias = t->P + ob;
            if (res) { p.res = res->p; p.res_mode = RS_NONE; p.res_Tin = Tout; }
            p.out = out.p; p.B = B;
            ck(launch_conv_simt(p, st));
            Op op; op.kind = CONV; op.in = in; op.out = out; op.has_res = res != nullptr; if (res) op.res = *res;
            op.ow = ow; op.ob = ob; op.taps = taps; op.stride = stride; op.pad = pad; op.ups = ups; op.in_needs_grad = in_needs_grad;
            tape.push_back(op);
        }
        return out;
    }
    TT norm(const TT& x, size_t og, size_t obe, int G, int silu) {
        float* ss = alloc((size_t)B * x.C * 2);
        float* mr = alloc((size_t)B * G * 2);
        const int nsplit = groupnorm_nsplit(x.C, x.T, G);
        float* part = alloc((size_t)B * nsplit * G * 3);
        TT a = tensor(x.C, x.T);
        if (!dry) {
            GnParams p{};
            p.src0 = x.p; p.C0 = x.C; p.T = x.T; p.G = G; p.gamma = t->P + og; p.beta = t->P + obe; p.eps = 1e-6f; p.B = B;
            p.nsplit = nsplit; p.scale = ss; p.shift = ss + (size_t)B * x.C; p.partial = part;
            p.mean_out = mr; p.rstd_out = mr + (size_t)B * G;
            ck(launch_gn_act_fwd(p, a.p, silu, st));
            Op op; op.kind = NORM; op.in = x; op.out = a; op.mean = p.mean_out; op.rstd = p.rstd_out; op.og = og; op.obe = obe; op.G = G; op.silu = silu;
            tape.push_back(op);
        }
        return a;
    }
    TT blocks(const std::vector<ALayer>& layers, const std::vector<TrainOff>& offs, TT cur, bool first_needs_grad) {
        const int G = h->cfg.norm_num_groups;
        bool needs = first_needs_grad;
        for (size_t i = 0; i < layers.size(); ++i) {
            const ALayer& l = layers[i]; const TrainOff& o = offs[i];
            switch (l.kind) {
                case ALayer::CONV: cur = conv(cur, o.w1, o.b1, l.cout, l.k, 1, l.k / 2, 0, nullptr, needs); break;
                case ALayer::RES: {
                    TT a1 = norm(cur, o.g1, o.be1, G, 1);
                    TT h1 = conv(a1, o.w1, o.b1, l.cout, 3, 1, 1, 0, nullptr);
                    TT a2 = norm(h1, o.g2, o.be2, G, 1);
                    if (l.cin != l.cout) {
                        TT sc = conv(cur, o.ws, o.bs, l.cout, 1, 1, 0, 0, nullptr);
                        cur = conv(a2, o.w2, o.b2, l.cout, 3, 1, 1, 0, &sc);
                    } else {
                        TT x = cur;
                        cur = conv(a2, o.w2, o.b2, l.cout, 3, 1, 1, 0, &x);
                    }
                    break;
                }
                case ALayer::DOWN: cur = conv(cur, o.w1, o.b1, l.cin, 3, 2, 0, 0, nullptr); break;
                case ALayer::UP: cur = conv(cur, o.w1, o.b1, l.cin, 3, 1, 1, 1, nullptr); break;
                case ALayer::NORM: cur = norm(cur, o.g1, o.be1, G, 0); break;
            }
            needs = true;
        }
        return cur;
    }
    // gradient slot of a tensor: (pointer, already holds a value?)
    std::pair<float*, bool>& slot(const TT& x) {
        auto it = grads.find(x.p);
        if (it == grads.end()) it = grads.emplace(x.p, std::make_pair(alloc(numel(x)), false)).first;
        return it->second;
    }
    void backward(float kl_weight) {
        for (auto it = tape.rbegin(); it != tape.rend(); ++it) {
            Op& op = *it;
            if (op.kind == CONV) {
                auto dy = grads.find(op.out.p);
                if (dy == grads.end() || !dy->second.second) continue;   // output unused by the loss
                ConvGradParams p{};
                if (op.has_res) {   // identity residual: its gradient slot receives dy in the same launch
                    auto& gr = slot(op.res);
                    p.dres = gr.first; p.dres_accumulate = gr.second;
                    gr.second = true;
                }
                p.dy = dy->second.first; p.a = op.in.p; p.w = t->P + op.ow; p.dw = t->G + op.ow; p.db = t->G + op.ob;
                p.Cin = op.in.C; p.Cout = op.out.C; p.taps = op.taps; p.stride = op.stride; p.pad = op.pad; p.ups = op.ups;
                p.Tin = op.in.T; p.Tc = op.ups ? op.in.T * 2 : op.in.T; p.Tout = op.out.T; p.B = B;
                if (op.in_needs_grad) {
                    auto& gi = slot(op.in);
                    p.da = gi.first; p.accumulate = gi.second;
                    gi.second = true;
                }
                ck(launch_conv_bwd(p, st));
            } else if (op.kind == NORM) {
                auto da = grads.find(op.out.p);
                if (da == grads.end() || !da->second.second) continue;
                auto& gx = slot(op.in);
                NormGradParams p{};
                p.da = da->second.first; p.x = op.in.p; p.mean = op.mean; p.rstd = op.rstd; p.gamma = t->P + op.og; p.beta = t->P + op.obe;
                p.m12 = alloc((size_t)B * op.G * 2); p.dgamma = t->G + op.og; p.dbeta = t->G + op.obe; p.dx = gx.first;
                p.C = op.in.C; p.T = op.in.T; p.G = op.G; p.B = B; p.silu = op.silu; p.accumulate = gx.second;
                ck(launch_norm_act_bwd(p, st));
                gx.second = true;
            } else {   // LATENT: z = mu + eps*sigma ; KL
                auto dz = grads.find(op.out.p);
                if (dz == grads.end() || !dz->second.second) continue;
                auto& gm = slot(op.mu); auto& gl = slot(op.lv);
                ck(launch_latent(op.mu.p, op.lv.p, op.eps, nullptr, nullptr, dz->second.first, gm.first, gl.first, nullptr, kl_weight, B,
                                 numel(op.mu), st, dmu_ext, dsigma_ext));
                gm.second = gl.second = true;
            }
        }
    }
};

struct AeklPending { TrainRun run; TT recon, mu, sigma; int B, L; };
void aekl_pending_free(AeklPending* p) { delete p; }

// the forward pass of the training step: encoder, reparameterisation with the caller's eps, decoder (tape recorded unless dry)
TT aekl_train_forward(eegldm_aekl* h, AeklTrain& t, TrainRun& tr, const float* x_dev, const float* eps_dev, int L, TT* mu_out, TT* sigma_out,
                      cudaStream_t st) {
    const int z = h->cfg.latent_channels, T = L / h->down_factor();
    TT x; x.p = const_cast<float*>(x_dev); x.C = 1; x.T = L;
    TT h3 = tr.blocks(h->enc, t.enc, x, /*first_needs_grad=*/false);
    TT mu = tr.conv(h3, t.qmu.w1, t.qmu.b1, z, 1, 1, 0, 0, nullptr);
    TT lv = tr.conv(h3, t.qls.w1, t.qls.b1, z, 1, 1, 0, 0, nullptr);
    TT sigma = tr.tensor(z, T), zz = tr.tensor(z, T);
    if (!tr.dry) {
        tr.ck(launch_latent(mu.p, lv.p, eps_dev, sigma.p, zz.p, nullptr, nullptr, nullptr, t.losses + 1, 0.f, tr.B, tr.numel(mu), st));
        TrainRun::Op op; op.kind = TrainRun::LATENT; op.out = zz; op.mu = mu; op.lv = lv; op.sigma = sigma; op.eps = eps_dev;
        tr.tape.push_back(op);
    }
    if (mu_out) *mu_out = mu;
    if (sigma_out) *sigma_out = sigma;
    TT pq = tr.conv(zz, t.pq.w1, t.pq.b1, z, 1, 1, 0, 0, nullptr);
    return tr.blocks(h->dec, t.dec, pq, true);
}
}  // namespace

extern "C" {

// AutoencoderKL.forward(x) in training mode across an autograd boundary (the reference's own loop, train_autoencoderkl.py:204-220:
// model(x) -> losses in PyTorch -> loss_g.backward() -> optimizer_g.step()): the forward pass keeps its tape inside the handle ...
int eegldm_aekl_forward_train(eegldm_aekl* h, const float* x_dev, const float* eps_dev, float* recon_dev, float* z_mu_dev, float* z_sigma_dev,
                              int B, int L, void* stream) {
    if (!h || !x_dev || !eps_dev || !recon_dev || !z_mu_dev || !z_sigma_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (h->cfg.in_channels != 1 || h->cfg.out_channels != 1 || h->cfg.latent_channels != 1)
        return fail(EEGLDM_ERR_INVALID, "training supports in / out / latent channels = 1 (every reference config)");
    const int f = h->down_factor();
    if (B <= 0 || L <= 0 || L % f) return fail(EEGLDM_ERR_SHAPE, "L must be a positive multiple of 2^(levels-1), B > 0");
    for (int i = 0; i < h->cfg.n_levels; ++i)
        if (h->cfg.num_channels[i] / h->cfg.norm_num_groups > 256) return fail(EEGLDM_ERR_INVALID, "channels per group > 256 not supported in training");
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->train) { int r = aekl_train_init(h); if (r) return r; }
    AeklTrain& t = *h->train;
    if (t.pending) { aekl_pending_free(t.pending); t.pending = nullptr; }
    TrainRun sizing{h, &t, B, st, true};
    aekl_train_forward(h, t, sizing, nullptr, nullptr, L, nullptr, nullptr, st);
    int r = ensure(t.arena, t.arena_cap, sizing.off * 2 + (size_t)B * (L / f) + (size_t)(1 << 20));
    if (r) return r;
    auto* pd = new AeklPending{TrainRun{h, &t, B, st, false}, TT{}, TT{}, TT{}, B, L};
    pd->run.base = t.arena;
    // eps is read again by the backward pass: keep a copy in the arena (the caller's tensor may be gone by then)
    const size_t nz = (size_t)B * (L / f);
    float* eps_keep = pd->run.alloc(nz);
    CU(cudaMemcpyAsync(eps_keep, eps_dev, nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(t.losses, 0, 4 * sizeof(float), st));
    pd->recon = aekl_train_forward(h, t, pd->run, x_dev, eps_keep, L, &pd->mu, &pd->sigma, st);
    if (pd->run.err != cudaSuccess) { cudaError_t e = pd->run.err; delete pd; return cuda_fail(e, "training forward"); }
    CU(cudaMemcpyAsync(recon_dev, pd->recon.p, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(z_mu_dev, pd->mu.p, nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(z_sigma_dev, pd->sigma.p, nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    t.pending = pd;
    return EEGLDM_OK;
}

// ... and the backward pass takes dL/d(reconstruction), dL/d(z_mu), dL/d(z_sigma) (each nullable = zero) and leaves the parameter
// gradients in the handle (eegldm_aekl_train_export(h, 1, name, ...)); dx_dev (nullable) receives dL/dx.
int eegldm_aekl_backward(eegldm_aekl* h, const float* d_recon_dev, const float* d_mu_dev, const float* d_sigma_dev, float* dx_dev, void* stream) {
    if (!h || !h->train || !h->train->pending) return fail(EEGLDM_ERR_MISSING, "no recorded forward pass (eegldm_aekl_forward_train)");
    if (dx_dev) return fail(EEGLDM_ERR_INVALID, "the gradient with respect to the input signal is not computed (the first conv skips it)");
    AeklTrain& t = *h->train;
    AeklPending* pd = t.pending;
    cudaStream_t st = (cudaStream_t)stream;
    TrainRun& tr = pd->run;
    tr.st = st;
    CU(cudaMemsetAsync(t.G, 0, t.n * sizeof(float), st));
    auto& gr = tr.slot(pd->recon);
    if (d_recon_dev) CU(cudaMemcpyAsync(gr.first, d_recon_dev, tr.numel(pd->recon) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else CU(cudaMemsetAsync(gr.first, 0, tr.numel(pd->recon) * sizeof(float), st));
    gr.second = true;
    tr.dmu_ext = d_mu_dev; tr.dsigma_ext = d_sigma_dev;
    tr.backward(0.f);
    const cudaError_t e = tr.err;
    const bool overflow = tr.off > t.arena_cap;
    aekl_pending_free(pd);
    t.pending = nullptr;
    if (e != cudaSuccess) return cuda_fail(e, "training backward");
    if (overflow) return fail(EEGLDM_ERR_NOMEM, "training arena overflow");
    return EEGLDM_OK;
}

int eegldm_jukebox_loss(const float* input_dev, const float* target_dev, int B, int C, int N, int reduction, float* loss_dev,
                        float* grad_input_dev, void* stream) {
    if (!input_dev || !target_dev || !loss_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (C != 1) return fail(EEGLDM_ERR_INVALID, "JukeboxLoss: only single-channel signals (the reference's in/out_channels = 1)");
    if (reduction != 0 && reduction != 1) return fail(EEGLDM_ERR_INVALID, "reduction must be 0 (sum) or 1 (mean)");
    if (B < 0 || N < 2) return fail(EEGLDM_ERR_SHAPE, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMemsetAsync(loss_dev, 0, sizeof(float), st));
    std::string err;
    const int r = spectral_loss(input_dev, target_dev, B, N, reduction, 1.f, loss_dev, grad_input_dev, 1.f, 0, st, &err);
    if (r == 1) return fail(EEGLDM_ERR_CUDA, "cuFFT: " + err);
    if (r) return cuda_fail((cudaError_t)r, "spectral loss");
    return EEGLDM_OK;
}

// one training step; disc == null: the generator half only (eegldm_aekl_train_step), else the full step of
// train_autoencoderkl.py:204-234 (eegldm_aekl_train_step_adv).  losses_host: {l1, kl, spectral, total_g[, gen_adv, disc]}
static int aekl_train_step_impl(eegldm_aekl* h, eegldm_disc* disc, const float* x_dev, const float* eps_dev, int B, int L,
                                const eegldm_aekl_train_cfg* cfg, float adv_weight, float lr_d, int no_act, float* losses_host, void* stream) {
    if (!h || !cfg) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (h->cfg.in_channels != 1 || h->cfg.out_channels != 1)
        return fail(EEGLDM_ERR_INVALID, "training step supports in/out_channels = 1 (every reference config)");
    const int f = h->down_factor();
    if (B <= 0 || L <= 0 || L % f) return fail(EEGLDM_ERR_SHAPE, "L must be a positive multiple of 2^(levels-1), B > 0");
    if (!x_dev || !eps_dev) return fail(EEGLDM_ERR_INVALID, "null argument");
    for (int i = 0; i < h->cfg.n_levels; ++i)
        if (h->cfg.num_channels[i] / h->cfg.norm_num_groups > 256) return fail(EEGLDM_ERR_INVALID, "channels per group > 256 not supported in training");
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->train) { int r = aekl_train_init(h); if (r) return r; }
    AeklTrain& t = *h->train;
    const int z = h->cfg.latent_channels, T = L / f;
    auto run = [&](TrainRun& tr) { return aekl_train_forward(h, t, tr, x_dev, eps_dev, L, nullptr, nullptr, st); };
    // eps is given in the reference NCL layout; the engine's latent tensors are channels-last
    if (z != 1) return fail(EEGLDM_ERR_INVALID, "training step supports latent_channels = 1 (config_aekl_eeg_2_2_4_spec.yaml)");
    TrainRun sizing{h, &t, B, st, true};
    run(sizing);
    const size_t need = sizing.off * 2 + (size_t)(1 << 20);
    int r = ensure(t.arena, t.arena_cap, need);
    if (r) return r;
    CU(cudaMemsetAsync(t.G, 0, t.n * sizeof(float), st));
    CU(cudaMemsetAsync(t.losses, 0, 4 * sizeof(float), st));
    TrainRun tr{h, &t, B, st, false};
    tr.base = t.arena;
    TT recon = run(tr);
    if (tr.err != cudaSuccess) return cuda_fail(tr.err, "training forward");
    // losses: L1 (mean) + spectral_weight * Jukebox(sum) on the reconstruction; KL handled in the latent op
    auto& gr = tr.slot(recon);
    CU(launch_l1_loss(recon.p, x_dev, gr.first, t.losses + 0, 1.f, tr.numel(recon), st));
    gr.second = true;
    if (cfg->spectral_weight != 0.f) {
        std::string err;
        const int sr = spectral_loss(recon.p, x_dev, B, L, 0, 1.f, t.losses + 2, gr.first, cfg->spectral_weight, 1, st, &err);
        if (sr == 1) return fail(EEGLDM_ERR_CUDA, "cuFFT: " + err);
        if (sr) return cuda_fail((cudaError_t)sr, "spectral loss");
    }
    if (disc) {   // logits_fake = D(recon)[-1]; generator_loss = adv_loss(logits_fake, real) -- its gradient joins d loss / d recon
        r = disc_prepare_step(disc, B, L, st);
        if (r) return r;
        r = disc_generator_term(disc, recon.p, B, L, adv_weight, no_act, gr.first, st);
        if (r) return r;
    }
    tr.backward(cfg->kl_weight);
    if (tr.err != cudaSuccess) return cuda_fail(tr.err, "training backward");
    if (tr.off > t.arena_cap) return fail(EEGLDM_ERR_NOMEM, "training arena overflow");
    CU(launch_axpy(t.losses + 0, t.losses + 3, 1.f, 0, 1, st));
    CU(launch_axpy(t.losses + 1, t.losses + 3, cfg->kl_weight, 1, 1, st));
    CU(launch_axpy(t.losses + 2, t.losses + 3, cfg->spectral_weight, 1, 1, st));
    if (disc) CU(launch_axpy(disc_losses_dev(disc) + 0, t.losses + 3, adv_weight, 1, 1, st));
    if (cfg->lr > 0.f) {
        t.step += 1;
        CU(launch_adam(t.P, t.G, t.M, t.V, cfg->lr, cfg->beta1, cfg->beta2, cfg->adam_eps, t.step, t.n, st));
        t.dirty = true;
    }
    if (disc) {   // discriminator part (train_autoencoderkl.py:223-234): recon is the pre-update reconstruction, detached
        r = disc_step(disc, x_dev, B, L, adv_weight, no_act, lr_d, cfg->beta1, cfg->beta2, cfg->adam_eps, st);
        if (r) return r;
    }
    if (losses_host) {
        CU(cudaMemcpyAsync(losses_host, t.losses, 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (disc) {
            float dl[4];
            CU(cudaMemcpyAsync(dl, disc_losses_dev(disc), 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            losses_host[4] = dl[0];                      // generator_loss  (train_autoencoderkl.py:214)
            losses_host[5] = 0.5f * (dl[1] + dl[2]);     // discriminator_loss (:229)
        }
        CU(cudaStreamSynchronize(st));
    }
    return EEGLDM_OK;
}

int eegldm_aekl_train_step(eegldm_aekl* h, const float* x_dev, const float* eps_dev, int B, int L, const eegldm_aekl_train_cfg* cfg,
                           float* losses_host, void* stream) {
    return aekl_train_step_impl(h, nullptr, x_dev, eps_dev, B, L, cfg, 0.f, 0.f, 0, losses_host, stream);
}

int eegldm_aekl_train_step_adv(eegldm_aekl* h, eegldm_disc* disc, const float* x_dev, const float* eps_dev, int B, int L,
                               const eegldm_aekl_adv_train_cfg* cfg, float* losses_host, void* stream) {
    if (!cfg || !disc) return fail(EEGLDM_ERR_INVALID, "null argument");
    eegldm_aekl_train_cfg g{cfg->kl_weight, cfg->spectral_weight, cfg->lr_g, cfg->beta1, cfg->beta2, cfg->adam_eps};
    return aekl_train_step_impl(h, disc, x_dev, eps_dev, B, L, &g, cfg->adv_weight, cfg->lr_d, cfg->no_activation_leastsq, losses_host, stream);
}

int eegldm_aekl_train_export(eegldm_aekl* h, int what, const char* name, float* host_out) {
    if (!h || !name || !host_out) return fail(EEGLDM_ERR_INVALID, "null argument");
    if (!h->train) return fail(EEGLDM_ERR_MISSING, "no training step has run");
    AeklTrain& t = *h->train;
    if (what != 0 && what != 1) return fail(EEGLDM_ERR_INVALID, "what must be 0 (parameter) or 1 (gradient)");
    for (auto& en : t.entries) {
        if (en.name != name) continue;
        size_t n = 1;
        for (auto s : en.shape) n *= (size_t)s;
        std::vector<float> tmp(n);
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(tmp.data(), (what ? t.G : t.P) + en.off, n * sizeof(float), cudaMemcpyDeviceToHost));
        if (en.conv) tmp = unpack_conv(tmp.data(), (int)en.shape[0], (int)en.shape[1], (int)en.shape[2]);
        std::memcpy(host_out, tmp.data(), n * sizeof(float));
        return EEGLDM_OK;
    }
    return fail(EEGLDM_ERR_MISSING, std::string("unknown state_dict key: ") + name);
}

int eegldm_aekl_train_sync(eegldm_aekl* h) {
    if (!h) return fail(EEGLDM_ERR_INVALID, "null handle");
    return aekl_train_sync(h);
}

}  // extern "C"

#include "unet_train.inc"
