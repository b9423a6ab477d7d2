// tcgen05 self-attention for the latent UNet (QKVAttentionLegacy.forward, src/models/unet.py:107-125):
//     w = softmax_fp32( (q * ch^-1/4)^T (k * ch^-1/4) )   over keys,   a = v w^T        per (sample, head)
// as two tensor-pipe GEMMs around an in-kernel softmax, with the scores never leaving the SM.
//
//   qkv_split_kernel   fp32 qkv [B][T][H*3*ch] (legacy head layout, unet.py:116-118) -> fp16 hi/lo images:
//                        K    : [k-step = ch/32][hi|lo][T/8][4][8 pos][8 ch]   (K-major core matrices; a row range is contiguous)
//                        All three images hold their rows (queries; keys of K and V alike -- the key order is free as long as K and V
//                        agree) in the conv kernel's phase-strided order, per 128-row tile:
//                               a tile holds nseg = min(8, T/16 - 8*tile) segments of 16 positions, and its row m is position
//                               tile*128 + (m % nseg)*16 + m / nseg (attn_q_row): consecutive rows = the same slot of consecutive
//                               segments.  A TMEM lane of the qkv conv and of this kernel is then (slot, segment) with the
//                               segment fastest, so eight consecutive lanes store whole 64- / 128-byte runs -- of the Q image in
//                               the qkv conv's epilogue, of proj_out's operand image in this kernel's epilogue -- instead of 16-byte
//                               pieces of 32 different lines
//                        V    : [ch/128][T/32][hi|lo][4][16][8 key][8 ch]      (MN-major B operand for P.V)
//   attn_tc_kernel     one CTA per (128-query tile, head, sample):
//                        S = Q K^T   : M=128, N=T (<=256), K=ch, accumulators in TMEM columns [0,T) and [256,256+T)
//                        softmax     : 128 threads, one query row each: TMEM -> registers, max, exp2, sum,
//                                      P = exp(.) split to fp16 hi/lo -> shared memory (K-major core matrices)
//                        O = P V     : per 128-channel chunk, M=128, N=128, K=T, double-buffered in TMEM;
//                                      epilogue scales by 1/rowsum and stores fp32 channels-last.
//   Arithmetic is the same f16x3 scheme as conv_tc.cu (hi*hi | hi*lo + lo*hi in a second accumulator); the
//   non-X3 instantiation issues hi*hi only (fast mode, not a parity mode).
#include "kernels.cuh"
#include "tc_common.cuh"

namespace eegldm {
using namespace tc;
namespace {

// operand stage ring: as many stages as fit next to the two P buffers (T = 192: 3, T = 256: 2, T <= 128: 4) -- the S phase
// streams 40 KB per k-step and is bound by the load round trip with only two in flight
__host__ __device__ inline int attn_stages(int T);
constexpr int Q_HALF = 8192;           // 128 rows x 32 ch x 2 B
constexpr int V_HALF = 8192;           // 32 keys x 128 ch x 2 B
constexpr int NUM_THREADS = 192;

// row of position t in the Q image: tile t / 128 holds nseg = min(8, T/16 - 8*tile) segments; its rows run over the segments first
__host__ __device__ inline int attn_q_row(int t, int T) {
    const int qt = t >> 7, tt = t & 127, nseg = min(8, (T >> 4) - 8 * qt);
    return qt * 128 + nseg * (tt & 15) + (tt >> 4);
}
__host__ __device__ inline int stage_bytes(int T) { return 2 * Q_HALF + 2 * T * 64; }   // Q hi/lo + K hi/lo of one 32-ch k-step
__host__ __device__ inline int p_half_bytes(int T) { return (T / 8) * 2048; }           // P hi (or lo): [T/8][16][8][8] fp16
__host__ __device__ inline int attn_stages(int T) {
    const int n = (227 * 1024 - 2 * p_half_bytes(T) - 256 - 2048) / stage_bytes(T);   // 256 B barriers + 2 KB row max / sum exchange
    return n > 4 ? 4 : n;
}

// ------------------------------------------------------------------------------------------------ qkv split
// one thread per (sample, head, q|k|v, 8-channel chunk, position); position fastest so that 8 lanes fill a 128-byte line
__global__ void qkv_split_kernel(const float* __restrict__ qkv, uint8_t* __restrict__ dst, int T, int H, int ch, size_t total,
                                 int* __restrict__ range_flag) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int t = (int)(idx % T);
    size_t rest = idx / T;
    const int c8 = (int)(rest % (ch / 8)); rest /= (ch / 8);
    const int which = (int)(rest % 3); rest /= 3;
    const int h = (int)(rest % H);
    const size_t b = rest / H;
    const int c = c8 * 8;
    const float* src = qkv + ((size_t)b * T + t) * ((size_t)H * 3 * ch) + (size_t)h * 3 * ch + (size_t)which * ch + c;
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(src)), x1 = __ldg(reinterpret_cast<const float4*>(src + 4));
    const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    uint4 hi, lo;
    if (range_flag) {   // f16x3 operand range (|x| < 65504, include/eegldm.h); inf and NaN compare above every finite magnitude
        uint32_t m = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) m = max(m, __float_as_uint(v[e]) & 0x7FFFFFFFu);
        if (m >= 0x477FE000u) atomicOr(range_flag, 1);
    }
    split8_f16(v, hi, lo);
    const size_t plane = (size_t)4 * ch * T;                       // bytes of one of q / k / v (hi + lo)
    uint8_t* base = dst + ((size_t)b * H + h) * 3 * plane + (size_t)which * plane;
    size_t ohi, olo;
    const int tr = attn_q_row(t, T);   // image row of position t: queries and keys in the phase-strided tile order
    if (which < 2) {   // Q, K: [ks][hi|lo][T/8][4][8][8]
        const int ks = c / 32, cg = (c % 32) / 8;
        const size_t o = (size_t)(tr / 8) * 512 + cg * 128 + (tr % 8) * 16;
        ohi = ((size_t)ks * 2 + 0) * ((size_t)T * 64) + o;
        olo = ((size_t)ks * 2 + 1) * ((size_t)T * 64) + o;
    } else {           // V: [ch/128][T/32][hi|lo][4][16][8][8]
        const size_t blk = (size_t)(c / 128) * (T / 32) + tr / 32;
        const size_t o = (size_t)((tr % 32) / 8) * 2048 + ((c % 128) / 8) * 128 + (tr % 8) * 16;
        ohi = (blk * 2 + 0) * V_HALF + o;
        olo = (blk * 2 + 1) * V_HALF + o;
    }
    *reinterpret_cast<uint4*>(base + ohi) = hi;
    *reinterpret_cast<uint4*>(base + olo) = lo;
}

// ------------------------------------------------------------------------------------------------ attention
// DIRECT: q, k, v are read as fp32 straight from the qkv conv's output [B][T][H*3*ch] (legacy head layout): six producer warps
// split them to fp16 hi/lo and write the stage images with st.shared (the qkv_split pass and its 2 x 1.2 GB of HBM traffic
// per block at B = 1024 disappear; measured: the attention kernel slows down by exactly what the pass cost, 0.47 -> 0.99 ms,
// so this form is OFF by default -- eegldm_set_conv_tuning bit 4).  Item = 8 channels of one row (two float4 loads, two 16-byte stores); one stage of loads
// is in flight in registers while the previous one is split.  Lane mapping: S stages -- 4 consecutive lanes cover one row's
// 32 channels (128 contiguous bytes in, 8 rows x 64 B = 512 contiguous bytes of the image out per warp); PV stages -- a warp
// covers 8 keys x 4 channel groups, which is conflict-free on the MN-major V image (4 x 128 contiguous bytes).
// Softmax / epilogue warps: ONE warpgroup leaves every TMEM load -> exp2 / split -> store chain exposed (one warp per scheduler);
// the pre-split form runs TWO (warps 0-3 and 6-9), each taking half of the score columns in the softmax (row maxima and sums are
// exchanged through 2 KB of shared memory) and half of every 128-channel output chunk in the epilogue (tools/attn_timeline.py:
// PV + epilogue 24.9 k -> see profiles/r02b_attn_timeline.txt).  The in-kernel-split form keeps one (warps 6-11 are its producers).
constexpr int ATTN_THREADS = 320;   // pre-split form: 4 softmax + loader + MMA issuer + 4 softmax
template <bool X3, bool DIRECT>
__global__ void __launch_bounds__(DIRECT ? 2 * NUM_THREADS : ATTN_THREADS, 1) attn_tc_kernel(const AttnTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int T = p.T, ch = p.ch;
    const int Tk = p.Tk ? p.Tk : T, nkb = T / Tk;   // keys of this CTA's block; nkb > 1: partial outputs, merged by attn_merge_kernel
    const int SB = stage_bytes(Tk), PH = p_half_bytes(Tk), NST = attn_stages(Tk);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sStage = sbase, sP = sbase + NST * SB, bars = sP + 2 * PH;
    // Stage slots.  The S phase streams 40 KB per k-step and is bound by the load round trip; the two P buffers are idle until the
    // scores are complete, so (pre-split form) the S phase also uses them as stage slots NST .. NSS-1.  The PV phase uses slots
    // 0 .. NST-1 only and packs VP = 2 of its 16 KB V stages into one slot when the slot is big enough (T >= 128).
    const int NSS = DIRECT ? NST : NST + (2 * PH) / SB;          // <= NST + 2
    const int VP = (!DIRECT && SB >= 4 * V_HALF) ? 2 : 1;
    const uint32_t barFull = bars, barEmpty = bars + 8 * 6, barS = bars + 16 * 6, barP = barS + 8, barOfull = barP + 8,
                   barOempty = barOfull + 16;                     // (room for 6 stage barriers each: NST <= 4, + 2)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + NST * SB + 2 * PH + 16 * 6 + 48);
    auto slot_addr = [&](int i) -> uint32_t { return i < NST ? sStage + i * SB : sP + (i - NST) * SB; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mt = blockIdx.x / nkb, kblk = blockIdx.x - mt * nkb, key0 = kblk * Tk, h = blockIdx.y, b = blockIdx.z;
    const int nks = ch / 32, nchunk = ch / 128, nss = Tk / 32;
    const int nseg = min(8, (T >> 4) - 8 * mt);   // 16-position segments of this query tile (rows nseg*16 .. 127 of the MMA are unused)
    const size_t plane = (size_t)4 * ch * T;
    const uint8_t* gq = p.qkv16 + ((size_t)b * p.H + h) * 3 * plane;
    const uint8_t* gk = gq + plane;
    const uint8_t* gv = gk + plane;

    if (tid == 0) {
        for (int i = 0; i < NSS; ++i) { mbar_init(barFull + 8 * i, DIRECT ? NUM_THREADS : 1); mbar_init(barEmpty + 8 * i, 1); }
        mbar_init(barS, 1);
        mbar_init(barP, DIRECT ? 128 : 256);
        for (int i = 0; i < 2; ++i) { mbar_init(barOfull + 8 * i, 1); mbar_init(barOempty + 8 * i, DIRECT ? 128 : 256); }
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // eegldm_bench_attention: per-CTA cycle stamps written by thread 0 (softmax warp 0): [0] start -> barriers ready, [1] -> scores
    // complete (S phase), [2] -> P written (softmax), [3] -> last output chunk stored (PV + epilogue), [4] total
    const bool tl = p.timeline != nullptr;
    const long long tl0 = tl ? clock64() : 0;
    long long tl1 = 0, tl2 = 0;

    if (warp < 4 || (!DIRECT && warp >= 6)) {
        // ================================================================ softmax, then epilogue (one or two warpgroups)
        constexpr int NWG = DIRECT ? 1 : 2;
        const int wg = warp < 4 ? 0 : 1;                            // column half this warpgroup takes
        const int row = (warp & 3) * 32 + lane;                    // MMA row = TMEM lane (a warp reads the lane quarter warp % 4):
                                                                   // (slot = row / nseg, segment = row % nseg) of the tile
        const int t = mt * 128 + (row % nseg) * 16 + row / nseg;  // its position (Q image row order, see the file comment)
        const bool rowv = row < 16 * nseg;
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        float* xmax = reinterpret_cast<float*>(smem + NST * SB + 2 * PH + 256);   // [2][128] row maxima, then [2][128] row sums
        float* xsum = xmax + 256;
        const int c_mid = NWG == 1 ? nss : (nss + 1) / 2;
        const int cb0 = (wg == 0 ? 0 : c_mid) * 32, cb1 = (wg == 0 ? c_mid : nss) * 32;   // this warpgroup's score columns
        mbar_wait(barS, 0);
        tc_fence_after();
        if (tl) tl1 = clock64();
        float mx = -INFINITY;
        for (int cb = cb0; cb < cb1; cb += 32) {
            uint32_t v[32];
            tmem_ld32(lane_addr + cb, v);
            if (X3) {
                uint32_t c2[32];
                tmem_ld32(lane_addr + 256 + cb, c2);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaf(__uint_as_float(c2[i]), 1.0f / LO_SCALE, __uint_as_float(v[i])));
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        if (NWG == 2) {
            xmax[wg * 128 + row] = mx;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            mx = fmaxf(xmax[row], xmax[128 + row]);
        }
        float sum = 0.f;
        const float sc = p.scale_log2e;               // ch^-1/2 * log2(e): scores are (q.k) * ch^-1/2 (unet.py:119-121)
        const float mxs = mx * sc;
        uint8_t* prow = smem + NST * SB + (row >> 3) * 128 + (row & 7) * 16;
        for (int cb = cb0; cb < cb1; cb += 32) {
            uint32_t v[32];
            tmem_ld32(lane_addr + cb, v);
            if (X3) {
                uint32_t c2[32];
                tmem_ld32(lane_addr + 256 + cb, c2);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaf(__uint_as_float(c2[i]), 1.0f / LO_SCALE, __uint_as_float(v[i])));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = exp2f(fmaf(__uint_as_float(v[8 * j + i]), sc, -mxs));
                    sum += e[i];
                }
                uint4 hi, lo;
                split8_f16(e, hi, lo);
                uint8_t* dst = prow + (size_t)((cb >> 3) + j) * 2048;
                *reinterpret_cast<uint4*>(dst) = hi;
                if (X3) *reinterpret_cast<uint4*>(dst + PH) = lo;
            }
        }
        fence_proxy_async_smem();      // P (generic-proxy stores) -> visible to the tensor core's async proxy
        tc_fence_before();             // the TMEM reads above are ordered before the MMAs that reuse the columns
        mbar_arrive(barP);
        if (NWG == 2) {
            xsum[wg * 128 + row] = sum;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            sum = xsum[row] + xsum[128 + row];
        }
        if (tl) tl2 = clock64();
        const float inv = 1.0f / sum;
        if (nkb > 1 && wg == 0 && rowv) p.part_ml[(((size_t)kblk * p.B + b) * p.H + h) * T + t] = make_float2(mxs, sum);
        float* orow = (nkb > 1 ? p.part_out + (size_t)kblk * p.B * T * ((size_t)p.H * ch) : p.out) + ((size_t)b * T + t) * ((size_t)p.H * ch) +
                      (size_t)h * ch;
        const int ob0 = NWG == 1 ? 0 : wg * 64, ob1 = NWG == 1 ? 128 : ob0 + 64;   // this warpgroup's columns of every output chunk
        for (int c = 0; c < nchunk; ++c) {
            const int buf = c & 1;
            mbar_wait(barOfull + 8 * buf, (c >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cb = ob0; cb < ob1; cb += 32) {
                uint32_t v[32];
                tmem_ld32(lane_addr + buf * 256 + cb, v);
                if (X3) {
                    uint32_t c2[32];
                    tmem_ld32(lane_addr + buf * 256 + 128 + cb, c2);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaf(__uint_as_float(c2[i]), 1.0f / LO_SCALE, __uint_as_float(v[i])));
                }
                if (rowv && p.out_u) {
                    // proj_out's operand image (conv_tc.cu U layout, 1x1 conv: halo slots are never read): this row is slot
                    // t%16 + 1 of 16-position segment b*T/16 + t/16; 8 channels = one 16-byte item per hi / lo half.  Eight
                    // consecutive lanes are the eight segments of one slot: one whole 128-byte line per store.
                    const int g16 = b * (T >> 4) + (t >> 4);
                    uint8_t* ub = p.out_u + (size_t)(g16 >> 3) * ((size_t)p.H * ch / 32) * (2 * TC_U_HALF_BYTES) +
                                  ((t & 15) + 1) * 128 + (g16 & 7) * 16;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float e[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) e[i] = __uint_as_float(v[8 * q + i]) * inv;
                        uint4 hi, lo;
                        if (X3) split8_f16(e, hi, lo);
                        else round8_bf16(e, hi);
                        const int cc = h * ch + c * 128 + cb + 8 * q;      // channel of the block output
                        uint8_t* dst = ub + (size_t)(cc >> 5) * (2 * TC_U_HALF_BYTES) + ((cc >> 3) & 3) * (TC_U_HALF_BYTES / 4);
                        *reinterpret_cast<uint4*>(dst) = hi;
                        if (X3) *reinterpret_cast<uint4*>(dst + TC_U_HALF_BYTES) = lo;
                    }
                } else if (rowv) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(orow + c * 128 + cb + 4 * q) =
                            make_float4(__uint_as_float(v[4 * q]) * inv, __uint_as_float(v[4 * q + 1]) * inv,
                                        __uint_as_float(v[4 * q + 2]) * inv, __uint_as_float(v[4 * q + 3]) * inv);
                }
            }
            tc_fence_before();
            mbar_arrive(barOempty + 8 * buf);
        }
        if (tl && tid == 0) {
            unsigned long long* o = p.timeline + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8;
            const long long t3 = clock64();
            o[1] = tl1 - tl0; o[2] = tl2 - tl1; o[3] = t3 - tl2; o[4] = t3 - tl0; o[5] = (unsigned long long)tl0;
        }
    } else if (DIRECT && warp >= 6) {
        // ================================================================ producers (192 threads): fp32 rows -> fp16 hi/lo stage images
        const int pt = tid - NUM_THREADS;              // warps 6-11
        const size_t rs = (size_t)p.H * 3 * ch;          // floats per qkv row
        const float* qrow0 = p.qkv32 + (size_t)b * T * rs + (size_t)h * 3 * ch;   // q of position 0; k at +ch, v at +2ch
        const int rowsQ = 16 * nseg, nS = (rowsQ + T) * 4, nV = 512;
        const int total = nks + nchunk * nss;
        constexpr int MAXI = 7;                          // items per thread and stage: (128 + 256) * 4 / 192 = 8 would need T = 256: see launcher
        struct Pre { float4 x[MAXI][2]; };
        auto src_of = [&](int it, int idx) -> const float* {
            if (it < nks) {
                const int row = idx >> 2, cg = idx & 3;
                if (row < rowsQ) {
                    const int tq = mt * 128 + (row % nseg) * 16 + row / nseg;  // phase-strided query row order
                    return qrow0 + (size_t)tq * rs + it * 32 + cg * 8;
                }
                return qrow0 + ch + (size_t)(row - rowsQ) * rs + it * 32 + cg * 8;
            }
            const int c = (it - nks) / nss, ss = (it - nks) % nss;
            const int cgrp = (idx & 3) | (((idx >> 5) & 3) << 2), key = ((idx >> 2) & 7) | ((idx >> 7) << 3);
            return qrow0 + 2 * ch + (size_t)(ss * 32 + key) * rs + c * 128 + cgrp * 8;
        };
        auto issue = [&](int it, Pre& P) {
            const int n = it < nks ? nS : nV;
#pragma unroll
            for (int j = 0; j < MAXI; ++j) {
                const int idx = pt + NUM_THREADS * j;
                if (idx < n) {
                    const float* sp = src_of(it, idx);
                    if (sp) {
                        P.x[j][0] = __ldg(reinterpret_cast<const float4*>(sp));
                        P.x[j][1] = __ldg(reinterpret_cast<const float4*>(sp + 4));
                    } else P.x[j][0] = P.x[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        Pre N;
        issue(0, N);
        for (int it = 0; it < total; ++it) {
            const Pre C = N;
            if (it + 1 < total) issue(it + 1, N);
            const int st = it % NST;
            mbar_wait(barEmpty + 8 * st, ((it / NST) & 1) ^ 1);
            uint8_t* stage = smem + (size_t)st * SB;
            const int n = it < nks ? nS : nV;
#pragma unroll
            for (int j = 0; j < MAXI; ++j) {
                const int idx = pt + NUM_THREADS * j;
                if (idx < n) {
                    const float v[8] = {C.x[j][0].x, C.x[j][0].y, C.x[j][0].z, C.x[j][0].w, C.x[j][1].x, C.x[j][1].y, C.x[j][1].z, C.x[j][1].w};
                    uint32_t off, half;
                    if (it < nks) {
                        const int row = idx >> 2, cg = idx & 3;
                        if (row < rowsQ) { off = (uint32_t)((row >> 3) * 512 + cg * 128 + (row & 7) * 16); half = Q_HALF; }
                        else {
                            const int tk = row - rowsQ;
                            off = (uint32_t)(2 * Q_HALF + (tk >> 3) * 512 + cg * 128 + (tk & 7) * 16); half = (uint32_t)T * 64;
                        }
                    } else {
                        const int cgrp = (idx & 3) | (((idx >> 5) & 3) << 2), key = ((idx >> 2) & 7) | ((idx >> 7) << 3);
                        off = (uint32_t)((key >> 3) * 2048 + cgrp * 128 + (key & 7) * 16); half = V_HALF;
                    }
                    uint4 hi, lo;
                    split8_f16(v, hi, lo);             // the attention MMAs are fp16 in both modes (the fast mode drops lo)
                    if (X3) *reinterpret_cast<uint4*>(stage + off + half) = lo;
                    *reinterpret_cast<uint4*>(stage + off) = hi;
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(barFull + 8 * st);
        }
    } else if (warp == 4) {
        // ================================================================ loader (pre-split images; idle in the DIRECT form)
        if (lane == 0 && !DIRECT) {
            uint32_t phE = 0;   // bit i: parity of the completed waits on slot i's "empty" barrier
            const uint32_t qb = (uint32_t)nseg * 1024, kb = (uint32_t)Tk * 64;   // the tile's 16*nseg query rows; this block's keys
            const size_t qhalf = (size_t)T * 64;                                  // bytes of one (k-step, hi | lo) plane of Q and of K
            for (int ks = 0; ks < nks; ++ks) {
                const int st = ks % NSS;
                const uint32_t dst = slot_addr(st);
                mbar_wait(barEmpty + 8 * st, ((phE >> st) & 1) ^ 1);
                phE ^= 1u << st;
                mbar_arrive_expect_tx(barFull + 8 * st, X3 ? 2 * (qb + kb) : qb + kb);
                const uint8_t* qs = gq + ((size_t)ks * 2) * qhalf + (size_t)mt * 16 * 512;
                const uint8_t* ks_ = gk + ((size_t)ks * 2) * qhalf + (size_t)key0 * 64;
                bulk_copy_g2s(dst, qs, qb, barFull + 8 * st);
                bulk_copy_g2s(dst + 2 * Q_HALF, ks_, kb, barFull + 8 * st);
                if (X3) {
                    bulk_copy_g2s(dst + Q_HALF, qs + qhalf, qb, barFull + 8 * st);
                    bulk_copy_g2s(dst + 2 * Q_HALF + kb, ks_ + qhalf, kb, barFull + 8 * st);
                }
            }
            int it = 0;
            for (int c = 0; c < nchunk; ++c)
                for (int ss = 0; ss < nss; ss += VP, ++it) {
                    const int st = it % NST, n = min(VP, nss - ss);
                    mbar_wait(barEmpty + 8 * st, ((phE >> st) & 1) ^ 1);
                    phE ^= 1u << st;
                    const uint8_t* src = gv + ((size_t)c * (T >> 5) + (key0 >> 5) + ss) * 2 * V_HALF;
                    if (X3) {   // hi and lo of n consecutive 32-key stages are contiguous in the V image: one copy
                        mbar_arrive_expect_tx(barFull + 8 * st, (uint32_t)n * 2 * V_HALF);
                        bulk_copy_g2s(slot_addr(st), src, (uint32_t)n * 2 * V_HALF, barFull + 8 * st);
                    } else {
                        mbar_arrive_expect_tx(barFull + 8 * st, (uint32_t)n * V_HALF);
                        for (int u = 0; u < n; ++u) bulk_copy_g2s(slot_addr(st) + u * 2 * V_HALF, src + (size_t)u * 2 * V_HALF, V_HALF, barFull + 8 * st);
                    }
                }
        }
    } else {
        // ================================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idescS = make_idesc(0u, 128u, (uint32_t)Tk);
            constexpr uint32_t idescO = make_idesc(0u, 128u, 128u, 1u);   // B (= V) is MN-major
            uint32_t phF = 0;   // bit i: parity of the next wait on slot i's "full" barrier
            uint32_t acc = 0, acc2 = 0;
            for (int ks = 0; ks < nks; ++ks) {
                const int st = ks % NSS;
                mbar_wait(barFull + 8 * st, (phF >> st) & 1);
                phF ^= 1u << st;
                tc_fence_after();
                const uint32_t q_hi = slot_addr(st), q_lo = q_hi + Q_HALF, k_hi = q_hi + 2 * Q_HALF, k_lo = k_hi + Tk * 64;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t dah = make_desc(q_hi + kk * 256, 128, 512), dbh = make_desc(k_hi + kk * 256, 128, 512);
                    umma_bf16(tmem, dah, dbh, idescS, acc);
                    acc = 1;
                    if (X3) {
                        const uint64_t dal = make_desc(q_lo + kk * 256, 128, 512), dbl = make_desc(k_lo + kk * 256, 128, 512);
                        umma_bf16(tmem + 256, dah, dbl, idescS, acc2);
                        umma_bf16(tmem + 256, dal, dbh, idescS, 1);
                        acc2 = 1;
                    }
                }
                umma_commit(barEmpty + 8 * st);
            }
            umma_commit(barS);
            mbar_wait(barP, 0);
            tc_fence_after();
            int it = DIRECT ? nks : 0;   // (in-kernel-split form: its producers walk one slot sequence through both phases)
            for (int c = 0; c < nchunk; ++c) {
                const int buf = c & 1;
                mbar_wait(barOempty + 8 * buf, ((c >> 1) & 1) ^ 1);
                tc_fence_after();
                uint32_t a0 = 0, a1 = 0;
                for (int ss = 0; ss < nss; ss += VP, ++it) {
                    const int st = it % NST, n = min(VP, nss - ss);
                    mbar_wait(barFull + 8 * st, (phF >> st) & 1);
                    phF ^= 1u << st;
                    tc_fence_after();
                    for (int u = 0; u < n; ++u) {
                        const uint32_t v_hi = slot_addr(st) + u * 2 * V_HALF, v_lo = v_hi + V_HALF;
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint32_t po = (uint32_t)((ss + u) * 4 + kk * 2) * 2048;
                            const uint64_t dah = make_desc(sP + po, 2048, 128), dbh = make_desc(v_hi + kk * 4096, 2048, 128);
                            umma_bf16(tmem + buf * 256, dah, dbh, idescO, a0);
                            a0 = 1;
                            if (X3) {
                                const uint64_t dal = make_desc(sP + PH + po, 2048, 128), dbl = make_desc(v_lo + kk * 4096, 2048, 128);
                                umma_bf16(tmem + buf * 256 + 128, dah, dbl, idescO, a1);
                                umma_bf16(tmem + buf * 256 + 128, dal, dbh, idescO, 1);
                                a1 = 1;
                            }
                        }
                    }
                    umma_commit(barEmpty + 8 * st);
                }
                umma_commit(barOfull + 8 * buf);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, 512);
}

}  // namespace

// out[b][t][h*ch + c] = sum_kb w_kb O_kb / sum_kb w_kb,  w_kb = exp2(m_kb - max_kb m_kb) * l_kb  (the key blocks' partial softmaxes merged)
__global__ void attn_merge_kernel(const float* __restrict__ part, const float2* __restrict__ ml, float* __restrict__ out, int nkb, int B, int T,
                                  int H, int ch, size_t total4) {
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= total4) return;
    const size_t i = i4 * 4, C = (size_t)H * ch;
    const int c = (int)(i % C), h = c / ch;
    const size_t bt = i / C;
    const int t = (int)(bt % T);
    const size_t b = bt / T;
    float m = -INFINITY, w[8];
    for (int kb = 0; kb < nkb; ++kb) m = fmaxf(m, ml[(((size_t)kb * B + b) * H + h) * T + t].x);
    float W = 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
        const float2 v = ml[(((size_t)kb * B + b) * H + h) * T + t];
        w[kb] = exp2f(v.x - m) * v.y;
        W += w[kb];
    }
    const float iW = 1.0f / W;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kb = 0; kb < nkb; ++kb) {
        const float4 pv = *reinterpret_cast<const float4*>(part + (size_t)kb * B * T * C + i);
        const float f = w[kb] * iW;
        o.x = fmaf(f, pv.x, o.x); o.y = fmaf(f, pv.y, o.y); o.z = fmaf(f, pv.z, o.z); o.w = fmaf(f, pv.w, o.w);
    }
    *reinterpret_cast<float4*>(out + i) = o;
}

bool attn_direct_eligible(int T, int ch) { return attn_tc_eligible(T, ch) && (128 + T) * 4 <= 7 * NUM_THREADS; }
// keys per CTA: the whole sequence up to 256 (one score tile in TMEM); longer sequences in blocks merged afterwards
int attn_tc_key_block(int T) {
    if (T <= 256) return T;
    for (int tk : {256, 192, 128}) if (T % tk == 0 && T / tk <= 8) return tk;
    return 0;
}
bool attn_tc_eligible(int T, int ch) {
    return T >= 32 && T % 32 == 0 && ch >= 128 && ch % 128 == 0 && attn_tc_key_block(T) > 0 && (T <= 256 || T % 16 == 0);
}
size_t attn_tc_scratch_bytes(int B, int T, int H, int ch) {
    const int tk = attn_tc_key_block(T);
    if (tk <= 0 || tk == T) return 0;
    const size_t nkb = (size_t)(T / tk);
    return nkb * B * T * ((size_t)H * ch * sizeof(float) + (size_t)H * sizeof(float2));
}
size_t attn_qkv16_bytes(int B, int T, int H, int ch) { return (size_t)B * H * 12 * ch * T; }

cudaError_t launch_qkv_split(const float* qkv, uint8_t* dst, int B, int T, int H, int ch, cudaStream_t st, int* range_flag) {
    const size_t total = (size_t)B * H * 3 * (ch / 8) * T;
    if (!total) return cudaSuccess;
    qkv_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(qkv, dst, T, H, ch, total, range_flag);
    g_launch_count += 1;
    return cudaGetLastError();
}

cudaError_t launch_attention_tc(const AttnTcParams& p_in, bool x3, cudaStream_t st) {
    if (p_in.B <= 0) return cudaSuccess;
    AttnTcParams p = p_in;
    const int Tk = attn_tc_key_block(p.T);
    if (Tk <= 0) return cudaErrorInvalidValue;
    const int nkb = p.T / Tk;
    if (nkb > 1) {   // key blocks + merge: fp32 output only, pre-split operand images, scratch from attn_tc_scratch_bytes
        if (!p.part_out || !p.out || p.out_u || p.qkv32) return cudaErrorInvalidValue;
        p.Tk = Tk;
        p.part_ml = reinterpret_cast<float2*>(p.part_out + (size_t)nkb * p.B * p.T * ((size_t)p.H * p.ch));
    } else p.Tk = 0;
    const int smem = attn_stages(Tk) * stage_bytes(Tk) + 2 * p_half_bytes(Tk) + 256 + 2048;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid((unsigned)((p.T + 127) / 128 * nkb), p.H, p.B);
    if (p.qkv32) {
        if ((128 + p.T) * 4 > 7 * NUM_THREADS) return cudaErrorInvalidValue;   // producer item budget: T <= 208 (attn_direct_eligible)
        if (x3) attn_tc_kernel<true, true><<<grid, 2 * NUM_THREADS, smem, st>>>(p);
        else attn_tc_kernel<false, true><<<grid, 2 * NUM_THREADS, smem, st>>>(p);
    } else if (x3) attn_tc_kernel<true, false><<<grid, ATTN_THREADS, smem, st>>>(p);
    else attn_tc_kernel<false, false><<<grid, ATTN_THREADS, smem, st>>>(p);
    g_launch_count += 1;
    if (nkb > 1) {
        const size_t total4 = (size_t)p.B * p.T * p.H * p.ch / 4;
        attn_merge_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(p.part_out, p.part_ml, p.out, nkb, p.B, p.T, p.H, p.ch, total4);
        g_launch_count += 1;
    }
    return cudaGetLastError();
}

}  // namespace eegldm
