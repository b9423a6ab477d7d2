// tcgen05 implicit-GEMM convolution for sm_100a (the tensor-pipe path of the eegldm engine).
//
// Computes, for channels-last activations,
//     out[b,t,co] = bias[co] + temb[b,co] + res[b,t,co] + sum_seg sum_{k,ci} W_seg[co,ci,k] * u_seg[b, t+k-pad, ci]
//     u = resample(silu?(scale*x + shift))           (GroupNorm apply + SiLU + AvgPool/nearest)
// i.e. the same contract as conv_simt_kernel (reference: src/models/unet.py:263,291,302,158,161 + 308-327),
// in one launch (fused-producer form, the default) or two (pre-pass form: AvgPool inputs, the qkv conv):
//
//   act_split_kernel   (pre-pass form only) one pass over x: apply the GroupNorm affine / SiLU / resample, split every fp32 value into
//                      16-bit parts and write them as ready-made shared-memory tile images ("U" tensors).
//   conv_tc_kernel     tcgen05 GEMM: M = 128 output positions per CTA, N = 128 or 256 output channels, K = taps*Cin; weights arrive by
//                      cp.async.bulk (single CTA) or cp.async.bulk.tensor (CTA pairs), activations from six producer warps that read
//                      the fp32 source themselves (or by bulk copy of the U image), D lives in TMEM.
//
//   * fp32 parity on a 16-bit tensor pipe ("f16x3"): every fp32 operand is split x = hi + lo/2048 with hi and lo
//     fp16 (11 + 11 significand bits, lo pre-scaled by 2^11 so it stays in fp16's normal range) and three products
//     are issued per K-slice: hi*hi into accumulator 0, hi*lo + lo*hi into accumulator 1; the epilogue returns
//     acc0 + acc1 * 2^-11.  Operand error ~2^-22; the tensor core's fp32 accumulate truncates (measured,
//     tools/conv_precision.py), which the separate correction accumulator keeps off the long chain.
//     EEGLDM_MATH_BF16_TC issues a single bf16 product (fast, NOT a parity mode).
//   * B operand = weights, pre-split and pre-packed on the host into the exact shared-memory image of one
//     pipeline stage (K-major, no swizzle, 8x16-byte core matrices, [k-step][tap][hi|lo][Cout/8]: the columns of any tile
//     width, or of one CTA of a pair, are one contiguous slice = one copy).
//   * A operand = activations in a "phase-strided halo" K-major layout: the 128 M rows of a tile are 8 segments of
//     16 consecutive positions; the 8 rows of one core matrix are the SAME offset in the 8 segments, and each
//     segment carries its own 2 halo positions (18 slots).  A tap shift of +-1 position is then a whole-core-matrix
//     shift = +-SBO bytes on the descriptor start address, so the 3 taps of a k=3 conv read ONE staged tile
//     (12.5 % halo overhead instead of 3 copies).  Segments may belong to different samples; halos at sample
//     edges are zero (the conv's padding).  Stage image: [hi|lo][kc][slot][segment][8 ch], 18 KB per (tile, k-step).
//   * persistent, warp-specialised: one CTA per SM; 8 epilogue warps in two warpgroups (TMEM -> registers -> global; 4 in the
//     one-warpgroup form kept for 32-channel GroupNorm groups), six producer warps (fused-producer form), one loader warp, one
//     MMA warp.  Rings: 5 x 18 KB activation stages + 96 KB of weight stages (3 x 32 KB ... 12 x 8 KB).  N = 128: two TMEM
//     accumulator sets, the epilogue of tile i overlaps the mainloop of tile i+1; N = 256 in f16x3 fills all 512 TMEM columns
//     (one set).  The 256-wide launches run as cta_group::2 CTA pairs (M = 256 per MMA), see the kernel comment.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstring>
#include <vector>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace eegldm {
using namespace tc;
namespace {

constexpr int BM = 128;            // output positions per CTA (8 segments x 16)
constexpr int BK = TC_BK;          // input channels per k-step
constexpr int SLOTS = 18;          // 16 positions + 2 halo slots per segment
constexpr int A_SBO = 128;         // bytes between core matrices along M (between slots)
constexpr int A_LBO = SLOTS * A_SBO;   // bytes between core matrices along K (8-channel chunks)
constexpr int A_TILE = (BK / 8) * A_LBO;   // one hi (or lo) activation tile: 9216 B
constexpr int A_STAGE = 2 * A_TILE;
static_assert(A_TILE == TC_U_HALF_BYTES, "engine and kernel disagree on the U tile size");
constexpr int B_LBO = 128;         // weight image: bytes between the 8-channel K chunks of one 8-column group
constexpr int B_SBO = (BK / 8) * B_LBO;    // bytes between 8-column groups (512): any N tile is a contiguous slice
constexpr int N_ITEMS = 8 * SLOTS * (BK / 8);   // 576 16-byte items per activation tile
// shared-memory plan: activation ring | weight ring | mbarriers (<= 8*(2*6 + 2*16 + 4) = 384 B), TMEM slot at +448 |
// 4 warps x [32][33] fp32 epilogue transpose buffers.  A CTA pair stages half the weight bytes per MMA, so it trades
// weight-ring bytes for two more activation stages (the activation stream comes from HBM: latency x bandwidth).
// alt = CTA-pair form or fused-producer form: one more activation stage, one 32 KB weight stage less (the producer's output is
// burstier than a bulk copy; measured 1-2 % faster than 4 + 128 KB for the fused-producer convs)
__host__ __device__ constexpr int ring_na(bool alt) { return alt ? 5 : 4; }                          // 18 KB stages
__host__ __device__ constexpr int ring_b_bytes(bool alt) { return (alt ? 96 : 128) * 1024; }         // 3 x 32 KB ... 12 x 8 KB stages
__host__ __device__ constexpr int bar_off(bool alt) { return ring_na(alt) * A_STAGE + ring_b_bytes(alt); }
constexpr int STG_LD = 36;         // staging row stride in floats: 16-byte aligned rows, conflict-free 128-bit writes and reads
constexpr int STAGING_BYTES = 4 * 32 * STG_LD * 4;
constexpr int STATS_BYTES = 8 * 256 * 8;   // per epilogue warp (8 in the two-warpgroup form): (mean, M2) of up to 8 segments x 32 groups
__host__ __device__ constexpr int smem_bytes(bool alt) { return bar_off(alt) + 512 + STAGING_BYTES + STATS_BYTES; }
constexpr int NUM_THREADS = 192;
constexpr int SPLIT_THREADS = 192;
static_assert(N_ITEMS % SPLIT_THREADS == 0, "items must divide evenly over the act_split block");

__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.f + __expf(-v)); }
// prologue activation after the affine: 0 none, 1 SiLU, 2 LeakyReLU(0.2) (PatchDiscriminator blocks)
__device__ __forceinline__ float act(float x, float a, float s, int silu) {
    const float v = fmaf(a, x, s);
    return silu == 1 ? silu_fast(v) : (silu == 2 ? (v > 0.f ? v : 0.2f * v) : v);
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// 8 consecutive floats (32-byte aligned) as ONE 256-bit load (LDG.E.256, sm_100): the fused producer's items are 8 channels wide,
// and two 128-bit loads per item each fetch half of every 32-byte sector they touch -- twice the LSU wavefronts (ncu: 67 % LSU
// utilisation, profiles/r02b_ncu_full_conv_pair_512.txt)
__device__ __forceinline__ void ldg8(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
// SiLU on the bare special-function instructions: v * rcp(1 + ex2(-v * log2 e)).  ex2.approx.ftz needs none of __expf's
// denormal-range fix-ups (3 extra instructions per value): an underflowing exponential is 0 (v large: silu = v), an
// overflowing one is +inf (v very negative: rcp = 0, silu = -0); both approximations are within 2 ulp, as before.
__device__ __forceinline__ float silu_ftz(float v) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return v * r;
}
// f16x3 operand range: hi = fp16(v) is inf from 65520 up and lo = (v - inf) * 2^11 is NaN -- the split is only meaningful for
// |v| < 65504 (include/eegldm.h).  true when any of the 8 values is outside.
__device__ __forceinline__ bool out_of_f16_range(const float (&v)[8]) {
    // 3-input max with |.| source modifiers: 4 instructions per 8 values.  (A NaN operand is not an overflow: it reaches the
    // output as NaN in every math mode.)
    const float m = fmaxf(fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))),
                          fmaxf(fmaxf(fabsf(v[4]), fabsf(v[5])), fmaxf(fabsf(v[6]), fabsf(v[7]))));
    return !(m < 65504.f);
}

// ------------------------------------------------------------------------------------------------ act_split
// grid (nks, n_mtiles), block 192.  Item (q, c, r): slot q of segment r, 8-channel chunk c of k-step blockIdx.x.
// Lanes run over r fastest (8 segments = one 128-byte line of the image), then c: full-line stores, and the 4 lanes
// of one (q, r) read 128 contiguous bytes of one position.
template <bool X3>
__global__ void __launch_bounds__(SPLIT_THREADS) act_split_kernel(const ActSplitParams p) {
    const int ks = blockIdx.x, m_tile = blockIdx.y;
    const int spt = p.Tout >> 4;
    uint8_t* img = p.U + ((size_t)m_tile * p.nks + ks) * A_STAGE;
    uint8_t* img2 = p.U_raw ? p.U_raw + ((size_t)m_tile * p.nks + ks) * A_STAGE : nullptr;   // raw (no affine / SiLU) twin
    const int Cin = p.C0 + p.C1;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < N_ITEMS / SPLIT_THREADS; ++j) {
        const int idx = threadIdx.x + SPLIT_THREADS * j;
        const int r = idx & 7, c = (idx >> 3) & 3, q = idx >> 5;
        const int g = m_tile * 8 + r;
        const bool segv = g < p.nsegs16;
        const int b = segv ? g / spt : 0;
        const int t = (g % spt) * 16 + q - 1;
        const bool inb = segv && t >= 0 && t < p.Tout;
        float v[8], vr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = vr[e] = 0.f;   // conv zero padding / rows past the batch
        if (inb) {
            const int cc = ks * BK + c * 8;
            const float* src; int ch, Cs;
            if (cc < p.C0) { src = p.src0; ch = cc; Cs = p.C0; } else { src = p.src1; ch = cc - p.C0; Cs = p.C1; }
            float a[8], s[8];
            if (p.scale) {
                const size_t o = (size_t)b * Cin + cc;
                const float4 a0 = ldg4(p.scale + o), a1 = ldg4(p.scale + o + 4), s0 = ldg4(p.shift + o), s1 = ldg4(p.shift + o + 4);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) { a[e] = 1.f; s[e] = 0.f; }
            }
            const float* base = src + (size_t)b * p.Tin * Cs + ch;
            if (p.resample == RS_AVGPOOL2) {
                const float* r0 = base + (size_t)(2 * t) * Cs;
                const float4 x0 = ldg4(r0), x1 = ldg4(r0 + 4), y0 = ldg4(r0 + Cs), y1 = ldg4(r0 + Cs + 4);
                const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const float ya[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    v[e] = 0.5f * (act(xa[e], a[e], s[e], p.silu) + act(ya[e], a[e], s[e], p.silu));
                    vr[e] = 0.5f * (xa[e] + ya[e]);
                }
            } else {
                const float* r0 = base + (size_t)(p.resample == RS_NEAREST2 ? (t >> 1) : t) * Cs;
                const float4 x0 = ldg4(r0), x1 = ldg4(r0 + 4);
                const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) { v[e] = act(xa[e], a[e], s[e], p.silu); vr[e] = xa[e]; }
            }
        }
        const uint32_t off = (uint32_t)(c * A_LBO + q * A_SBO + r * 16);
        uint4 hi, lo;
        if (X3) {
            bad |= out_of_f16_range(v);
            split8_f16(v, hi, lo);
            *reinterpret_cast<uint4*>(img + A_TILE + off) = lo;
        } else round8_bf16(v, hi);
        *reinterpret_cast<uint4*>(img + off) = hi;
        if (img2) {
            if (X3) {
                bad |= out_of_f16_range(vr);
                split8_f16(vr, hi, lo);
                *reinterpret_cast<uint4*>(img2 + A_TILE + off) = lo;
            } else round8_bf16(vr, hi);
            *reinterpret_cast<uint4*>(img2 + off) = hi;
        }
    }
    if (X3 && bad && p.range_flag) atomicOr(p.range_flag, 1);
}

// ------------------------------------------------------------------------------------------------ conv_tc
// Persistent: one CTA per SM walks tiles (n_tile fastest, so concurrently running CTAs share activation tiles in L2);
// two TMEM accumulator sets, so the epilogue of tile i overlaps the mainloop of tile i+1.
// CL CTAs of a cluster work on CL consecutive M tiles of the same N tile: the weight stage is identical for all of them,
// so each CTA fetches 1/CL of it and multicasts it to the whole cluster (L2 -> SM weight traffic / CL).
//
// PAIR (cta_group::2): the two CTAs of a cluster form one M=256 x N=BN MMA.  Each CTA stages its own 128-row activation
// tile and only ITS HALF of the weight columns (BN/2), the leader (cluster rank 0) issues every MMA for both, and each
// CTA drains its own 128 TMEM lanes.  Per MMA an SM then reads 4 KB of A + BN/2 x 32 B of B instead of BN x 32 B, and
// stages half the weight bytes: the shared-memory operand traffic that bounds the single-CTA shape (profiles/r01_summary.md)
// drops below the tensor pipe's time, and N = 128 (two TMEM accumulator sets: overlapped epilogue) costs no more operand
// bandwidth than the single-CTA N = 256 shape.  Synchronisation without a relay hop: every "full" barrier lives in the LEADER.
// Both CTAs load with cp.async.bulk.tensor.cta_group::2 (tensor maps over the weight / U images), whose complete_tx reaches the
// leader's barrier from either CTA (only the leader arrives, with the pair's byte count); the fused producers of both CTAs
// arrive on the leader's barrier through the cluster (one elected lane per warp after __syncwarp); the leader's tcgen05.commit
// multicasts the "empty" / "accumulator full" arrivals to both CTAs; both CTAs' epilogue threads arrive on the leader's
// "accumulator empty".  (Round 1's form relayed the peer's own full barriers through a warp: 25-45 % slower.)
//
// DIRECT: no activation pre-pass.  Six producer warps (the act_split thread mapping: 192 threads x 3 sixteen-byte items per
// k-step) read the fp32 source rows themselves, apply the GroupNorm affine / SiLU / nearest-x2, split to fp16 hi/lo and
// write the stage image into the activation ring with st.shared (+ fence.proxy.async, 192 arrivals on the "full" barrier),
// one k-step of global loads in flight in registers ahead of the stage being written.  The U tensors -- 8.5 B per element
// of HBM traffic in act_split plus 4.5 B per element read here -- disappear; the conv reads 4 B per element instead.
// Registers: launched at 384 threads (168 registers), then setmaxnreg moves registers from the producer / loader / MMA
// warpgroups (136) to the epilogue warpgroup (232): 128*232 + 256*136 = 384*168 exactly -- the CTA's register pool is fixed at
// launch, and a split that needs more makes setmaxnreg.inc wait forever.  The producer is bounded by the XU pipe (ex2 + rcp
// of SiLU, fp16 pack / unpack: ~1000 cycles per k-step), which hides under a 3-tap N=256 k-step (2304 tensor cycles) but
// not under 1-tap or N=128 k-steps; a two-k-step prefetch distance measured no different from one, and nine producer warps
// with two items each (512 threads, 232 / 88 registers) measured 6 % slower than six with three.  tools/producer_bench.py
// (profiles/r01_producer_bench.txt) switches parts of the producer off: no single part dominates.  Interleaving the 1-tap
// k-steps of a skip_connection segment with the 3-tap ones (so the ring averages their tensor time) measured 10 % SLOWER on the
// two-segment layers and was dropped.  Pulling the rows of the CTA's NEXT tile into L2 when a tile starts (prefetch.global.L2 per 128-byte line;
// for the short-K level-0 tiles, whose k-step is shorter than a DRAM round trip) measured 5 % SLOWER on those layers (0.31 vs 0.29 ms):
// spreading the same prefetches over the k-steps (one 128-byte row piece per k-step, a constant element offset from the current rows)
// measured no different (+-1 %).  Those tiles are not waiting for DRAM: per tile the MMAs read 480 KB of operands from shared memory
// (N = 128: 20 KB per K-slice pair), the copies and the producer write 264 KB and the overlapped epilogue stages 128 KB -- 870 KB at
// 128 B/clk is 6.8 k of the 11.7 k cycles a tile takes; the shared-memory pipe, which the three f16x3 products keep busy, bounds them.
//
// EPI8: two epilogue warpgroups instead of one.  With N = 256 in f16x3 the two accumulators fill TMEM, so the epilogue of a tile cannot
// overlap the next mainloop and its duration is lost tensor time (10-17 k cycles per tile, 20-40 % on the K <= 1536 layers:
// profiles/r02_summary.md).  One warp per scheduler is latency-bound (every TMEM load -> transpose -> global store chain is exposed);
// two warps per scheduler, each draining half of the tile's columns in 16-column chunks, halve that time.  Thread t of warp w
// (lane quarter ew = w & 3, column half eh = w >> 2) owns 4 consecutive positions (ew*4 .. +3) of ONE 16-position segment (lane / 4)
// x 4 channels per chunk.
constexpr int conv_tc_threads(bool direct, bool epi8) { return direct ? (epi8 ? 512 : 384) : (epi8 ? 320 : NUM_THREADS); }
template <bool X3, int BN, int CL, bool PAIR, bool DIRECT, bool EPI8 = false>
__global__ void __launch_bounds__(conv_tc_threads(DIRECT, EPI8), 1) conv_tc_kernel(const __grid_constant__ TcConvParams p) {
    static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of 2");
    static_assert(!PAIR || EPI8, "the CTA-pair form uses the two-warpgroup epilogue");
    constexpr int NEPI = EPI8 ? 8 : 4;               // epilogue warps 0 .. NEPI-1
    constexpr int NPROD = DIRECT ? 6 : 0;            // producer warps NEPI .. NEPI+NPROD-1
    constexpr int W_LOAD = NEPI + NPROD, W_MMA = NEPI + 1 + NPROD;
    constexpr int BNL = PAIR ? BN / 2 : BN;    // weight columns staged in THIS CTA's shared memory
    constexpr int B_HALF = BNL * BK * 2;       // hi (or lo) weight tile of one (tap, k-step): BNL x 32 x 2 B
    constexpr int B_STAGE = 2 * B_HALF;
    constexpr bool ALT = true;   // ring plan of every form: 5 activation stages + 96 KB of weight stages (the 4 + 128 KB plan of the
                                 // single-CTA pre-pass form no longer fits next to the 16 KB statistics buffer)
    constexpr int NA = ring_na(ALT), BAR_OFF = bar_off(ALT), STAGING_OFF = BAR_OFF + 512;
    constexpr int NB = ring_b_bytes(ALT) / B_STAGE;
    static_assert(8 * (3 * NA + 2 * NB + 4) <= 448, "barrier area: 3 NA + 2 NB + 4 mbarriers in front of the TMEM slot");
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sA = sbase, sB = sbase + NA * A_STAGE;
    const uint32_t bars = sbase + BAR_OFF;             // 8-byte mbarriers
    const uint32_t barAfull = bars, barAempty = bars + 8 * NA, barBfull = bars + 16 * NA, barBempty = barBfull + 8 * NB,
                   barAccFull = barBempty + 8 * NB, barAccEmpty = barAccFull + 16, barApeer = barAccEmpty + 16;   // barApeer: pair + fused producer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 448);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nks0 = p.seg[0].nks, nks = nks0 + (p.nseg > 1 ? p.seg[1].nks : 0);
    const int spt = p.Tout >> 4;   // 16-position segments per sample
    const int n_ntiles = p.Cout / BN;
    const int n_mtiles = (p.nsegs16 + 7) / 8;
    const int nwork = n_ntiles * ((n_mtiles + CL - 1) / CL);   // work item = (group of CL M tiles, N tile), per cluster
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int cid = CL > 1 ? (int)cluster_id_x() : (int)blockIdx.x;
    const int ncl = CL > 1 ? (int)num_clusters_x() : (int)gridDim.x;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

    const bool leader = !PAIR || crank == 0;
    if (tid == 0) {
        // pair: the "full" barriers of the async loads that count are the leader's (its one arrive.expect_tx for the pair's bytes);
        // fused producers arrive on their own CTA's barAfull, the peer's idle issuer warp forwards that to the leader's barApeer
        for (int i = 0; i < NA; ++i) { mbar_init(barAfull + 8 * i, DIRECT ? 32 * NPROD : 1); mbar_init(barAempty + 8 * i, 1); mbar_init(barApeer + 8 * i, 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(barBfull + 8 * i, 1); mbar_init(barBempty + 8 * i, PAIR ? 1 : CL); }
        for (int i = 0; i < 2; ++i) { mbar_init(barAccFull + 8 * i, 1); mbar_init(barAccEmpty + 8 * i, (PAIR ? 2 : 1) * 32 * NEPI); }
        fence_mbar_init();
    }
    constexpr uint32_t ACC_COLS = X3 ? 2 * BN : BN;    // X3: accumulator 0 = hi*hi, accumulator 1 = cross terms * 2^11
    constexpr int NSETS = ACC_COLS * 2 <= 512 ? 2 : 1; // N=256 in f16x3 fills TMEM: no epilogue overlap (used for long K only)
    constexpr uint32_t TMEM_COLS = NSETS * ACC_COLS;
    constexpr uint32_t IDESC = make_idesc(X3 ? 0u : 1u, PAIR ? 2 * BM : BM, BN);
    constexpr bool CAT = X3 && BN == 128 && !PAIR;                 // hi x [hi | lo] as one N = 256 MMA (see the MMA issuer)
    constexpr uint32_t IDESC_CAT = make_idesc(0u, BM, 2 * BN);
    static_assert(!CAT || B_HALF == (BN / 8) * B_SBO, "the lo weight tile must continue the hi tile's 8-column group stride");
    if (warp == W_MMA) {
        if (PAIR) tmem_alloc2(smem_u32(tmem_slot), TMEM_COLS);
        else tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // peers' mbarriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // eegldm_bench_conv_timeline: cycle counters per warp role (one writer each: epilogue warp 0, producer warp 4, loader, MMA warp)
    const bool tl = p.timeline != nullptr;
    unsigned long long* tlo = tl ? p.timeline + (size_t)blockIdx.x * TC_TL_N : nullptr;
    long long tl_a = 0, tl_b = 0, tl_c = 0;
    const long long tl_start = tl ? clock64() : 0;
#define TL_WAIT(acc, stmt) do { if (tl) { const long long t0__ = clock64(); stmt; acc += clock64() - t0__; } else { stmt; } } while (0)
    if (DIRECT && !EPI8) {   // warpgroup 0 = epilogue, warpgroups 1-2 = producers + loader + MMA issuer
        if (warp < 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 136;");
    }
    if (DIRECT && EPI8) {    // 512 threads x 128 registers at launch: 256 x 136 (epilogue) + 256 x 120 (producers, loader, MMA issuer)
        if (warp < 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");
    }

    if (EPI8 && warp < NEPI) {
        // ================================================================ epilogue, two warpgroups (see the kernel comment)
        const int ew = warp & 3, eh = warp >> 2;
        constexpr int HC = BN / 2;                                   // columns per warpgroup
        // per-warp transpose buffer [32 rows][16 columns], 16-byte quads XOR-swizzled by (row >> 1) & 3: conflict-free 128-bit
        // writes (thread = row) and reads (thread = (segment, quad)) without padding
        uint32_t* stg = reinterpret_cast<uint32_t*>(smem + STAGING_OFF) + warp * 512;
        float2* stats = reinterpret_cast<float2*>(smem + STAGING_OFF + STAGING_BYTES);   // [warp][256]: segment * gpt + group
        const int cpg = p.gn_cpg, gpt = p.gn_partial ? HC / cpg : 0;                        // groups per warpgroup half-tile (<= 32)
        const int cpg_sh = 31 - __clz(max(cpg, 1));
        const int seg = lane >> 2, quad = lane & 3, col4 = quad * 4;
        int lt = 0;
        for (int w = cid; w < nwork; w += ncl, ++lt) {
            const int n_tile = w % n_ntiles, m_tile = (w / n_ntiles) * CL + crank;
            const int as = lt % NSETS, use = lt / NSETS;
            const int co0 = n_tile * BN + eh * HC;
            if (X3 && p.qkv16) {
                // qkv conv of an AttentionBlock (unet.py:158): the output's only consumer is attn_tc.cu, so q, k, v go straight into
                // its fp16 hi/lo operand images (layout: qkv_split_kernel) -- no fp32 tensor, no qkv_split pass.  A TMEM lane is one
                // position and 16 consecutive columns are two 8-channel image items of 16 bytes each (x hi, lo): no shared-memory
                // transpose.  The rows of all three images are in this tile's phase-strided order (attn_tc.cu, attn_q_row: queries AND
                // keys -- the attention only needs K and V to agree on the key order), so lane = (slot, segment) with the segment
                // fastest and eight consecutive lanes store one 128-byte core-matrix row set.  (In position order the four lanes
                // l, l+8, l+16, l+24 of a segment hold consecutive rows: 16-byte pieces of eight lines per store cost 5.8 k cycles per
                // tile, a 16-shuffle lane permutation to 64-byte runs 0.9 k + 1.2 k, this order 0.9 k: tools/conv_timeline.py.)
                const int ch = p.qkv_ch, T = p.Tout;
                const int cq = co0 / (3 * ch), rem = co0 - cq * 3 * ch, which = rem / ch, c0 = rem - which * ch;   // head, q|k|v, channel
                const int sg = lane & 7, pos = ew * 4 + (lane >> 3);   // (segment, position) of this lane's TMEM row
                const int g = m_tile * 8 + sg;
                const bool valid = g < p.nsegs16;
                const int b = valid ? g / spt : 0, s_in = g % spt;
                // image row of this position (attn_tc.cu, attn_q_row): tile s_in / 8 holds nseg segments, rows run over them first
                const int nseg = min(8, spt - (s_in & ~7)), t = (s_in >> 3) * 128 + nseg * pos + (s_in & 7);
                const size_t plane = (size_t)4 * ch * T;
                const size_t half = which < 2 ? (size_t)T * 64 : 8192;
                uint8_t* base = p.qkv16 + (((size_t)b * p.qkv_H + cq) * 3 + which) * plane +
                                (which < 2 ? (size_t)(t >> 3) * 512 + (t & 7) * 16
                                           : (size_t)(t >> 5) * 2 * half + ((t & 31) >> 3) * 2048 + (t & 7) * 16);
                const float* bias_p = p.bias && !(p.debug & 512) ? p.bias + co0 : nullptr;
                const bool do_store = valid && !(p.debug & 256);   // (debug bits 256 / 512: timing experiments, tools/conv_timeline.py)
                TL_WAIT(tl_a, mbar_wait(barAccFull + 8 * as, use & 1));
                tc_fence_after();
                const long long tl_e0 = tl ? clock64() : 0;
                const uint32_t acc_addr = tmem + ((uint32_t)(ew * 32) << 16) + as * ACC_COLS + eh * HC;
                uint32_t vn[16], c2n[16];
                tmem_ld16_async(acc_addr, vn);
                tmem_ld16_async(acc_addr + (uint32_t)BN, c2n);
                bool bad = false;
#pragma unroll 1
                for (int cb = 0; cb < HC; cb += 16) {
                    float4 bi[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bi[j] = bias_p ? ldg4(bias_p + cb + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    tmem_ld_wait();
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = fmaf(__uint_as_float(c2n[i]), 1.0f / LO_SCALE, __uint_as_float(vn[i]));
                    if (cb + 16 < HC) {
                        tmem_ld16_async(acc_addr + (uint32_t)(cb + 16), vn);
                        tmem_ld16_async(acc_addr + (uint32_t)(BN + cb + 16), c2n);
                    } else {
                        tc_fence_before();
                        if (PAIR) mbar_arrive_cluster(barAccEmpty + 8 * as, 0);
                        else mbar_arrive(barAccEmpty + 8 * as);
                    }
                    const int c = c0 + cb;
                    uint8_t* dst = base + (which < 2 ? (size_t)(c >> 5) * 2 * half + ((c & 31) >> 3) * 128
                                                     : (size_t)(c >> 7) * (T >> 5) * 2 * half + ((c & 127) >> 3) * 128);
                    uint4 pc[4];   // hi item 0, lo item 0, hi item 1, lo item 1 of this lane's TMEM row
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const float v8[8] = {f[8 * it] + bi[2 * it].x,     f[8 * it + 1] + bi[2 * it].y,     f[8 * it + 2] + bi[2 * it].z,
                                             f[8 * it + 3] + bi[2 * it].w, f[8 * it + 4] + bi[2 * it + 1].x, f[8 * it + 5] + bi[2 * it + 1].y,
                                             f[8 * it + 6] + bi[2 * it + 1].z, f[8 * it + 7] + bi[2 * it + 1].w};
                        bad |= out_of_f16_range(v8);
                        split8_f16(v8, pc[2 * it], pc[2 * it + 1]);
                    }
                    if (do_store) {
                        *reinterpret_cast<uint4*>(dst) = pc[0];
                        *reinterpret_cast<uint4*>(dst + half) = pc[1];
                        *reinterpret_cast<uint4*>(dst + 128) = pc[2];
                        *reinterpret_cast<uint4*>(dst + 128 + half) = pc[3];
                    }
                }
                if (bad && p.range_flag) atomicOr(p.range_flag, 1);
                if (tl) tl_b += clock64() - tl_e0;
                continue;
            }
            const int g = m_tile * 8 + seg;
            const int rb = g < p.nsegs16 ? g / spt : -1;             // sample (< 0: segment past the batch)
            const int rt = (g % spt) * 16 + ew * 4;                  // first of this thread's 4 positions
            const int bb = max(rb, 0);
            // polyphase up-conv (TcConvParams.poly): this half tile is one output phase; Cr real channels, rows 2t + phase
            const int Cr = p.poly ? p.Cout >> 1 : p.Cout;
            const int ph = p.poly && co0 >= Cr ? 1 : 0, cor = co0 - ph * Cr;
            const size_t ostride = p.poly ? 2 * (size_t)Cr : (size_t)Cr;           // elements between consecutive tile positions
            const size_t obase = p.poly ? ((size_t)bb * 2 * p.Tout + 2 * rt + ph) * Cr + cor + col4
                                        : ((size_t)bb * p.Tout + rt) * p.Cout + co0 + col4;
            const int tr = p.res_mode == RS_AVGPOOL2 ? 2 * rt : (p.res_mode == RS_NEAREST2 ? (rt >> 1) : rt);
            const size_t rbase = ((size_t)bb * p.res_Tin + tr) * p.Cout + co0 + col4;
            const int rmul = p.res_mode == RS_AVGPOOL2 ? 4 : (p.res_mode == RS_NEAREST2 ? 1 : 2);   // half-rows per position
#define OOFF(i) (obase + (size_t)(i) * ostride)
#define ROFF(i) (rbase + (size_t)(((i) * rmul) >> 1) * p.Cout)
            // Prefetched one chunk ahead: plain loads only.  (The pooled residual's second row is loaded where the value is consumed:
            // an average computed here, even predicated off, makes the warp wait for the loads it has just issued -- ncu showed
            // that scoreboard wait as 41 % of the epilogue loop's samples, profiles/r02_summary.md.)
            auto load_res = [&](int cb, float4 (&R)[4]) {
#pragma unroll
                for (int i = 0; i < 4; ++i) R[i] = ldg4(p.res + ROFF(i) + cb);
            };
            float4 Rn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) Rn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.res) {
                if (quad == 0) {   // pull this warp's residual rows into L2 while the mainloop of the tile is still running
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        for (int cb = 0; cb < HC; cb += 32) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + ROFF(i) + cb));
                            if (p.res_mode == RS_AVGPOOL2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + ROFF(i) + p.Cout + cb));
                        }
                }
                load_res(0, Rn);
            }
            // bias / time-embedding values of a chunk are loaded one chunk ahead as well (the first ones while the mainloop of the tile
            // still runs): a load consumed in the iteration that issues it costs a full L2 round trip per chunk
            const float* bias_p = p.bias ? p.bias + cor + col4 : nullptr;
            const float* temb_p = p.temb ? p.temb + (size_t)bb * p.temb_stride + cor + col4 : nullptr;
            float4 biasn = bias_p ? ldg4(bias_p) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 Tmn = temb_p ? ldg4(temb_p) : make_float4(0.f, 0.f, 0.f, 0.f);
            TL_WAIT(tl_a, mbar_wait(barAccFull + 8 * as, use & 1));
            tc_fence_after();
            const long long tl_e0 = tl ? clock64() : 0;
            const uint32_t acc_addr = tmem + ((uint32_t)(ew * 32) << 16) + as * ACC_COLS + eh * HC;
            uint32_t vn[16], c2n[16];
            tmem_ld16_async(acc_addr, vn);
            if (X3) tmem_ld16_async(acc_addr + (uint32_t)BN, c2n);
#pragma unroll 1
            for (int cb = 0; cb < HC; cb += 16) {
                float4 R[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) R[i] = Rn[i];
                if (p.res && p.res_mode == RS_AVGPOOL2) {   // AvgPool1d(2) of the residual (the two down-sampling ResBlocks): second row now
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 x1 = ldg4(p.res + ROFF(i) + p.Cout + cb);
                        R[i] = make_float4(0.5f * (R[i].x + x1.x), 0.5f * (R[i].y + x1.y), 0.5f * (R[i].z + x1.z), 0.5f * (R[i].w + x1.w));
                    }
                }
                // bias + time embedding once per chunk; the residual add only where there is one (three adds per value -> one or two)
                const float4 bt = make_float4(biasn.x + Tmn.x, biasn.y + Tmn.y, biasn.z + Tmn.z, biasn.w + Tmn.w);
                if (cb + 16 < HC) {
                    if (p.res) load_res(cb + 16, Rn);
                    if (bias_p) biasn = ldg4(bias_p + cb + 16);
                    if (temb_p) Tmn = ldg4(temb_p + cb + 16);
                }
                uint32_t v[16];
                tmem_ld_wait();
                if (X3) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(fmaf(__uint_as_float(c2n[i]), 1.0f / LO_SCALE, __uint_as_float(vn[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = vn[i];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(stg + lane * 16 + 4 * (j ^ ((lane >> 1) & 3))) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (cb + 16 < HC) {
                    tmem_ld16_async(acc_addr + (uint32_t)(cb + 16), vn);
                    if (X3) tmem_ld16_async(acc_addr + (uint32_t)(BN + cb + 16), c2n);
                } else {
                    tc_fence_before();
                    // every TMEM read of this thread has completed: the set may be overwritten (pair: the leader's MMA warp owns both
                    // CTAs' accumulators)
                    if (PAIR) mbar_arrive_cluster(barAccEmpty + 8 * as, 0);
                    else mbar_arrive(barAccEmpty + 8 * as);
                }
                __syncwarp();
                float4 O[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = 8 * i + seg;
                    const float4 a = *reinterpret_cast<const float4*>(stg + row * 16 + 4 * (quad ^ ((row >> 1) & 3)));
                    float4 o = make_float4(a.x + bt.x, a.y + bt.y, a.z + bt.z, a.w + bt.w);
                    if (p.res) o = make_float4(o.x + R[i].x, o.y + R[i].y, o.z + R[i].z, o.w + R[i].w);
                    if (rb >= 0) *reinterpret_cast<float4*>(p.out + OOFF(i) + cb) = o;
                    O[i] = o;
                }
                if (p.gn_partial) {
                    // GroupNorm statistics of the tensor being written: 4 positions x 4 channels of one segment per thread, shifted
                    // one-pass sums, then Chan's equal-count combination over the lanes (quads) that share a group
                    const float K = O[0].x;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float d0 = O[j].x - K, d1 = O[j].y - K, d2 = O[j].z - K, d3 = O[j].w - K;
                        s0 += d0; s1 += d1; s2 += d2; s3 += d3;
                        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
                    }
                    const float sd = (s0 + s1) + (s2 + s3);
                    float mean = fmaf(sd, 1.f / 16.f, K);
                    float m2 = fmaxf((q0 + q1) + (q2 + q3) - sd * sd * (1.f / 16.f), 0.f);
                    float n = 16.f;
                    for (int off = 1; off * 4 < cpg; off <<= 1) {
                        const float mo = __shfl_xor_sync(0xffffffffu, mean, off), qo = __shfl_xor_sync(0xffffffffu, m2, off);
                        const float d = mo - mean;
                        m2 = m2 + qo + d * d * (0.5f * n);
                        mean = 0.5f * (mean + mo);
                        n *= 2.f;
                    }
                    if ((col4 & (cpg - 1)) == 0) stats[warp * 256 + seg * gpt + ((cb + col4) >> cpg_sh)] = make_float2(mean, m2);
                }
                __syncwarp();   // the transpose buffer is reused by the next chunk
            }
#undef OOFF
#undef ROFF
            if (tl) tl_b += clock64() - tl_e0;
            if (p.gn_partial) {
                // the 4 warps of this warpgroup hold the 4 position-quarters of every segment: combine them into one
                // (count, mean, M2) record per (sample, 16-position segment, group) for gn_finalize
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eh) : "memory");
                const float nq = 4.f * cpg;
                const float2* sw = stats + eh * 1024;
                if (p.gn_tile) {
                    // Tout % 128 == 0: the tile's eight segments are one sample's -- ONE record per (tile, group): the 32 equal-count
                    // moments (8 segments x 4 position quarters) combined by Chan's formula
                    for (int gi = tid & 127; gi < gpt; gi += 128) {
                        float msum = 0.f;
                        for (int i = 0; i < 32; ++i) msum += sw[(i >> 3) * 256 + (i & 7) * gpt + gi].x;
                        const float mean = msum * (1.f / 32.f);
                        float m2 = 0.f, dd = 0.f;
                        for (int i = 0; i < 32; ++i) {
                            const float2 v = sw[(i >> 3) * 256 + (i & 7) * gpt + gi];
                            m2 += v.y; dd = fmaf(v.x - mean, v.x - mean, dd);
                        }
                        if (m_tile < n_mtiles) {
                            float* o = p.gn_partial + ((size_t)(p.poly ? 2 * m_tile + ph : m_tile) * (Cr / cpg) + cor / cpg + gi) * 3;
                            o[0] = 32.f * nq; o[1] = mean; o[2] = fmaf(nq, dd, m2);
                        }
                    }
                } else
                for (int e = tid & 127; e < 8 * gpt; e += 128) {
                    const int sg = e / gpt, gi = e - sg * gpt, g16 = m_tile * 8 + sg;
                    const float2 a = sw[e], b = sw[256 + e], c = sw[512 + e], d = sw[768 + e];
                    const float mean = 0.25f * ((a.x + b.x) + (c.x + d.x));
                    const float da = a.x - mean, db = b.x - mean, dc = c.x - mean, dd = d.x - mean;
                    const float m2 = (a.y + b.y) + (c.y + d.y) + nq * ((da * da + db * db) + (dc * dc + dd * dd));
                    if (g16 < p.nsegs16) {
                        // [b][segment][group][3]; polyphase: the 16 even (odd) rows of output segment pair 2 g16, 2 g16 + 1 are
                        // recorded as "segment" 2 g16 + phase (statistics do not care which 16 rows a record holds)
                        float* o = p.gn_partial + ((size_t)(p.poly ? 2 * g16 + ph : g16) * (Cr / cpg) + cor / cpg + gi) * 3;
                        o[0] = 4.f * nq; o[1] = mean; o[2] = m2;
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eh) : "memory");   // the stats buffer is rewritten by the next tile
            }
        }
        if (tl && tid == 0) { tlo[TC_TL_EPI_WAIT] = tl_a; tlo[TC_TL_EPI_BUSY] = tl_b; tlo[TC_TL_TILES] = lt; tlo[TC_TL_TOTAL] = clock64() - tl_start; }
    } else if (!EPI8 && warp < 4) {
        // ================================================================ epilogue
        // TMEM gives each thread one M row (32 columns per load).  The 32x32 block is transposed through a padded
        // per-warp staging buffer so that global loads (residual, time embedding) and stores are 128-byte row segments:
        // lane -> (row 4*i + lane/8, columns 4*(lane%8)..+3), i = 0..7.
        float* stg = reinterpret_cast<float*>(smem + STAGING_OFF) + warp * (32 * STG_LD);
        float2* stats = reinterpret_cast<float2*>(smem + STAGING_OFF + STAGING_BYTES);   // [warp][segment * gpt + group]
        const int cpg = p.gn_cpg, gpt = p.gn_partial ? BN / cpg : 0;                        // channels per group, groups per tile
        const int cpg_sh = 31 - __clz(max(cpg, 1));                                         // cpg is a power of two
        const int col4 = (lane & 7) * 4;
        int lt = 0;
        for (int w = cid; w < nwork; w += ncl, ++lt) {
            const int n_tile = w % n_ntiles, m_tile = (w / n_ntiles) * CL + crank;   // m_tile >= n_mtiles: idle slot of the last group
            const int as = lt % NSETS, use = lt / NSETS;
            const int co0 = n_tile * BN;
            // This thread's 8 rows: row i is M row (TMEM lane) warp*32 + 4*i + lane/8 = segment 4*(i&1) + lane/8, position
            // warp*4 + i/2 of that segment -- two segments (h = i & 1), four consecutive positions (j = i >> 1) each.
            int rb[2], rt[2];   // sample (< 0: segment past the batch) and first position per segment
            size_t obase[2], rbase[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int g = m_tile * 8 + 4 * h + (lane >> 3);
                rb[h] = g < p.nsegs16 ? g / spt : -1;
                rt[h] = (g % spt) * 16 + warp * 4;
                const int bb = max(rb[h], 0), t = rt[h];            // clamped: loads stay in bounds, stores are predicated
                obase[h] = ((size_t)bb * p.Tout + t) * p.Cout + co0 + col4;
                const int tr = p.res_mode == RS_AVGPOOL2 ? 2 * t : (p.res_mode == RS_NEAREST2 ? (t >> 1) : t);
                rbase[h] = ((size_t)bb * p.res_Tin + tr) * p.Cout + co0 + col4;
            }
            // residual row of position j: 2j (pooled pair), j/2 (nearest x2; rt is a multiple of 4) or j
            const int rmul = p.res_mode == RS_AVGPOOL2 ? 4 : (p.res_mode == RS_NEAREST2 ? 1 : 2);   // half-rows per position
#define OOFF(i) (obase[(i) & 1] + (size_t)((i) >> 1) * p.Cout)
#define ROFF(i) (rbase[(i) & 1] + (size_t)((((i) >> 1) * rmul) >> 1) * p.Cout)
            // residual rows of one 32-column chunk, all 8 loads in flight at once (prefetched one chunk ahead)
            // plain loads only: the pooled residual's second row is loaded where the value is consumed (see the EPI8 epilogue)
            auto load_res = [&](int cb, float4 (&R)[8]) {
#pragma unroll
                for (int i = 0; i < 8; ++i) R[i] = ldg4(p.res + ROFF(i) + cb);
            };
            float4 Rn[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) Rn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.res) {
                // the mainloop of this tile is still running: pull the tile's residual rows into L2 now so that the
                // epilogue's loads below are L2 hits (each lane owns one 128-byte line per row and 32-column chunk pair)
                if ((lane & 7) == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        for (int cb = 0; cb < BN; cb += 32) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + ROFF(i) + cb));
                            if (p.res_mode == RS_AVGPOOL2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + ROFF(i) + p.Cout + cb));
                        }
                }
                load_res(0, Rn);
            }
            TL_WAIT(tl_a, mbar_wait(barAccFull + 8 * as, use & 1));
            tc_fence_after();
            const long long tl_e0 = tl ? clock64() : 0;
            const uint32_t acc_addr = tmem + ((uint32_t)(warp * 32) << 16) + as * ACC_COLS;
            // TMEM reads run one 32-column chunk ahead of the chunk being stored: the load of chunk i+1 is issued as soon as
            // chunk i has been moved to the staging buffer, so its latency overlaps the transpose, the adds, the global stores
            // and the GroupNorm statistics of chunk i (one warp per scheduler: nothing else hides it)
            uint32_t vn[32], c2n[32];
            tmem_ld32_async(acc_addr, vn);
            if (X3) tmem_ld32_async(acc_addr + (uint32_t)BN, c2n);
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += 32) {
                float4 R[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) R[i] = Rn[i];
                if (p.res && p.res_mode == RS_AVGPOOL2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 x1 = ldg4(p.res + ROFF(i) + p.Cout + cb);
                        R[i] = make_float4(0.5f * (R[i].x + x1.x), 0.5f * (R[i].y + x1.y), 0.5f * (R[i].z + x1.z), 0.5f * (R[i].w + x1.w));
                    }
                }
                if (p.res && cb + 32 < BN) load_res(cb + 32, Rn);
                // bias / time-embedding rows of this chunk: issued before the TMEM loads so their latency is covered
                const int co = co0 + cb + col4;
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) bias4 = ldg4(p.bias + co);
                float4 Tm[2];   // rows i and i+2 lie in the same segment, hence the same sample: two distinct rows per thread
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    Tm[h] = p.temb ? ldg4(p.temb + (size_t)max(rb[h], 0) * p.temb_stride + co) : make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t v[32];
                tmem_ld_wait();
                if (X3) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaf(__uint_as_float(c2n[i]), 1.0f / LO_SCALE, __uint_as_float(vn[i])));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = vn[i];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(stg + lane * STG_LD + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (cb + 32 < BN) {
                    tmem_ld32_async(acc_addr + (uint32_t)(cb + 32), vn);
                    if (X3) tmem_ld32_async(acc_addr + (uint32_t)(BN + cb + 32), c2n);
                }
                __syncwarp();
                float4 O[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 a = *reinterpret_cast<const float4*>(stg + (4 * i + (lane >> 3)) * STG_LD + col4);
                    const float4 tb = Tm[i & 1];
                    const float4 o = make_float4(a.x + bias4.x + tb.x + R[i].x, a.y + bias4.y + tb.y + R[i].y,
                                                 a.z + bias4.z + tb.z + R[i].z, a.w + bias4.w + tb.w + R[i].w);
                    if (rb[i & 1] >= 0) *reinterpret_cast<float4*>(p.out + OOFF(i) + cb) = o;
                    O[i] = o;
                }
                if (p.gn_partial) {
                    // GroupNorm statistics of the tensor being written (the consumer's Normalize, unet.py:71-74): this thread
                    // holds 4 positions x 4 channels of segment lane/8 (even i) and of segment 4 + lane/8 (odd i).  One pass
                    // of sums shifted by the first value (no E[x^2]-E[x]^2 cancellation), four independent chains per
                    // segment, then Chan's equal-count combination over the lanes that share a group.
                    float mean[2], m2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float K = O[h].x;
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float d0 = O[2 * j + h].x - K, d1 = O[2 * j + h].y - K, d2 = O[2 * j + h].z - K, d3 = O[2 * j + h].w - K;
                            s0 += d0; s1 += d1; s2 += d2; s3 += d3;
                            q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
                        }
                        const float sd = (s0 + s1) + (s2 + s3);
                        mean[h] = fmaf(sd, 1.f / 16.f, K);
                        m2[h] = fmaxf((q0 + q1) + (q2 + q3) - sd * sd * (1.f / 16.f), 0.f);
                    }
                    float n = 16.f;
                    for (int off = 1; off * 4 < cpg; off <<= 1) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float mo = __shfl_xor_sync(0xffffffffu, mean[h], off), qo = __shfl_xor_sync(0xffffffffu, m2[h], off);
                            const float d = mo - mean[h];
                            m2[h] = m2[h] + qo + d * d * (0.5f * n);
                            mean[h] = 0.5f * (mean[h] + mo);
                        }
                        n *= 2.f;
                    }
                    if (((col4) & (cpg - 1)) == 0) {
                        const int e = (lane >> 3) * gpt + ((cb + col4) >> cpg_sh);
                        stats[warp * 256 + e] = make_float2(mean[0], m2[0]);
                        stats[warp * 256 + 4 * gpt + e] = make_float2(mean[1], m2[1]);
                    }
                }
                __syncwarp();   // staging buffer is reused by the next chunk
            }
#undef OOFF
#undef ROFF
            tc_fence_before();
            if (PAIR) mbar_arrive_cluster(barAccEmpty + 8 * as, 0);   // the leader's MMA warp owns both CTAs' accumulators
            else mbar_arrive(barAccEmpty + 8 * as);                   // this accumulator set may be overwritten
            if (tl) tl_b += clock64() - tl_e0;
            if (p.gn_partial) {
                // the 4 epilogue warps hold the 4 position-quarters of every segment: combine them and write one
                // (count, mean, M2) record per (sample, 16-position segment, group) for gn_finalize
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const float nq = 4.f * cpg;
                for (int e = tid; e < 8 * gpt; e += 128) {
                    const float2 a = stats[e], b = stats[256 + e], c = stats[512 + e], d = stats[768 + e];
                    const float mean = 0.25f * ((a.x + b.x) + (c.x + d.x));
                    const float da = a.x - mean, db = b.x - mean, dc = c.x - mean, dd = d.x - mean;
                    const float m2 = (a.y + b.y) + (c.y + d.y) + nq * ((da * da + db * db) + (dc * dc + dd * dd));
                    const int seg = e / gpt, gi = e - seg * gpt, g16 = m_tile * 8 + seg;
                    if (g16 < p.nsegs16) {
                        float* o = p.gn_partial + ((size_t)g16 * (p.Cout / cpg) + co0 / cpg + gi) * 3;   // [b][segment][group][3]
                        o[0] = 4.f * nq; o[1] = mean; o[2] = m2;
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");   // the stats buffer is rewritten by the next tile
            }
        }
        if (tl && tid == 0) { tlo[TC_TL_EPI_WAIT] = tl_a; tlo[TC_TL_EPI_BUSY] = tl_b; tlo[TC_TL_TILES] = lt; tlo[TC_TL_TOTAL] = clock64() - tl_start; }
    } else if (DIRECT && warp < W_LOAD) {
        // ================================================================ activation producers (192 threads)
        // act_split_kernel's mapping: item idx = pt + 192*j -> segment r = idx & 7, 8-channel chunk c = (idx >> 3) & 3, slot
        // q = idx >> 5.  192 = 6 * 32, so the three items of a thread share r and c (one sample, one channel chunk -> ONE
        // set of GroupNorm scale / shift values per k-step) and differ in the slot only: q = pt/32 + 6*j.
        // The loop is written for instruction count and latency: with one or two producer warps per scheduler every dependent
        // instruction costs its full latency (ncu: 725 instructions per k-step and thread, issue slots 43 % used, profiles/
        // r02_summary.md), so everything that does not change per k-step is hoisted to tile / segment changes -- row pointers and
        // validity per tile, source selection per (tile, source) -- the k-step itself advances three pointers, issues 6 + 4
        // vector loads and runs ~11 instructions per value (affine, SiLU on ex2.approx / rcp.approx without the denormal
        // fix-ups of __expf, fp16 hi/lo split, range check).
        const int pt = tid - 32 * NEPI;
        bool bad = false;
        const int r = pt & 7, c = (pt >> 3) & 3, q0 = pt >> 5;
        struct Pre { float4 x[3][2]; float4 a[2], s[2]; };
        // producer-side cursor over (work item, k-step): `nxt` is what the in-flight loads belong to
        struct Cur {
            const float* row[3];     // row pointers of the three items at channel chunk c of the CURRENT k-step's source
            const float* sc; const float* sh;   // scale / shift at this k-step's channels (null: no affine)
            uint32_t ok;             // bit j: item j lies inside the batch and the sample (else it is conv zero padding)
            int silu;
        };
        int w = cid, ks = 0;
        // tile constants (per work item): sample, first position, validity
        int tb = 0, tile_b = 0; bool segv = false;
        auto tile_setup = [&](int w_) {
            const int m_tile = min((w_ / n_ntiles) * CL + crank, n_mtiles - 1);
            const int g = m_tile * 8 + r;
            segv = g < p.nsegs16;
            tile_b = g / spt;
            tb = (g - tile_b * spt) * 16 - 1 + q0;      // position of item 0; item j: + 6 j
        };
        // pointers for k-step ks_ of the current tile (called when the tile, the segment or the concat source changes)
        auto src_setup = [&](int ks_, Cur& cu) {
            const bool first = ks_ < nks0;
            const TcSeg& sg = first ? p.seg[0] : p.seg[1];
            const int kl = first ? ks_ : ks_ - nks0;
            const int cc = kl * BK + c * 8;
            const float* src; int ch, Cs;
            if (cc < sg.C0) { src = sg.src0; ch = cc; Cs = sg.C0; } else { src = sg.src1; ch = cc - sg.C0; Cs = sg.C1; }
            cu.ok = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int t = tb + 6 * j;
                const bool v = segv && t >= 0 && t < p.Tout;
                const int tr = v ? (sg.resample == RS_NEAREST2 ? (t >> 1) : t) : 0;
                cu.row[j] = src + ((size_t)(v ? tile_b : 0) * sg.Tin + tr) * Cs + ch;   // invalid items: a harmless in-bounds address, never loaded
                cu.ok |= (v ? 1u : 0u) << j;
            }
            if (sg.scale && segv) {
                const size_t o = (size_t)tile_b * (sg.C0 + sg.C1) + cc;
                cu.sc = sg.scale + o; cu.sh = sg.shift + o;
            } else cu.sc = cu.sh = nullptr;
            cu.silu = sg.silu;
        };
        auto issue = [&](const Cur& cu, Pre& P) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if ((cu.ok >> j) & 1) ldg8(cu.row[j], P.x[j][0], P.x[j][1]);
                else P.x[j][0] = P.x[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (cu.sc) { ldg8(cu.sc, P.a[0], P.a[1]); ldg8(cu.sh, P.s[0], P.s[1]); }
            else {
                P.a[0] = P.a[1] = make_float4(1.f, 1.f, 1.f, 1.f);
                P.s[0] = P.s[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // k-step boundaries at which the pointers must be rebuilt instead of advanced by 32 channels: segment change, and the
        // switch from src0 to src1 inside a virtual concat (C0 is a multiple of 32, so a k-step never straddles the two)
        const int sw0 = p.seg[0].src1 ? p.seg[0].C0 / BK : -1;
        const int sw1 = p.nseg > 1 && p.seg[1].src1 ? nks0 + p.seg[1].C0 / BK : -1;
        Cur nxt;
        bool have = w < nwork;
        Pre N;
        if (have) { tile_setup(w); src_setup(0, nxt); issue(nxt, N); }
        int ia = 0;
        while (have) {
            const Pre C = N;
            const uint32_t ok = nxt.ok;
            const int silu = nxt.silu;
            // advance the cursor and put the next k-step's loads in flight while this one is transformed
            if (++ks == nks) { ks = 0; w += ncl; have = w < nwork; if (have) { tile_setup(w); src_setup(0, nxt); } }
            else if (ks == nks0 || ks == sw0 || ks == sw1) src_setup(ks, nxt);
            else {
#pragma unroll
                for (int j = 0; j < 3; ++j) nxt.row[j] += BK;
                if (nxt.sc) { nxt.sc += BK; nxt.sh += BK; }
            }
            if (have) issue(nxt, N);
            const float a[8] = {C.a[0].x, C.a[0].y, C.a[0].z, C.a[0].w, C.a[1].x, C.a[1].y, C.a[1].z, C.a[1].w};
            const float sh[8] = {C.s[0].x, C.s[0].y, C.s[0].z, C.s[0].w, C.s[1].x, C.s[1].y, C.s[1].z, C.s[1].w};
            const int sa = ia % NA;
            TL_WAIT(tl_a, mbar_wait(barAempty + 8 * sa, ((ia / NA) & 1) ^ 1));
            uint8_t* img = smem + sa * A_STAGE + c * A_LBO + q0 * A_SBO + r * 16;
            // (measured: one straight-line block of all 24 values with a select for the padding items is SLOWER than three
            // 8-value items behind their validity branch -- 13.1 vs 11.1 k cycles per level-0 tile, profiles/r02_summary.md)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float v[8] = {C.x[j][0].x, C.x[j][0].y, C.x[j][0].z, C.x[j][0].w, C.x[j][1].x, C.x[j][1].y, C.x[j][1].z, C.x[j][1].w};
                if ((ok >> j) & 1) {                     // padding / rows past the batch stay exactly zero
                    if (silu == 1) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = silu_ftz(fmaf(a[e], v[e], sh[e]));
                    } else if (silu == 2) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) { const float u = fmaf(a[e], v[e], sh[e]); v[e] = u > 0.f ? u : 0.2f * u; }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaf(a[e], v[e], sh[e]);   // no affine: a = 1, s = 0 (exact)
                    }
                }
                uint4 hi, lo;
                if (X3) {
                    bad |= out_of_f16_range(v);
                    split8_f16(v, hi, lo);
                    *reinterpret_cast<uint4*>(img + j * 6 * A_SBO + A_TILE) = lo;
                } else round8_bf16(v, hi);
                *reinterpret_cast<uint4*>(img + j * 6 * A_SBO) = hi;
            }
            fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core's async proxy
            mbar_arrive(barAfull + 8 * sa);            // (pair: the peer's stage is forwarded to the leader by its issuer warp)
            ++ia;
        }
        if (X3 && bad && p.range_flag) atomicOr(p.range_flag, 1);
        if (tl && pt == 0) { tlo[TC_TL_PROD_WAIT] = tl_a; tlo[TC_TL_PROD_BUSY] = clock64() - tl_start - tl_a; }
    } else if (warp == W_LOAD) {
        // ================================================================ loader (whole warp runs the loop, one elected lane issues)
        {
            const uint32_t a_bytes = X3 ? A_STAGE : A_TILE, b_bytes = X3 ? B_STAGE : B_HALF;
            int ia = 0, ib = 0;
            for (int w = cid; w < nwork; w += ncl) {
                const int n_tile = w % n_ntiles;
                const int m_tile = min((w / n_ntiles) * CL + crank, n_mtiles - 1);   // idle slot: reload a valid tile, never stored
                for (int ks = 0; ks < nks; ++ks, ++ia) {
                    const bool first = ks < nks0;
                    const TcSeg& sg = first ? p.seg[0] : p.seg[1];
                    const int kl = first ? ks : ks - nks0;
                    const int sa = ia % NA;
                    if (!DIRECT) {   // DIRECT: the producer warps fill the activation ring
                        mbar_wait(barAempty + 8 * sa, ((ia / NA) & 1) ^ 1);
                        if (elect_one()) {
                            if (p.debug & 1) { if (leader) mbar_arrive(barAfull + 8 * sa); }   // timing experiment: no operand traffic
                            else if (PAIR) {   // U image as rows of 512 B: one stage = 36 (hi + lo) or 18 rows
                                if (leader) mbar_arrive_expect_tx(barAfull + 8 * sa, 2 * a_bytes);
                                tma_load_2d_pair(sA + sa * A_STAGE, &p.tmap_u[first ? 0 : 1], 0, (int)(((size_t)m_tile * sg.nks + kl) * (A_STAGE / 512)),
                                                 barAfull + 8 * sa);
                            } else {
                                mbar_arrive_expect_tx(barAfull + 8 * sa, a_bytes);
                                bulk_copy_g2s(sA + sa * A_STAGE, sg.U + ((size_t)m_tile * sg.nks + kl) * A_STAGE, a_bytes, barAfull + 8 * sa);
                            }
                        }
                        __syncwarp();
                    }
                    // weight image [k-step][tap][hi|lo][Cout/8][4 kc][8][8]: the BN columns of this tile are one contiguous slice
                    const size_t whalf = (size_t)p.Cout * (BK * 2);
                    const uint8_t* wsrc = sg.w + (size_t)kl * sg.taps * 2 * whalf + (size_t)n_tile * BN * (BK * 2);
                    // polyphase up-conv: the even phase (first half of the columns) has no tap 2, the odd phase no tap 0
                    const int tap_lo = p.poly && 2 * n_tile * BN >= p.Cout ? 1 : 0, tap_hi = p.poly ? tap_lo + 2 : sg.taps;
                    for (int tap = tap_lo; tap < tap_hi; ++tap, ++ib) {
                        const int sb = ib % NB;
                        TL_WAIT(tl_a, mbar_wait(barBempty + 8 * sb, ((ib / NB) & 1) ^ 1));   // all consumers of this stage are done
                        if (elect_one()) {
                          if (p.debug & 1) { if (leader) mbar_arrive(barBfull + 8 * sb); }
                          else if (PAIR) {
                            // this CTA's half of the tile's columns: rows of 8 columns (512 B) of the weight image
                            // [k-step][tap][hi|lo][Cout/8], box = BNL/8 rows; the leader expects both CTAs' bytes
                            const uint32_t dst = sB + sb * B_STAGE, bar = barBfull + 8 * sb;
                            if (leader) mbar_arrive_expect_tx(bar, 2 * b_bytes);
                            const int row = ((kl * sg.taps + tap) * 2) * (p.Cout >> 3) + ((n_tile * BN + crank * BNL) >> 3);
                            tma_load_2d_pair(dst, &p.tmap_w[first ? 0 : 1], 0, row, bar);
                            if (X3) tma_load_2d_pair(dst + B_HALF, &p.tmap_w[first ? 0 : 1], 0, row + (p.Cout >> 3), bar);
                          } else {
                            mbar_arrive_expect_tx(barBfull + 8 * sb, b_bytes);
                            const uint8_t* hi = wsrc + (size_t)tap * 2 * whalf;
                            const uint32_t dst = sB + sb * B_STAGE, bar = barBfull + 8 * sb;
                            if (CL > 1) {   // fetch 1/CL of the stage, deliver it to every CTA of the cluster
                                const uint32_t part = B_HALF / CL, off = crank * part;
                                bulk_copy_g2s_multicast(dst + off, hi + off, part, bar, CMASK);
                                if (X3) bulk_copy_g2s_multicast(dst + B_HALF + off, hi + whalf + off, part, bar, CMASK);
                            } else {
                                bulk_copy_g2s(dst, hi, B_HALF, bar);
                                if (X3) bulk_copy_g2s(dst + B_HALF, hi + whalf, B_HALF, bar);
                            }
                        }
                        }
                        __syncwarp();
                    }
                }
            }
            if (tl && lane == 0) tlo[TC_TL_LOAD_WAIT_B] = tl_a;
        }
    } else if (PAIR && !leader) {
        // ================================================================ pair peer: the leader issues the MMAs of both CTAs
        // Fused producer: forward "this CTA's activation stage is written" to the leader.  A remote arrive stalls the arriving warp
        // for the cluster round trip (measured: +330 cycles per k-step when the producer warps did it themselves), this warp has
        // nothing else to do; the five-stage activation ring hides the extra hop.
        if (DIRECT) {
            int ia = 0;
            for (int w = cid; w < nwork; w += ncl)
                for (int ks = 0; ks < nks; ++ks, ++ia) {
                    const int sa = ia % NA;
                    mbar_wait(barAfull + 8 * sa, (ia / NA) & 1);
                    if (elect_one()) mbar_arrive_cluster(barApeer + 8 * sa, 0);
                    __syncwarp();
                }
        }
    } else {
        // ================================================================ MMA issuer (whole warp runs the loop, one elected lane issues)
        {
            int ia = 0, ib = 0, lt = 0;
            for (int w = cid; w < nwork; w += ncl, ++lt) {
                const int as = lt % NSETS, use = lt / NSETS;
                const uint32_t d0 = tmem + as * ACC_COLS, d1 = d0 + BN;
                const int poly_lo = p.poly && 2 * (w % n_ntiles) * BN >= p.Cout ? 1 : 0;   // polyphase up-conv: first tap of this N tile
                TL_WAIT(tl_a, mbar_wait(barAccEmpty + 8 * as, (use & 1) ^ 1));   // (pair: both CTAs' epilogues have drained this set)
                tc_fence_after();
                uint32_t accum = 0, accum2 = 0;
                const bool skip_mma = (p.debug & 2) != 0;   // timing experiment: operand traffic only
                auto mma = [skip_mma](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
                    if (skip_mma) return;
                    if (PAIR) umma2_bf16(d, da, db, IDESC, acc);
                    else umma_bf16(d, da, db, IDESC, acc);
                };
                for (int ks = 0; ks < nks; ++ks, ++ia) {
                    const int sa = ia % NA;
                    const int taps = ks < nks0 ? p.seg[0].taps : p.seg[1].taps;
                    const int tap_lo = p.poly ? poly_lo : 0, tap_hi = p.poly ? poly_lo + 2 : taps;
                    TL_WAIT(tl_b, mbar_wait(barAfull + 8 * sa, (ia / NA) & 1));
                    if (PAIR && DIRECT) TL_WAIT(tl_b, mbar_wait(barApeer + 8 * sa, (ia / NA) & 1));   // the peer's half of the M = 256 rows
                    tc_fence_after();
                    for (int tap = tap_lo; tap < tap_hi; ++tap, ++ib) {
                        const int sb = ib % NB;
                        TL_WAIT(tl_c, mbar_wait(barBfull + 8 * sb, (ib / NB) & 1));
                        tc_fence_after();
                        const int shift = taps == 3 ? tap : 1;   // slot of the first row: position - 1 + tap
                        const uint32_t a_hi = sA + sa * A_STAGE + shift * A_SBO, a_lo = a_hi + A_TILE;
                        const uint32_t b_hi = sB + sb * B_STAGE, b_lo = b_hi + B_HALF;
                        if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            const uint64_t dah = make_desc(a_hi + kk * 2 * A_LBO, A_LBO, A_SBO);
                            const uint64_t dbh = make_desc(b_hi + kk * 2 * B_LBO, B_LBO, B_SBO);
                            if (CAT && p.cat) {
                                // N = 128 tiles: the lo weight tile follows the hi tile in the stage with the same 512-byte group
                                // stride and accumulator 1 follows accumulator 0 in TMEM, so a_hi x [b_hi | b_lo] is ONE N = 256
                                // MMA into [d0 | d1]: a_hi is read from shared memory once instead of twice (the N = 128 shape is
                                // bound by shared-memory operand reads: 128 -> 107 B/clk)
                                const uint64_t dal = make_desc(a_lo + kk * 2 * A_LBO, A_LBO, A_SBO);
                                if (!skip_mma) {
                                    umma_bf16(d0, dah, dbh, IDESC_CAT, kk == 0 ? accum : 1u);
                                    umma_bf16(d1, dal, dbh, IDESC, 1u);
                                }
                                continue;
                            }
                            mma(d0, dah, dbh, kk == 0 ? accum : 1u);
                            if (X3) {
                                const uint64_t dal = make_desc(a_lo + kk * 2 * A_LBO, A_LBO, A_SBO);
                                const uint64_t dbl = make_desc(b_lo + kk * 2 * B_LBO, B_LBO, B_SBO);
                                mma(d1, dah, dbl, kk == 0 ? accum2 : 1u);
                                mma(d1, dal, dbh, 1u);
                            }
                        }
                        if (PAIR) {   // one commit per barrier, delivered to the same barrier of both CTAs
                            umma2_commit_multicast(barBempty + 8 * sb, CMASK);
                            if (tap == tap_hi - 1) umma2_commit_multicast(barAempty + 8 * sa, CMASK);
                            if (tap == tap_hi - 1 && ks == nks - 1) umma2_commit_multicast(barAccFull + 8 * as, CMASK);
                        } else {
                            if (CL > 1) umma_commit_multicast(barBempty + 8 * sb, CMASK);   // every CTA's loader writes into this stage
                            else umma_commit(barBempty + 8 * sb);                           // weight stage free once these MMAs retire
                            if (tap == tap_hi - 1) umma_commit(barAempty + 8 * sa);
                            if (tap == tap_hi - 1 && ks == nks - 1) umma_commit(barAccFull + 8 * as);
                        }
                        }
                        __syncwarp();
                        accum = 1; accum2 = 1;
                    }
                }
            }
            if (tl && lane == 0) {
                tlo[TC_TL_MMA_WAIT_ACC] = tl_a; tlo[TC_TL_MMA_WAIT_A] = tl_b; tlo[TC_TL_MMA_WAIT_B] = tl_c;
            }
        }
    }
#undef TL_WAIT
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into its shared memory
    if (warp == W_MMA) {
        if (PAIR) tmem_dealloc2(tmem, TMEM_COLS);
        else tmem_dealloc(tmem, TMEM_COLS);
    }
}

uint16_t bf16_rn(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);   // inf / nan
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
uint16_t f16_rn(float f) { return __half_as_ushort(__float2half_rn(f)); }
float f16_to_f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

}  // namespace

bool conv_tc_eligible(int Cin0, int Cin1, int Cout, int Tout, int taps, int stride) {
    return stride == 1 && (taps == 1 || taps == 3) && Cin0 > 0 && Cin0 % TC_BK == 0 && Cin1 % TC_BK == 0 && Cout % 128 == 0 &&
           Tout > 0 && Tout % 16 == 0;
}

// Output channels per tile.  N=128 reads 128 B/cycle of operands from shared memory: the MMAs alone (no copies) stop
// near 400 TFLOP/s fp32-equivalent; N=256 needs 96 B/cycle and reaches 590, but with the second f16x3 accumulator it
// fills TMEM, so its epilogue does not overlap the next mainloop.  Measured per layer shape (tools/conv_bench.py,
// profiles/r01_conv_bench.txt) N=256 wins wherever Cout % 256 == 0, including the K=512 1x1 convs (qkv 0.93 vs 1.32 ms).
// The CTA-pair shapes (cta_group::2) issue MMAs at the same rate and halve the weight staging, but the peer -> leader
// barrier relay doubles the load -> consume -> free round trip (2.7 vs 1.3 us) and the rings that fit in 227 KB no longer
// cover it: 25-45 % slower on every layer, so they are off by default.
int g_conv_tc_cluster = 2;          // CTAs per cluster sharing weight stages by multicast (1, 2 or 4); eegldm_set_conv_cluster
int g_conv_tc_pair = 1;             // cta_group::2 CTA pairs (M=256 per MMA): bit 0 = the 256-wide launches (default), bit 1 = the 128-wide ones
int g_conv_tc_bn256_stages = 1;     // minimum weight stages per tile for the N=256 shape
int g_conv_tc_cat = 1;              // N=128 f16x3 tiles: a_hi x [w_hi | w_lo] as one N=256 MMA
int g_conv_tc_epi8 = 1;             // two epilogue warpgroups (EPI8); eegldm_set_conv_tuning bit 6 switches it off (A/B timing)
bool conv_tc_gn_ok(int Cout, int G) {
    if (G <= 0 || Cout % G) return false;
    const int cpg = Cout / G;
    // groups per epilogue warpgroup (half a tile in the two-warpgroup form) <= 32; 32-channel groups take the one-warpgroup form
    const int bn = Cout % 256 == 0 ? 256 : 128;
    return (cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) && Cout % 128 == 0 && (g_conv_tc_epi8 && cpg != 32 ? bn / 2 : bn) / cpg <= 32;
}
int conv_tc_bn(int Cout, int weight_stages) { return (Cout % 256 == 0 && weight_stages >= g_conv_tc_bn256_stages) ? 256 : 128; }

// [Cout][Cin][k] fp32 -> [k-step][tap][hi|lo][Cout/8][kc 4][r 8][e 8] 16-bit: per (k-step, tap, half) the 8-column
// groups are consecutive 512-byte blocks of 4 K-major core matrices, so the columns of ANY tile width are one contiguous
// slice that is also the tile's shared-memory image (LBO 128, SBO 512):
//   element (co, ci) at  ((ks*k + tap)*2 + half) * Cout*32  +  (co/8)*256 + kc*64 + (co%8)*8 + e,   ci = ks*32 + kc*8 + e
// x3: hi = fp16(w), lo = fp16((w - hi) * 2^11);  otherwise hi = bf16(w), lo unused (left zero).
void pack_conv_tc(const float* w, int Cout, int Cin, int k, bool x3, std::vector<uint16_t>& out) {
    const int nks = Cin / TC_BK;
    const size_t half = (size_t)Cout * TC_BK;   // u16 elements per half
    out.assign((size_t)nks * k * 2 * half, 0);
    for (int ks = 0; ks < nks; ++ks)
        for (int tap = 0; tap < k; ++tap) {
            uint16_t* hi = out.data() + ((size_t)ks * k + tap) * 2 * half;
            uint16_t* lo = hi + half;
            for (int co = 0; co < Cout; ++co)
                for (int kc = 0; kc < TC_BK / 8; ++kc)
                    for (int e = 0; e < 8; ++e) {
                        const int ci = ks * TC_BK + kc * 8 + e;
                        const float v = w[((size_t)co * Cin + ci) * k + tap];
                        const size_t o = (size_t)(co >> 3) * 256 + kc * 64 + (co & 7) * 8 + e;
                        if (x3) {
                            const uint16_t h = f16_rn(v);
                            hi[o] = h;
                            lo[o] = f16_rn((v - f16_to_f(h)) * LO_SCALE);
                        } else hi[o] = bf16_rn(v);
                    }
        }
}

// Polyphase image of a 3-tap conv that follows a nearest x2 upsample (TcConvParams.poly): on the low-resolution input u,
//   out[2t]   = w0 u[t-1] + (w1 + w2) u[t]            (even phase: columns 0 .. Cout-1,      taps (w0, w1+w2, 0))
//   out[2t+1] = (w0 + w1) u[t] + w2 u[t+1]            (odd phase:  columns Cout .. 2Cout-1,  taps (0, w0+w1, w2))
// -- the same sums the reference forms on the upsampled signal (unet.py:263-268, 291), with the two coincident taps added in fp32
// before the 16-bit split.  The zero taps are kept in the image (uniform layout) and skipped by the kernel per N tile.
void pack_conv_tc_poly(const float* w, int Cout, int Cin, bool x3, std::vector<uint16_t>& out) {
    std::vector<float> wp((size_t)2 * Cout * Cin * 3);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci) {
            const float* s = w + ((size_t)co * Cin + ci) * 3;
            float* e = wp.data() + ((size_t)co * Cin + ci) * 3;
            float* o = wp.data() + ((size_t)(Cout + co) * Cin + ci) * 3;
            e[0] = s[0]; e[1] = s[1] + s[2]; e[2] = 0.f;
            o[0] = 0.f; o[1] = s[0] + s[1]; o[2] = s[2];
        }
    pack_conv_tc(wp.data(), 2 * Cout, Cin, 3, x3, out);
}

size_t act_split_bytes(int nsegs16, int Cin) { return (size_t)((nsegs16 + 7) / 8) * (Cin / TC_BK) * A_STAGE; }

cudaError_t launch_act_split(const ActSplitParams& p, bool x3, cudaStream_t st) {
    if (p.nsegs16 <= 0) return cudaSuccess;
    dim3 grid(p.nks, (p.nsegs16 + 7) / 8);
    if (x3) act_split_kernel<true><<<grid, SPLIT_THREADS, 0, st>>>(p);
    else act_split_kernel<false><<<grid, SPLIT_THREADS, 0, st>>>(p);
    g_launch_count += 1;
    return cudaGetLastError();
}

template <bool X3, int BN, int CL, bool PAIR, bool DIRECT = false, bool EPI8 = false>
cudaError_t launch_conv_tc_t(const TcConvParams& p, int num_sms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<X3, BN, CL, PAIR, DIRECT, EPI8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(true));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int n_mtiles = (p.nsegs16 + 7) / 8;
    const int nwork = (p.Cout / BN) * ((n_mtiles + CL - 1) / CL);
    int nclusters = num_sms / CL;
    if (nwork < nclusters) nclusters = nwork;
    if (p.debug & 64) nclusters = 1;   // timing experiment: one cluster alone on the device (no memory-system contention)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nclusters * CL);   // persistent: one CTA per SM
    cfg.blockDim = dim3(conv_tc_threads(DIRECT, EPI8));
    cfg.dynamicSmemBytes = smem_bytes(true);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv_tc_kernel<X3, BN, CL, PAIR, DIRECT, EPI8>, p);
}

// 2-D tensor map over an image of 512-byte rows (256 u16), box = `box_rows` whole rows (CTA-pair loads, see the kernel comment).
// cuTensorMapEncodeTiled is a pure host-side encoder; it is resolved through the runtime so the library does not link libcuda.
static cudaError_t encode_rows512(CUtensorMap* tm, const void* base, size_t rows, int box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
        if (e != cudaSuccess) return e;
        if (!f || qr != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        fn = (EncodeFn)f;
    }
    const cuuint64_t dims[2] = {256, (cuuint64_t)rows}, strides[1] = {512};
    const cuuint32_t box[2] = {256, (cuuint32_t)box_rows}, es[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_conv_tc(const TcConvParams& p_in, bool x3, cudaStream_t st) {
    if (p_in.nsegs16 <= 0) return cudaSuccess;
    TcConvParams p = p_in;
    p.cat = g_conv_tc_cat;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    if (p.bn != 128 && p.bn != 256) return cudaErrorInvalidValue;
    cudaError_t e;
    // two-warpgroup epilogue (EPI8): every launch except the ones whose epilogue emits 32-channel GroupNorm groups or bf16-mode
    // attention operand images (one-warpgroup epilogue only); cluster sizes 1 and 2
    if (p.qkv16 && (!x3 || p.gn_partial)) return cudaErrorInvalidValue;   // operand-image output: f16x3, two-warpgroup epilogue only
    // polyphase up-conv: one segment of 3 taps, whole N tiles per phase, no residual, two-warpgroup epilogue only
    if (p.gn_tile && (!p.gn_partial || p.Tout % 128 || p.gn_cpg == 32)) return cudaErrorInvalidValue;   // tile records: two-warpgroup epilogue
    if (p.poly && (p.nseg != 1 || p.seg[0].taps != 3 || p.bn != 256 || (p.Cout / 2) % 256 || p.res || p.qkv16 ||
                   (p.gn_partial && p.gn_cpg == 32)))
        return cudaErrorInvalidValue;
    const bool epi8_ok = !(p.gn_partial && p.gn_cpg == 32);
    // CTA pairs (cta_group::2): g_conv_tc_pair bit 0 = the N = 256 launches, bit 1 = the N = 128 launches
    const bool pair = epi8_ok && ((p.bn == 256 && (g_conv_tc_pair & 1)) || (p.bn == 128 && (g_conv_tc_pair & 2)));
    if (pair) {
        const int n_mtiles = (p.nsegs16 + 7) / 8;
        for (int s = 0; s < p.nseg; ++s) {
            const TcSeg& sg = p.seg[s];
            e = encode_rows512(&p.tmap_w[s], sg.w, (size_t)sg.nks * sg.taps * 2 * (p.Cout / 8), p.bn / 16);   // BN/2 columns = BN/16 rows
            if (e == cudaSuccess && !p.direct) e = encode_rows512(&p.tmap_u[s], sg.U, (size_t)n_mtiles * sg.nks * 36, x3 ? 36 : 18);
            if (e != cudaSuccess) return e;
        }
#define EEGLDM_TCP(X3, BN) (p.direct ? launch_conv_tc_t<X3, BN, 2, true, true, true>(p, num_sms, st) : launch_conv_tc_t<X3, BN, 2, true, false, true>(p, num_sms, st))
        if (p.bn == 256) e = x3 ? EEGLDM_TCP(true, 256) : EEGLDM_TCP(false, 256);
        else e = x3 ? EEGLDM_TCP(true, 128) : EEGLDM_TCP(false, 128);
#undef EEGLDM_TCP
        if (e != cudaSuccess) return e;
        g_launch_count += 1;
        return cudaGetLastError();
    }
    const bool epi8 = (g_conv_tc_epi8 || p.qkv16 || p.poly || p.gn_tile) && epi8_ok && (p.direct || g_conv_tc_cluster == 2 || g_conv_tc_cluster == 1 || p.qkv16 || p.poly || p.gn_tile);
#define EEGLDM_TC8(X3, BN)                                                                               \
    (p.direct ? (g_conv_tc_cluster == 1 ? launch_conv_tc_t<X3, BN, 1, false, true, true>(p, num_sms, st)   \
                                        : launch_conv_tc_t<X3, BN, 2, false, true, true>(p, num_sms, st))  \
     : g_conv_tc_cluster == 2 ? launch_conv_tc_t<X3, BN, 2, false, false, true>(p, num_sms, st)          \
                              : launch_conv_tc_t<X3, BN, 1, false, false, true>(p, num_sms, st))
    if (epi8) {
        if (p.bn == 256) e = x3 ? EEGLDM_TC8(true, 256) : EEGLDM_TC8(false, 256);
        else e = x3 ? EEGLDM_TC8(true, 128) : EEGLDM_TC8(false, 128);
        if (e != cudaSuccess) return e;
        g_launch_count += 1;
        return cudaGetLastError();
    }
#undef EEGLDM_TC8
#define EEGLDM_TC(X3, BN)                                                                         \
    (p.direct ? (g_conv_tc_cluster == 1 ? launch_conv_tc_t<X3, BN, 1, false, true>(p, num_sms, st)   \
                                        : launch_conv_tc_t<X3, BN, 2, false, true>(p, num_sms, st))  \
     : g_conv_tc_cluster == 4 ? launch_conv_tc_t<X3, BN, 4, false>(p, num_sms, st)                 \
     : g_conv_tc_cluster == 2 ? launch_conv_tc_t<X3, BN, 2, false>(p, num_sms, st) : launch_conv_tc_t<X3, BN, 1, false>(p, num_sms, st))
    if (p.bn == 256) e = x3 ? EEGLDM_TC(true, 256) : EEGLDM_TC(false, 256);
    else e = x3 ? EEGLDM_TC(true, 128) : EEGLDM_TC(false, 128);
#undef EEGLDM_TC
    if (e != cudaSuccess) return e;
    g_launch_count += 1;
    return cudaGetLastError();
}

}  // namespace eegldm
