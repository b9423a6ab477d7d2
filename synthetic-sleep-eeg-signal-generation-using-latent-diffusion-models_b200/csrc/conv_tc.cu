// tcgen05 implicit-GEMM convolution for sm_100a (the tensor-pipe path of the eegldm engine).
//
// Computes, for channels-last activations,
//     out[b,t,co] = bias[co] + temb[b,co] + res[b,t,co] + sum_seg sum_{k,ci} W_seg[co,ci,k] * u_seg[b, t+k-pad, ci]
//     u = resample(silu?(scale*x + shift))           (GroupNorm apply + SiLU + AvgPool/nearest)
// i.e. the same contract as conv_simt_kernel (reference: src/models/unet.py:263,291,302,158,161 + 308-327),
// in two launches:
//
//   act_split_kernel   one pass over x: apply the GroupNorm affine / SiLU / resample, split every fp32 value into
//                      16-bit parts and write them as ready-made shared-memory tile images ("U" tensors).
//   conv_tc_kernel     a pure async-copy + tcgen05 GEMM: M = 128 output positions, N = 128 output channels,
//                      K = taps*Cin per CTA; both operands arrive by cp.async.bulk (UBLKCP), D lives in TMEM.
//
//   * fp32 parity on a 16-bit tensor pipe ("f16x3"): every fp32 operand is split x = hi + lo/2048 with hi and lo
//     fp16 (11 + 11 significand bits, lo pre-scaled by 2^11 so it stays in fp16's normal range) and three products
//     are issued per K-slice: hi*hi into accumulator 0, hi*lo + lo*hi into accumulator 1; the epilogue returns
//     acc0 + acc1 * 2^-11.  Operand error ~2^-22; the tensor core's fp32 accumulate truncates (measured,
//     tools/conv_precision.py), which the separate correction accumulator keeps off the long chain.
//     EEGLDM_MATH_BF16_TC issues a single bf16 product (fast, NOT a parity mode).
//   * B operand = weights, pre-split and pre-packed on the host into the exact shared-memory image of one
//     pipeline stage (K-major, no swizzle, 8x16-byte core matrices): a stage is ONE bulk copy.
//   * A operand = activations in a "phase-strided halo" K-major layout: the 128 M rows of a tile are 8 segments of
//     16 consecutive positions; the 8 rows of one core matrix are the SAME offset in the 8 segments, and each
//     segment carries its own 2 halo positions (18 slots).  A tap shift of +-1 position is then a whole-core-matrix
//     shift = +-SBO bytes on the descriptor start address, so the 3 taps of a k=3 conv read ONE staged tile
//     (12.5 % halo overhead instead of 3 copies).  Segments may belong to different samples; halos at sample
//     edges are zero (the conv's padding).  U is stored [m_tile][k-step][hi|lo][kc][slot][segment][8 ch], i.e. one
//     contiguous 18 KB block per (tile, k-step): ONE bulk copy per stage.
//   * persistent, warp-specialised: one CTA per SM; warps 0-3 epilogue (TMEM -> registers -> global), warp 4 loader
//     (one thread), warp 5 MMA issuer (one thread); 4 x 18 KB activation + 8 x 16 KB weight stages (~205 KB smem);
//     two TMEM accumulator sets (512 columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
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
constexpr int BN = TC_BN;          // output channels per CTA
constexpr int BK = TC_BK;          // input channels per k-step
constexpr int SLOTS = 18;          // 16 positions + 2 halo slots per segment
constexpr int A_SBO = 128;         // bytes between core matrices along M (between slots)
constexpr int A_LBO = SLOTS * A_SBO;   // bytes between core matrices along K (8-channel chunks)
constexpr int A_TILE = (BK / 8) * A_LBO;   // one hi (or lo) activation tile: 9216 B
constexpr int A_STAGE = 2 * A_TILE;
static_assert(A_TILE == TC_U_HALF_BYTES, "engine and kernel disagree on the U tile size");
constexpr int B_SBO = 128;
constexpr int B_LBO = (BN / 8) * B_SBO;    // 2048
constexpr int B_HALF = (BK / 8) * B_LBO;   // 8192: hi (or lo) weight tile of one (tap, k-step)
constexpr int B_STAGE = 2 * B_HALF;
static_assert(B_HALF == TC_W_HALF_BYTES, "host packing and kernel disagree");
constexpr int NA = 4;              // activation ring (18 KB stages)
constexpr int NB = 8;              // weight ring (16 KB stages)
constexpr int N_ITEMS = 8 * SLOTS * (BK / 8);   // 576 16-byte items per activation tile
constexpr int SMEM_BYTES = NA * A_STAGE + NB * B_STAGE + 256;
constexpr int NUM_THREADS = 192;
constexpr int SPLIT_THREADS = 192;
static_assert(N_ITEMS % SPLIT_THREADS == 0, "items must divide evenly over the act_split block");

__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.f + __expf(-v)); }
__device__ __forceinline__ float act(float x, float a, float s, int silu) {
    const float v = fmaf(a, x, s);
    return silu ? silu_fast(v) : v;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

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
            split8_f16(v, hi, lo);
            *reinterpret_cast<uint4*>(img + A_TILE + off) = lo;
        } else round8_bf16(v, hi);
        *reinterpret_cast<uint4*>(img + off) = hi;
        if (img2) {
            if (X3) {
                split8_f16(vr, hi, lo);
                *reinterpret_cast<uint4*>(img2 + A_TILE + off) = lo;
            } else round8_bf16(vr, hi);
            *reinterpret_cast<uint4*>(img2 + off) = hi;
        }
    }
}

// ------------------------------------------------------------------------------------------------ conv_tc
// Persistent: one CTA per SM walks tiles (n_tile fastest, so concurrently running CTAs share activation tiles in L2);
// two TMEM accumulator sets, so the epilogue of tile i overlaps the mainloop of tile i+1.
template <bool X3>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const TcConvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sA = sbase, sB = sbase + NA * A_STAGE;
    const uint32_t bars = sB + NB * B_STAGE;           // 8-byte mbarriers
    const uint32_t barAfull = bars, barAempty = bars + 8 * NA, barBfull = bars + 16 * NA, barBempty = barBfull + 8 * NB,
                   barAccFull = barBempty + 8 * NB, barAccEmpty = barAccFull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + NA * A_STAGE + NB * B_STAGE + 16 * NA + 16 * NB + 32);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nks0 = p.seg[0].nks, nks = nks0 + (p.nseg > 1 ? p.seg[1].nks : 0);
    const int spt = p.Tout >> 4;   // 16-position segments per sample
    const int n_ntiles = p.Cout / BN;
    const int ntiles = n_ntiles * ((p.nsegs16 + 7) / 8);

    if (tid == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(barAfull + 8 * i, 1); mbar_init(barAempty + 8 * i, 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(barBfull + 8 * i, 1); mbar_init(barBempty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(barAccFull + 8 * i, 1); mbar_init(barAccEmpty + 8 * i, 128); }
        fence_mbar_init();
    }
    constexpr uint32_t ACC_COLS = X3 ? 2 * BN : BN;    // X3: accumulator 0 = hi*hi, accumulator 1 = cross terms * 2^11
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;       // two sets
    constexpr uint32_t IDESC = make_idesc(X3 ? 0u : 1u, BM, BN);
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
        // ================================================================ epilogue
        const int row = warp * 32 + lane;          // TMEM lane = M row
        const int r = row & 7, m = row >> 3;       // segment / offset inside the segment
        int lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int n_tile = tile % n_ntiles, m_tile = tile / n_ntiles;
            const int as = lt & 1;
            const int g = m_tile * 8 + r;
            const bool rowv = g < p.nsegs16;
            const int b = rowv ? g / spt : 0;
            const int t = (g % spt) * 16 + m;
            const int co0 = n_tile * BN;
            float* orow = p.out + ((size_t)b * p.Tout + t) * p.Cout + co0;
            const float* tb = p.temb ? p.temb + (size_t)b * p.temb_stride + co0 : nullptr;
            const float* bs = p.bias ? p.bias + co0 : nullptr;
            const float *res0 = nullptr, *res1 = nullptr;
            if (p.res) {
                const float* rb = p.res + (size_t)b * p.res_Tin * p.Cout + co0;
                if (p.res_mode == RS_AVGPOOL2) { res0 = rb + (size_t)(2 * t) * p.Cout; res1 = res0 + p.Cout; }
                else res0 = rb + (size_t)(p.res_mode == RS_NEAREST2 ? (t >> 1) : t) * p.Cout;
            }
            mbar_wait(barAccFull + 8 * as, (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t acc_addr = tmem + ((uint32_t)(warp * 32) << 16) + as * ACC_COLS;
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += 32) {
                uint32_t v[32];
                tmem_ld32(acc_addr + (uint32_t)cb, v);   // warp-collective: no divergence before this
                if (X3) {
                    uint32_t c2[32];
                    tmem_ld32(acc_addr + (uint32_t)(BN + cb), c2);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaf(__uint_as_float(c2[i]), 1.0f / LO_SCALE, __uint_as_float(v[i])));
                }
                if (!rowv) continue;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                           __uint_as_float(v[4 * q + 3]));
                    const int co = cb + 4 * q;
                    if (bs) { const float4 x = ldg4(bs + co); o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w; }
                    if (tb) { const float4 x = ldg4(tb + co); o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w; }
                    if (res0) {
                        float4 x = ldg4(res0 + co);
                        if (res1) { const float4 y = ldg4(res1 + co); x.x = 0.5f * (x.x + y.x); x.y = 0.5f * (x.y + y.y); x.z = 0.5f * (x.z + y.z); x.w = 0.5f * (x.w + y.w); }
                        o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
                    }
                    *reinterpret_cast<float4*>(orow + co) = o;
                }
            }
            tc_fence_before();
            mbar_arrive(barAccEmpty + 8 * as);   // this accumulator set may be overwritten
        }
    } else if (warp == 4) {
        // ================================================================ loader (one thread, bulk async copies)
        if (lane == 0) {
            const uint32_t a_bytes = X3 ? A_STAGE : A_TILE, b_bytes = X3 ? B_STAGE : B_HALF;
            int ia = 0, ib = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int n_tile = tile % n_ntiles, m_tile = tile / n_ntiles;
                for (int ks = 0; ks < nks; ++ks, ++ia) {
                    const bool first = ks < nks0;
                    const TcSeg& sg = first ? p.seg[0] : p.seg[1];
                    const int kl = first ? ks : ks - nks0;
                    const int sa = ia % NA;
                    mbar_wait(barAempty + 8 * sa, ((ia / NA) & 1) ^ 1);
                    mbar_arrive_expect_tx(barAfull + 8 * sa, a_bytes);
                    bulk_copy_g2s(sA + sa * A_STAGE, sg.U + ((size_t)m_tile * sg.nks + kl) * A_STAGE, a_bytes, barAfull + 8 * sa);
                    const uint8_t* wsrc = sg.w + ((size_t)n_tile * sg.nks + kl) * sg.taps * B_STAGE;
                    for (int tap = 0; tap < sg.taps; ++tap, ++ib) {
                        const int sb = ib % NB;
                        mbar_wait(barBempty + 8 * sb, ((ib / NB) & 1) ^ 1);
                        mbar_arrive_expect_tx(barBfull + 8 * sb, b_bytes);
                        bulk_copy_g2s(sB + sb * B_STAGE, wsrc + (size_t)tap * B_STAGE, b_bytes, barBfull + 8 * sb);
                    }
                }
            }
        }
    } else {
        // ================================================================ MMA issuer (one thread)
        if (lane == 0) {
            int ia = 0, ib = 0, lt = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
                const int as = lt & 1;
                const uint32_t d0 = tmem + as * ACC_COLS, d1 = d0 + BN;
                mbar_wait(barAccEmpty + 8 * as, ((lt >> 1) & 1) ^ 1);   // epilogue has drained this set
                tc_fence_after();
                uint32_t accum = 0, accum2 = 0;
                for (int ks = 0; ks < nks; ++ks, ++ia) {
                    const int sa = ia % NA;
                    const int taps = ks < nks0 ? p.seg[0].taps : p.seg[1].taps;
                    mbar_wait(barAfull + 8 * sa, (ia / NA) & 1);
                    tc_fence_after();
                    for (int tap = 0; tap < taps; ++tap, ++ib) {
                        const int sb = ib % NB;
                        mbar_wait(barBfull + 8 * sb, (ib / NB) & 1);
                        tc_fence_after();
                        const int shift = taps == 3 ? tap : 1;   // slot of the first row: position - 1 + tap
                        const uint32_t a_hi = sA + sa * A_STAGE + shift * A_SBO, a_lo = a_hi + A_TILE;
                        const uint32_t b_hi = sB + sb * B_STAGE, b_lo = b_hi + B_HALF;
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            const uint64_t dah = make_desc(a_hi + kk * 2 * A_LBO, A_LBO, A_SBO);
                            const uint64_t dbh = make_desc(b_hi + kk * 2 * B_LBO, B_LBO, B_SBO);
                            umma_bf16(d0, dah, dbh, IDESC, accum);
                            accum = 1;
                            if (X3) {
                                const uint64_t dal = make_desc(a_lo + kk * 2 * A_LBO, A_LBO, A_SBO);
                                const uint64_t dbl = make_desc(b_lo + kk * 2 * B_LBO, B_LBO, B_SBO);
                                umma_bf16(d1, dah, dbl, IDESC, accum2);
                                umma_bf16(d1, dal, dbh, IDESC, 1);
                                accum2 = 1;
                            }
                        }
                        umma_commit(barBempty + 8 * sb);   // weight stage free once these MMAs retire
                    }
                    umma_commit(barAempty + 8 * sa);
                }
                umma_commit(barAccFull + 8 * as);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, TMEM_COLS);
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
    return stride == 1 && (taps == 1 || taps == 3) && Cin0 > 0 && Cin0 % TC_BK == 0 && Cin1 % TC_BK == 0 && Cout % TC_BN == 0 &&
           Tout > 0 && Tout % 16 == 0;
}

// [Cout][Cin][k] fp32 -> per (n_tile, k-step, tap): [hi 8 KB | lo 8 KB], each the shared-memory image
//   byte(kc, ng, r, e) = kc*2048 + ng*128 + r*16 + e*2   for  co = n_tile*128 + ng*8 + r,  ci = ks*32 + kc*8 + e
// x3: hi = fp16(w), lo = fp16((w - hi) * 2^11);  otherwise hi = bf16(w), lo = 0 (never read).
void pack_conv_tc(const float* w, int Cout, int Cin, int k, bool x3, std::vector<uint16_t>& out) {
    const int nt = Cout / TC_BN, nks = Cin / TC_BK;
    out.assign((size_t)nt * nks * k * (2 * TC_W_HALF_BYTES / 2), 0);
    for (int n = 0; n < nt; ++n)
        for (int ks = 0; ks < nks; ++ks)
            for (int tap = 0; tap < k; ++tap) {
                uint16_t* hi = out.data() + (((size_t)n * nks + ks) * k + tap) * TC_W_HALF_BYTES;   // 2 halves of HALF/2 u16
                uint16_t* lo = hi + TC_W_HALF_BYTES / 2;
                for (int kc = 0; kc < TC_BK / 8; ++kc)
                    for (int ng = 0; ng < TC_BN / 8; ++ng)
                        for (int r = 0; r < 8; ++r)
                            for (int e = 0; e < 8; ++e) {
                                const int co = n * TC_BN + ng * 8 + r, ci = ks * TC_BK + kc * 8 + e;
                                const float v = w[((size_t)co * Cin + ci) * k + tap];
                                const size_t o = (size_t)kc * 1024 + ng * 64 + r * 8 + e;
                                if (x3) {
                                    const uint16_t h = f16_rn(v);
                                    hi[o] = h;
                                    lo[o] = f16_rn((v - f16_to_f(h)) * LO_SCALE);
                                } else hi[o] = bf16_rn(v);
                            }
            }
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

cudaError_t launch_conv_tc(const TcConvParams& p, bool x3, cudaStream_t st) {
    if (p.nsegs16 <= 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    const int ntiles = (p.Cout / BN) * ((p.nsegs16 + 7) / 8);
    dim3 grid(ntiles < num_sms ? ntiles : num_sms);   // persistent: one CTA per SM
    if (x3) conv_tc_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
    else conv_tc_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
    g_launch_count += 1;
    return cudaGetLastError();
}

}  // namespace eegldm
