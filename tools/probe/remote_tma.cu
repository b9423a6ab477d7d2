// Probe: cp.async.bulk.tensor.2d with .cta_group::2 -- the peer CTA copies into its own shared memory and signals complete_tx on the
// LEADER's mbarrier (address with the CTA-rank bit cleared).  Also times the hop.  (tools/probe: hardware experiments.)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __cluster_dims__(2, 1, 1) probe(const __grid_constant__ CUtensorMap tm, uint32_t* out) {
    __shared__ __align__(128) uint16_t buf[16 * 256];   // 16 rows x 512 B
    __shared__ __align__(8) uint64_t bar;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t b = smem_u32(&bar), d = smem_u32(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        if (rank == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(16384u) : "memory");
        const uint32_t lb = b & 0xFEFFFFFFu;   // the leader's barrier from either CTA
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(d), "l"(&tm), "r"(0), "r"((int)(rank * 16)), "r"(lb) : "memory");
    }
    if (rank == 0 && threadIdx.x == 0) {
        uint32_t ok = 0;
        while (!ok && clock64() - t0 < 2000000) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(0u) : "memory");
        }
        out[0] = ok; out[1] = buf[5]; out[3] = (uint32_t)(clock64() - t0);
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank == 1 && threadIdx.x == 0) out[2] = buf[5];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int rows = 64;
    uint16_t* h = new uint16_t[rows * 256];
    for (int i = 0; i < rows * 256; ++i) h[i] = (uint16_t)(i & 0xFFFF);
    uint16_t* src; uint32_t *out, ho[4];
    cudaMalloc(&src, rows * 512); cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
    cudaMemcpy(src, h, rows * 512, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    if (e != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled: %s\n", cudaGetErrorString(e)); return 1; }
    CUtensorMap tm;
    cuuint64_t dims[2] = {256, (cuuint64_t)rows}, strides[1] = {512};
    cuuint32_t box[2] = {256, 16}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    probe<<<2, 32>>>(tm, out);
    e = cudaDeviceSynchronize();
    cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
    printf("%s  barrier completed=%u leader buf[5]=%u (want 5) peer buf[5]=%u (want %u) cycles=%u\n", cudaGetErrorString(e), ho[0], ho[1], ho[2],
           (16 * 256 + 5) & 0xFFFF, ho[3]);
    return 0;
}
