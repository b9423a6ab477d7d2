// Probe: can a non-tensor cp.async.bulk (global -> own shared memory) signal complete_tx on an mbarrier that lives in the PEER
// CTA of the cluster?  (tools/probe: hardware experiments, not part of the library.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __cluster_dims__(2, 1, 1) probe(const uint32_t* src, uint32_t* out, int mode) {
    __shared__ __align__(128) uint32_t buf[1024];
    __shared__ __align__(8) uint64_t bar;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t b = smem_u32(&bar), d = smem_u32(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {
        if (rank == 0) {
            // leader: expects its own 4096 bytes + the peer's 4096 bytes on ITS barrier
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(8192u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(4096u), "r"(b) : "memory");
        } else {
            uint32_t rb;   // the leader's barrier, as a shared::cluster address
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(b), "r"(0u));
            if (mode == 1) rb = b & 0xFEFFFFFFu;   // CUTLASS's Sm100MmaPeerBitMask trick
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src + 1024), "r"(4096u), "r"(rb) : "memory");
            out[8] = b; out[9] = rb; out[10] = d;
        }
    }
    if (rank == 0 && threadIdx.x == 0) {
        uint32_t ok = 0; long long t0 = clock64();
        while (!ok && clock64() - t0 < 2000000) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(0u) : "memory");
        }
        out[0] = ok; out[1] = buf[5];
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank == 1 && threadIdx.x == 0) out[2] = buf[5];   // the peer's own data (valid once the leader saw the barrier complete)
}
int main() {
    uint32_t *src, *out, h[2048], ho[16];
    for (int i = 0; i < 2048; ++i) h[i] = 1000 + i;
    cudaMalloc(&src, sizeof(h)); cudaMalloc(&out, 64);
    cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(out, 0, 64);
        probe<<<2, 32>>>(src, out, mode);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(ho, out, 64, cudaMemcpyDeviceToHost);
        printf("mode %d: %s  barrier completed=%u leader buf[5]=%u (want 1005) peer buf[5]=%u (want 2029)  bar=%08x remote=%08x dst=%08x\n", mode,
               cudaGetErrorString(e), ho[0], ho[1], ho[2], ho[8], ho[9], ho[10]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
