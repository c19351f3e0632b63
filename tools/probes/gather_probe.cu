// gather_probe.cu -- how many random 32-byte sectors per second one B200 delivers from an L2-resident (or DRAM-resident)
// array, by the path the request takes: (a) divergent LDG (one 32-B sector per lane, through the L1TEX tag stage),
// (b) one 32-B cp.async.bulk per lane into shared memory (through the TMA unit), completion on a per-warp mbarrier.
// Measurement helper, not on the product path.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// (a) DEPTH independent divergent loads in flight per lane
template <int DEPTH>
__global__ void __launch_bounds__(256) ldg_kernel(const uint4* __restrict__ a, uint32_t n_sectors, int iters, uint32_t* out) {
    uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int it = 0; it < iters; it++) {
        uint4 v[DEPTH][2];
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {
            seed = mix(seed + d + 1);
            const uint32_t s = (uint32_t)(((uint64_t)seed * n_sectors) >> 32);
            v[d][0] = __ldg(a + 2 * (size_t)s);
            v[d][1] = __ldg(a + 2 * (size_t)s + 1);
        }
#pragma unroll
        for (int d = 0; d < DEPTH; d++) acc += v[d][0].x ^ v[d][1].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (b) each lane copies its 32-B sector with cp.async.bulk; DEPTH stages per warp, one mbarrier per stage
template <int DEPTH>
__global__ void __launch_bounds__(256) bulk_kernel(const uint4* __restrict__ a, uint32_t n_sectors, int iters, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                       // [nwarps][DEPTH]
    uint4* buf = reinterpret_cast<uint4*>(smem + 1024);                       // [nwarps][DEPTH][32 lanes][2]
    uint64_t* mybar = bars + warp * DEPTH;
    uint4* mybuf = buf + (size_t)warp * DEPTH * 64;
    if (lane == 0)
        for (int d = 0; d < DEPTH; d++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mybar + d)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    (void)nwarps;
    uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    auto issue = [&](int d) {
        seed = mix(seed + d + 1);
        const uint32_t s = (uint32_t)(((uint64_t)seed * n_sectors) >> 32);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mybar + d)), "r"(32 * 32) : "memory");
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                     ::"r"(smem_u32(mybuf + d * 64 + lane * 2)), "l"(a + 2 * (size_t)s), "r"(smem_u32(mybar + d)) : "memory");
    };
    for (int d = 0; d < DEPTH; d++) issue(d);
    for (int it = 0; it < iters; it++) {
        const uint32_t parity = it & 1;
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(mybar + d)), "r"(parity) : "memory");
            const uint4 x = mybuf[d * 64 + lane * 2], y = mybuf[d * 64 + lane * 2 + 1];
            acc += x.x ^ y.w;
            __syncwarp();
            if (it + 1 < iters) issue(d);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atol(argv[1]) : 48;
    const size_t bytes = mb << 20;
    const uint32_t n_sectors = (uint32_t)(bytes / 32);
    uint4* a; uint32_t* out;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&out, 4)); CK(cudaMemset(a, 1, bytes));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto report = [&](const char* name, int depth, int bps, int iters, float ms) {
        const double sectors = (double)sms * bps * 256 * (double)iters * depth;
        printf("%-6s array %4zu MB depth %d blocks/SM %d : %7.3f ms  %7.1f G sectors/s  (%.2f sectors/clk/SM at %d MHz)\n", name, mb, depth, bps, ms,
               sectors / ms / 1e6, sectors / (ms * 1e-3) / sms / (pr.clockRate * 1e3), pr.clockRate / 1000);
    };
#define RUN_LDG(D, BPS) { const int iters = 2000 / D; for (int rep = 0; rep < 3; rep++) { CK(cudaEventRecord(e0)); ldg_kernel<D><<<sms * BPS, 256>>>(a, n_sectors, iters, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); } \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); report("ldg", D, BPS, iters, ms); }
#define RUN_BULK(D, BPS) { const int iters = 2000 / D; const size_t sh = 1024 + 8 * D * 32 * 32; CK(cudaFuncSetAttribute(bulk_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh)); \
        for (int rep = 0; rep < 3; rep++) { CK(cudaEventRecord(e0)); bulk_kernel<D><<<sms * BPS, 256, sh>>>(a, n_sectors, iters, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); } \
        float ms; CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms, e0, e1)); report("bulk", D, BPS, iters, ms); }
    RUN_LDG(1, 3) RUN_LDG(1, 8) RUN_LDG(2, 4) RUN_LDG(4, 4) RUN_LDG(8, 2)
    RUN_BULK(1, 3) RUN_BULK(1, 8) RUN_BULK(2, 4) RUN_BULK(4, 4) RUN_BULK(4, 8)
    return 0;
}
