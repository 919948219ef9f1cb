// tools/dsmem_bench.cu -- micro-benchmark behind DESIGN.md's hub-cache model: random 4-byte gathers from a table held in
// (distributed) shared memory of a thread-block cluster of C CTAs (1 CTA per SM, 1024 threads), index stream from HBM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bench dsmem_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void fill_idx(uint32_t* idx, size_t n, uint32_t mod) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        idx[i] = (uint32_t) (x % mod);
    }
}

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(uint32_t saddr, uint32_t rank) {
    uint32_t ra; float v;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(saddr), "r"(rank));
    asm("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
    return v;
}

// frac_local: share (0..256) of gathers served by the table; the rest go to the global table (L2 path)
template<int C>
__global__ void __launch_bounds__(1024, 1) dsmem_gather(const uint4* __restrict__ idx, const float* __restrict__ gtab, size_t n4, uint32_t per_cta,
                                                        uint32_t hub_share, float* out) {
    extern __shared__ float tab[];
    for (uint32_t i = threadIdx.x; i < per_cta; i += blockDim.x) tab[i] = (float) i;
    if (C > 1) cluster_sync(); else __syncthreads();
    const uint32_t base = (uint32_t) __cvta_generic_to_shared(tab);
    float acc = 0.f;
    auto g = [&](uint32_t j) -> float {
        if ((j & 255u) < hub_share) {
            const uint32_t s = j >> 8;// slot
            if (C == 1) return tab[s % per_cta];
            return ld_dsmem(base + ((s / C) % per_cta) * 4u, s % C);
        }
        return __ldg(gtab + (j & 0xffffffu));// 64 MiB table in L2
    };
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x * 2) {
        uint4 j[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            size_t q = i + (size_t) c * gridDim.x * blockDim.x;
            j[c]     = q < n4 ? __ldcs(idx + q) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) acc += g(j[c].x) + g(j[c].y) + g(j[c].z) + g(j[c].w);
    }
    if (acc == 123.456f) *out = acc;
    if (C > 1) cluster_sync();
}

template<int C> float run(const uint4* idx, const float* gtab, size_t n4, uint32_t share, float* out, int* n_cta) {
    const size_t smem = 192 * 1024;
    cudaFuncSetAttribute(dsmem_gather<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    cfg.gridDim = dim3(C);
    cudaOccupancyMaxActiveClusters(&nclusters, dsmem_gather<C>, &cfg);
    cfg.gridDim = dim3(nclusters * C);
    *n_cta = nclusters * C;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, dsmem_gather<C>, idx, gtab, n4, (uint32_t) (smem / 4), share, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t n = (size_t) 1 << 28;
    uint32_t* idx; float *gtab, *out;
    cudaMalloc(&idx, n * 4); cudaMalloc(&gtab, (size_t) 64 << 20); cudaMalloc(&out, 4);
    cudaMemset(gtab, 0, (size_t) 64 << 20);
    fill_idx<<<592, 256>>>(idx, n, 0xffffffffu);
    cudaDeviceSynchronize();
    for (uint32_t share : {256u, 192u, 128u, 64u, 0u}) {
        int nc; float ms;
        ms = run<1>((const uint4*) idx, gtab, n / 4, share, out, &nc); printf("C=1 ctas=%3d hub share %3u/256: %.3f ms  %.1f G gathers/s\n", nc, share, ms, n / ms / 1e6);
        ms = run<2>((const uint4*) idx, gtab, n / 4, share, out, &nc); printf("C=2 ctas=%3d hub share %3u/256: %.3f ms  %.1f G gathers/s\n", nc, share, ms, n / ms / 1e6);
        ms = run<4>((const uint4*) idx, gtab, n / 4, share, out, &nc); printf("C=4 ctas=%3d hub share %3u/256: %.3f ms  %.1f G gathers/s\n", nc, share, ms, n / ms / 1e6);
        ms = run<8>((const uint4*) idx, gtab, n / 4, share, out, &nc); printf("C=8 ctas=%3d hub share %3u/256: %.3f ms  %.1f G gathers/s\n", nc, share, ms, n / ms / 1e6);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
