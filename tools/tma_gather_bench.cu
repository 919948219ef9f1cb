// tools/tma_gather_bench.cu -- can the TMA unit serve random gathers BESIDE the LSU path? Random 4-byte gathers through L1
// are capped at one L1->L2 request per clock per SM (tools/gather_bench.cu). Here every gather is a 16-byte bulk copy
// (cp.async.bulk global -> shared, completion on an mbarrier) of the aligned 16 bytes around the wanted element;
// modes: all LSU, all TMA, half / half in the same batch.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather_bench tma_gather_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void fill_idx(uint32_t* idx, size_t n, uint32_t mask) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        idx[i] = (uint32_t) x & mask;
    }
}

constexpr int THREADS = 256;
constexpr int K       = 8;// gathers per thread per batch

// N_LSU of the K gathers of a thread go through ld.global.nc, the other K - N_LSU through 16-byte bulk copies
template<int N_LSU>
__global__ void __launch_bounds__(THREADS, 2) gather_kernel(const uint32_t* __restrict__ idx, const float* __restrict__ table, size_t n, float* out) {
    __shared__ __align__(16) float4 buf[THREADS * K];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_addr = (uint32_t) __cvta_generic_to_shared(&bar);
    constexpr int  N_TMA    = K - N_LSU;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    float    acc    = 0.f;
    uint32_t parity = 0;
    const size_t per_batch = (size_t) THREADS * K;
    for (size_t base = blockIdx.x * per_batch; base + per_batch <= n; base += (size_t) gridDim.x * per_batch) {
        uint32_t j[K];
#pragma unroll
        for (int k = 0; k < K; ++k) j[k] = __ldcs(idx + base + (size_t) k * THREADS + threadIdx.x);
        if (N_TMA > 0) {
            if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"((uint32_t) (THREADS * N_TMA * 16)));
#pragma unroll
            for (int k = N_LSU; k < K; ++k) {
                const uint32_t dst = (uint32_t) __cvta_generic_to_shared(&buf[threadIdx.x * K + k]);
                const float*   src = table + (j[k] & ~3u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(dst), "l"(src), "r"(bar_addr)
                             : "memory");
            }
        }
#pragma unroll
        for (int k = 0; k < N_LSU; ++k) acc += __ldg(table + j[k]);
        if (N_TMA > 0) {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done)
                             : "r"(bar_addr), "r"(parity)
                             : "memory");
            }
            parity ^= 1u;
#pragma unroll
            for (int k = N_LSU; k < K; ++k) acc += reinterpret_cast<const float*>(&buf[threadIdx.x * K + k])[j[k] & 3u];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        }
    }
    if (acc == 123.456f) *out = acc;
}

template<int N_LSU>
static void run(const char* name, const uint32_t* idx, const float* table, size_t n, float* out, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        gather_kernel<N_LSU><<<grid, THREADS>>>(idx, table, n, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    printf("  %-28s %.3f ms  %.1f G gathers/s  (%s)\n", name, best, n / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t n = (size_t) 1 << 28;
    uint32_t*    idx;
    float *      table, *out;
    cudaMalloc(&idx, n * 4);
    cudaMalloc(&table, (size_t) 1 << 28);
    cudaMalloc(&out, 4);
    cudaMemset(table, 0, (size_t) 1 << 28);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 2;
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    for (int lg = 22; lg <= 28; lg += 2) {// table bytes 4 MiB .. 256 MiB
        const uint32_t mask = (uint32_t) (((size_t) 1 << lg) / 4 - 1);
        fill_idx<<<grid, 256>>>(idx, n, mask);
        printf("table %.0f MiB\n", (double) ((size_t) 1 << lg) / 1048576);
        run<K>("all LSU", idx, table, n, out, grid);
        run<0>("all TMA (16 B bulk copies)", idx, table, n, out, grid);
        run<K / 2>("half LSU + half TMA", idx, table, n, out, grid);
        run<K - 2>("6 LSU + 2 TMA", idx, table, n, out, grid);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
