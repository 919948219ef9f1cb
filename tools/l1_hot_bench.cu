// tools/l1_hot_bench.cu -- can the L1 keep a small popularity-packed "hot" region of v resident while cold random gathers stream
// through it? A fraction p of the gathers goes to the first H bytes of a 64 MiB table, the rest anywhere in it. Variants:
//   0 plain ld.global.nc for both        1 hot L1::evict_last, cold L1::evict_first
//   2 hot L1::evict_last, cold L1::no_allocate      3 hot from shared memory (the table copy a hub class uses), cold plain
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_hot_bench l1_hot_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void fill_idx(uint32_t* idx, size_t n, uint32_t cold_mask, uint32_t hot_mask, uint32_t p1024) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        const uint32_t lo = (uint32_t) x, hi = (uint32_t) (x >> 32);
        const bool hot = (hi & 1023u) < p1024;
        idx[i] = hot ? ((lo & hot_mask) | 0x80000000u) : (lo & cold_mask);
    }
}

__device__ __forceinline__ float ld_last(const float* p) { float v; asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ float ld_first(const float* p) { float v; asm volatile("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ float ld_noalloc(const float* p) { float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }

template<int MODE>
__device__ __forceinline__ float gather(const float* __restrict__ table, const float* sh, uint32_t j) {
    const bool     hot = (j >> 31) != 0u;
    const uint32_t k   = j & 0x7fffffffu;
    if (MODE == 0) return __ldg(table + k);
    if (MODE == 1) return hot ? ld_last(table + k) : ld_first(table + k);
    if (MODE == 2) return hot ? ld_last(table + k) : ld_noalloc(table + k);
    return hot ? sh[k] : __ldg(table + k);
}

template<int MODE>
__global__ void __launch_bounds__(640, 1) gather_kernel(const uint4* __restrict__ idx, const float* __restrict__ table, size_t n4, float* out, uint32_t hot_n) {
    extern __shared__ float sh[];
    if (MODE == 3) {
        for (uint32_t i = threadIdx.x; i < hot_n; i += blockDim.x) sh[i] = table[i];
        __syncthreads();
    }
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x * 4) {
        uint4 j[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            size_t q = i + (size_t) c * gridDim.x * blockDim.x;
            j[c]     = q < n4 ? __ldcs(idx + q) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            acc += gather<MODE>(table, sh, j[c].x) + gather<MODE>(table, sh, j[c].y) + gather<MODE>(table, sh, j[c].z) + gather<MODE>(table, sh, j[c].w);
    }
    if (acc == 123.456f) *out = acc;
}

template<int MODE>
static float run(const uint32_t* idx, const float* table, size_t n, float* out, int grid, uint32_t hot_n) {
    const size_t smem = MODE == 3 ? (size_t) hot_n * 4 : 40960;// the tail class kernel carries 40 KB of warp slices
    cudaFuncSetAttribute(gather_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        gather_kernel<MODE><<<grid, 640, smem>>>((const uint4*) idx, table, n / 4, out, hot_n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best;
}

// second table: plain gathers only -- how much do the thread count and the L1 share left by the shared-memory carve-out matter?
template<int THREADS>
__global__ void __launch_bounds__(THREADS, 1) gather_plain(const uint4* __restrict__ idx, const float* __restrict__ table, size_t n4, float* out) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x * 4) {
        uint4 j[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            size_t q = i + (size_t) c * gridDim.x * blockDim.x;
            j[c]     = q < n4 ? __ldcs(idx + q) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) acc += __ldg(table + j[c].x) + __ldg(table + j[c].y) + __ldg(table + j[c].z) + __ldg(table + j[c].w);
    }
    if (acc == 123.456f) *out = acc;
}
template<int THREADS>
static void sweep_plain(const uint32_t* idx, const float* table, size_t n, float* out, int grid) {
    cudaFuncSetAttribute(gather_plain<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int smem_kib : {0, 16, 32, 40, 64, 100, 132}) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            gather_plain<THREADS><<<grid, THREADS, (size_t) smem_kib * 1024>>>((const uint4*) idx, table, n / 4, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        printf("plain gathers, %4d threads / SM, %3d KiB dynamic shared memory: %.3f ms  %.0f G/s\n", THREADS, smem_kib, best, n / best / 1e6);
    }
}

int main() {
    const size_t n = (size_t) 1 << 28;// 256 Mi gathers
    uint32_t*    idx;
    float *      table, *out;
    cudaMalloc(&idx, n * 4);
    cudaMalloc(&table, (size_t) 1 << 26);
    cudaMalloc(&out, 4);
    cudaMemset(table, 0, (size_t) 1 << 26);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount;
    printf("device %s, %d SMs, one CTA of 640 threads per SM\n", prop.name, grid);
    const uint32_t cold_mask = (1u << 24) - 1;// 64 MiB of floats
    for (uint32_t p1024 : {0u, 123u, 256u, 512u}) {
        for (uint32_t hot_kib : {32u, 64u, 128u, 160u}) {
            if (p1024 == 0 && hot_kib != 32u) continue;
            const uint32_t hot_n    = hot_kib * 256;
            uint32_t       hot_mask = 1;
            while (hot_mask * 2 <= hot_n) hot_mask *= 2;
            // hot indices uniform in [0, hot_n): mask to the next power of two and fold
            fill_idx<<<grid * 8, 256>>>(idx, n, cold_mask, hot_mask - 1, p1024);
            const uint32_t eff_n = hot_mask;// power-of-two part that is really addressed
            float ms0 = run<0>(idx, table, n, out, grid, eff_n), ms1 = run<1>(idx, table, n, out, grid, eff_n), ms2 = run<2>(idx, table, n, out, grid, eff_n),
                  ms3 = run<3>(idx, table, n, out, grid, eff_n);
            printf("hot share %.3f  hot region %4u KiB: plain %.3f ms %.0f G/s | evict_last/first %.3f ms %.0f G/s | evict_last/no_allocate %.3f ms %.0f G/s | smem hot %.3f ms %.0f G/s\n",
                   p1024 / 1024.0, eff_n / 256, ms0, n / ms0 / 1e6, ms1, n / ms1 / 1e6, ms2, n / ms2 / 1e6, ms3, n / ms3 / 1e6);
        }
    }
    fill_idx<<<grid * 8, 256>>>(idx, n, cold_mask, 0u, 0u);
    sweep_plain<512>(idx, table, n, out, grid);
    sweep_plain<640>(idx, table, n, out, grid);
    sweep_plain<768>(idx, table, n, out, grid);
    sweep_plain<1024>(idx, table, n, out, grid);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
