// tools/gather_bench.cu -- micro-benchmark behind DESIGN.md's gather model: how many random 4-byte gathers per second
// does a B200 sustain from a table of W bytes (L1 / L2 / HBM resident), with the index stream read from HBM like
// the Aj stream of the pull kernel?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__global__ void fill_idx(uint32_t* idx, size_t n, uint32_t mask) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        idx[i] = (uint32_t) x & mask;
    }
}

template<bool GATHER>
__global__ void __launch_bounds__(256, 4) gather_kernel(const uint4* __restrict__ idx, const float* __restrict__ table, size_t n4, float* out) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x * 4) {
        uint4 j[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            size_t q = i + (size_t) c * gridDim.x * blockDim.x;
            j[c]     = q < n4 ? __ldcs(idx + q) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (GATHER) acc += __ldg(table + j[c].x) + __ldg(table + j[c].y) + __ldg(table + j[c].z) + __ldg(table + j[c].w);
            else acc += __uint_as_float(j[c].x ^ j[c].y ^ j[c].z ^ j[c].w);
        }
    }
    if (acc == 123.456f) *out = acc;
}

int main() {
    const size_t n = (size_t) 1 << 29;// 512 Mi indices = 2 GiB stream
    uint32_t*    idx;
    float *      table, *out;
    cudaMalloc(&idx, n * 4);
    cudaMalloc(&table, (size_t) 1 << 30);
    cudaMalloc(&out, 4);
    cudaMemset(table, 0, (size_t) 1 << 30);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    for (int lg = 14; lg <= 30; lg += 2) {// table bytes 16 KiB .. 1 GiB
        const uint32_t mask = (uint32_t) (((size_t) 1 << lg) / 4 - 1);
        fill_idx<<<grid, 256>>>(idx, n, mask);
        for (int g = 0; g < 2; ++g) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (g) gather_kernel<true><<<grid, 256>>>((const uint4*) idx, table, n / 4, out);
                else gather_kernel<false><<<grid, 256>>>((const uint4*) idx, table, n / 4, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("table %8.0f KiB  %s  %.3f ms  %.1f G idx/s  stream %.0f GB/s\n", (double) ((size_t) 1 << lg) / 1024, g ? "gather" : "stream", best,
                   n / best / 1e6, n * 4 / best / 1e6);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
