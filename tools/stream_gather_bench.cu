// tools/stream_gather_bench.cu -- does moving the Aj / Ax tile slices of the pull tail pass with 1-D bulk copies (cp.async.bulk, the
// TMA unit) instead of 128-bit LSU loads free the L1 -> L2 request port for the gathers?  The tail pass issues 512 gathers and streams
// 4 KB per tile; the port is ~89 % busy. Both variants run the same persistent grid (148 CTAs x WARPS warps, a warp owns
// one 512-entry tile at a time, slices requested one tile ahead, 16 gathers per lane from a 64 MB table).
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_gather_bench stream_gather_bench.cu
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

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D;\n bra W;\n D:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

// MODE 0: LSU loads of the slices; 1: bulk copies of the slices; 2: LSU slices, no gathers; 3: bulk slices, no gathers
template<int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) tile_kernel(const uint4* __restrict__ idx, const uint4* __restrict__ vals, const float* __restrict__ table,
                                                             uint32_t n_tiles, float* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr bool BULK   = (MODE & 1) != 0;
    constexpr bool GATHER = MODE < 2;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint4*         s_i  = reinterpret_cast<uint4*>(smem + warp * 4096);
    uint4*         s_v  = s_i + 128;
    uint64_t*      bar  = reinterpret_cast<uint64_t*>(smem + WARPS * 4096) + warp;
    const uint32_t n_warps = gridDim.x * WARPS, first = blockIdx.x * WARPS + warp;
    uint4          xi[4], xv[4];
    uint32_t       phase = 0;
    if (BULK) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0 && first < n_tiles) {
            mbar_expect(bar, 4096);
            bulk_g2s(s_i, idx + (size_t) first * 128, 2048, bar);
            bulk_g2s(s_v, vals + (size_t) first * 128, 2048, bar);
        }
    } else if (first < n_tiles) {
#pragma unroll
        for (int q = 0; q < 4; ++q) xi[q] = ld_stream_u4(idx + (size_t) first * 128 + q * 32 + lane), xv[q] = ld_stream_u4(vals + (size_t) first * 128 + q * 32 + lane);
    }
    float acc = 0.f;
    for (uint32_t t = first; t < n_tiles; t += n_warps) {
        uint4 ci[4], cv[4];
        if (BULK) {
            mbar_wait(bar, phase);
            phase ^= 1u;
#pragma unroll
            for (int q = 0; q < 4; ++q) ci[q] = s_i[q * 32 + lane], cv[q] = s_v[q * 32 + lane];
            __syncwarp();
            if (lane == 0 && t + n_warps < n_tiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect(bar, 4096);
                bulk_g2s(s_i, idx + (size_t) (t + n_warps) * 128, 2048, bar);
                bulk_g2s(s_v, vals + (size_t) (t + n_warps) * 128, 2048, bar);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) ci[q] = xi[q], cv[q] = xv[q];
            if (t + n_warps < n_tiles) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    xi[q] = ld_stream_u4(idx + (size_t) (t + n_warps) * 128 + q * 32 + lane), xv[q] = ld_stream_u4(vals + (size_t) (t + n_warps) * 128 + q * 32 + lane);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (GATHER) {
                acc += __uint_as_float(cv[q].x) * __ldg(table + ci[q].x) + __uint_as_float(cv[q].y) * __ldg(table + ci[q].y) +
                       __uint_as_float(cv[q].z) * __ldg(table + ci[q].z) + __uint_as_float(cv[q].w) * __ldg(table + ci[q].w);
            } else {
                acc += __uint_as_float(cv[q].x ^ ci[q].x) + __uint_as_float(cv[q].y ^ ci[q].y) + __uint_as_float(cv[q].z ^ ci[q].z) + __uint_as_float(cv[q].w ^ ci[q].w);
            }
        }
    }
    if (acc == 123.456f) *out = acc;
}

template<int MODE, int WARPS>
static float run(const uint4* idx, const uint4* vals, const float* table, uint32_t n_tiles, float* out, int sms) {
    auto         k    = tile_kernel<MODE, WARPS>;
    const size_t smem = WARPS * 4096 + WARPS * 8 + 128;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<<<sms, WARPS * 32, smem>>>(idx, vals, table, n_tiles, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t n = (size_t) 1 << 28;// 256 Mi entries: 1 GiB of indices + 1 GiB of values
    uint32_t *   idx, *vals;
    float *      table, *out;
    cudaMalloc(&idx, n * 4);
    cudaMalloc(&vals, n * 4);
    cudaMalloc(&table, (size_t) 1 << 26);
    cudaMalloc(&out, 4);
    cudaMemset(table, 0, (size_t) 1 << 26);
    cudaMemset(vals, 0, n * 4);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    fill_idx<<<sms * 4, 256>>>(idx, n, (uint32_t) (((size_t) 1 << 26) / 4 - 1));
    const uint32_t n_tiles = (uint32_t) (n / 512);
    printf("device %s, %d SMs; %u tiles of 512 entries, table 64 MiB\n", prop.name, sms, n_tiles);
    const char* names[4] = {"LSU slices + gathers ", "bulk slices + gathers", "LSU slices only      ", "bulk slices only     "};
    float       ms[4][2];
    ms[0][0] = run<0, 16>((const uint4*) idx, (const uint4*) vals, table, n_tiles, out, sms), ms[0][1] = run<0, 24>((const uint4*) idx, (const uint4*) vals, table, n_tiles, out, sms);
    ms[1][0] = run<1, 16>((const uint4*) idx, (const uint4*) vals, table, n_tiles, out, sms), ms[1][1] = run<1, 24>((const uint4*) idx, (const uint4*) vals, table, n_tiles, out, sms);
    ms[2][0] = run<2, 16>((const uint4*) idx, (const uint4*) vals, table, n_tiles, out, sms), ms[2][1] = run<2, 24>((const uint4*) idx, (const uint4*) vals, table, n_tiles, out, sms);
    ms[3][0] = run<3, 16>((const uint4*) idx, (const uint4*) vals, table, n_tiles, out, sms), ms[3][1] = run<3, 24>((const uint4*) idx, (const uint4*) vals, table, n_tiles, out, sms);
    for (int m = 0; m < 4; ++m)
        for (int w = 0; w < 2; ++w)
            printf("%s  %2d warps  %.3f ms  %.1f G entries/s  stream %.0f GB/s\n", names[m], w ? 24 : 16, ms[m][w], n / ms[m][w] / 1e6, n * 8 / ms[m][w] / 1e6);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
