// warp_prim_latency.cu — dependent-chain latency (cycles per operation) of the warp-level primitives the manager warp of
// q8_beam_kernel is built from, one warp on an otherwise idle SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o opensearch-jvector_b200/build/warp_prim_latency scripts/warp_prim_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAIN(NAME, ...)                                                    \
    {                                                                       \
        uint32_t x = seed;                                                  \
        const long long t0 = clock64();                                     \
        _Pragma("unroll 1") for (int it = 0; it < 256; it++) {              \
            _Pragma("unroll") for (int u = 0; u < 8; u++) { __VA_ARGS__; }        \
        }                                                                   \
        const long long t1 = clock64();                                     \
        sink += x;                                                          \
        if (lane == 0) printf("%-34s %6.1f cycles\n", NAME, (double)(t1 - t0) / 2048.0); \
    }

__global__ void k(uint32_t seed, uint32_t *out) {
    __shared__ uint64_t sm[256];
    const int lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) sm[i] = (uint64_t)((i * 7 + 1) & 255);
    __syncwarp();
    uint32_t sink = 0;
    CHAIN("iadd (dependent)", x = x + 0x9e3779b1u)
    CHAIN("imad (dependent)", x = x * 0x9e3779b1u + 12345u)
    CHAIN("popc", x = __popc(x) + seed)
    CHAIN("ballot + use", x = __ballot_sync(0xffffffffu, (x + lane) & 1u) + seed)
    CHAIN("shfl (32-bit)", x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31) + 1u)
    CHAIN("shfl x2 (64-bit value)", { uint32_t a = __shfl_sync(0xffffffffu, x, 5), b = __shfl_sync(0xffffffffu, x ^ seed, 5); x = a + b; })
    CHAIN("reduce_or (REDUX)", x = __reduce_or_sync(0xffffffffu, x + lane) + seed)
    CHAIN("reduce_max (REDUX)", x = __reduce_max_sync(0xffffffffu, x + lane) + seed)
    CHAIN("match_any (32 distinct)", x = __match_any_sync(0xffffffffu, x + lane) + seed)
    CHAIN("match_any (all equal)", x = __match_any_sync(0xffffffffu, x & 0u) + seed + x)
    CHAIN("lds.64 (dependent address)", x = (uint32_t)sm[x & 255])
    CHAIN("lds.64 + 64-bit compare + select", { uint64_t v = sm[x & 255]; x = (v | 1ull) > ((uint64_t)seed << 3) ? (x + 17u) : (x + 3u); })
    CHAIN("sts + syncwarp + lds", { sm[lane] = x; __syncwarp(); x = (uint32_t)sm[(lane + 1) & 31] + 1u; __syncwarp(); })
    CHAIN("atomicOr smem (spread)", x = atomicOr((uint32_t *)sm + ((x + lane * 17) & 511), 1u << (x & 31)) + seed + x)
    CHAIN("atomicAdd smem lane0 + shfl", { uint32_t s = 0; if (lane == 0) s = atomicAdd((uint32_t *)sm, 1u); x = __shfl_sync(0xffffffffu, s, 0) + x; })
    out[lane] = sink;
}

int main() {
    uint32_t *d;
    cudaMalloc(&d, 128);
    k<<<1, 32>>>(1u, d);
    cudaDeviceSynchronize();
    k<<<1, 32>>>(1u, d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
