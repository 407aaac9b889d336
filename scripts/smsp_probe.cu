// smsp_probe.cu — which SM sub-partition (warp scheduler) does warp w of resident CTA slot s run on?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/smsp_probe scripts/smsp_probe.cu && gpurun_out/smsp_probe
//
// 148 x 4 CTAs of 5 warps (48 KB of shared memory each: 4 per SM, the shape of q8_beam_kernel).  On one SM a reference warp
// (slot 0, warp 0) and one other warp (slot s, warp w) run the same issue-bound FFMA loop at the same time: if they share a
// scheduler the loop takes twice as long.  Prints the slow-down of the reference warp for every (s, w).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(160, 4) probe(int *slot_ctr, int *arrive, long long *out, int ref_slot, int ref_warp, int oth_slot, int oth_warp,
                                                 int n_active, int target_sm) {
    extern __shared__ unsigned char smem[];
    __shared__ int s_slot;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((int)smid != target_sm) return;
    if (threadIdx.x == 0) s_slot = atomicAdd(slot_ctr, 1);
    __syncthreads();
    const int slot = s_slot, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_ref = slot == ref_slot && warp == ref_warp, is_oth = slot == oth_slot && warp == oth_warp && n_active == 2;
    if (!is_ref && !is_oth) return;
    if (lane == 0) {
        atomicAdd(arrive, 1);
        const long long t_in = clock64();
        while (atomicAdd(arrive, 0) < n_active && clock64() - t_in < 200000000ll) { // bounded: a missing partner must not hang the box
        }
    }
    __syncwarp();
    float a[8];
    for (int i = 0; i < 8; i++) a[i] = (float)(lane + i);
    const float b = 1.0000001f, c = 0.5f;
    const long long t0 = clock64();
    for (int it = 0; it < 8192; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 123.456f) smem[0] = 1;
    if (lane == 0) out[is_ref ? 0 : 1] = t1 - t0;
}

int main() {
    int *d_ctr;
    long long *d_out;
    cudaMalloc(&d_ctr, 8);
    cudaMalloc(&d_out, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    auto run = [&](int os, int ow, int n_active, int sm) -> long long {
        cudaMemset(d_ctr, 0, 8);
        cudaMemset(d_out, 0, 16);
        probe<<<148 * 4, 160, 48 * 1024>>>(d_ctr, d_ctr + 1, d_out, 0, 0, os, ow, n_active, sm);
        if (cudaDeviceSynchronize() != cudaSuccess) {
            printf("launch failed\n");
            exit(1);
        }
        long long h[2];
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        return h[0];
    };
    for (int sm : {0, 77}) {
        const long long alone = run(0, 0, 1, sm);
        printf("sm %d: reference warp alone: %lld cycles\n", sm, alone);
        for (int s = 0; s < 4; s++) {
            printf("  slot %d:", s);
            for (int w = 0; w < 5; w++) {
                if (s == 0 && w == 0) {
                    printf("   ref ");
                    continue;
                }
                printf("  %.2f ", (double)run(s, w, 2, sm) / (double)alone);
            }
            printf("\n");
        }
    }
    return 0;
}
