// What does an accumulator read cost, and does it slow the tensor core down?  (fused decode + top-K.)
// One CTA per SM shaped like k_itemtile: warp 1 issues tcgen05.mma (M = 128, N = 256, K = 16, bf16, operands = whatever
// is in shared memory) into accumulator 0, NW epilogue warps read accumulator 1 with tcgen05.ld 32x32b.x16 (or .x32).
// mode 1: loads only, 2: MMAs only, 3: both.  Prints cycles per 128 x 256 fp32 accumulator read (all warps together) and
// cycles per MMA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I spotify_recsys_challenge_2018_b200/csrc -o tools/probes/tmem_probe.bin tools/probes/tmem_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "umma.cuh"
using namespace dae;

template <int X>
__global__ void __launch_bounds__(576, 1) k_probe(int mode, int nw, int iters, long long* out, unsigned int* sink) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar_done, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    long long t0 = clock64();
    if (warp == 1) {
        if (mode & 2) {
            const uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
            constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo = ((smem_u32(smem + 65536) & 0x3FFFFu) >> 4) | (1u << 16);
            for (int it = 0; it < iters; ++it) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        umma_bf16(tmem_base, umma_desc_pack(a_lo + 2 * (k & 3) + 1024 * (k >> 2), kHi),
                                  umma_desc_pack(b_lo + 2 * (k & 3) + 2048 * (k >> 2), kHi), idesc, k != 0);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&bar_done);
            __syncwarp();
            mbar_wait(&bar_done, 0);
            long long t1 = clock64();
            if (lane == 0) out[blockIdx.x * 4 + 0] = t1 - t0;
        }
    } else if (warp >= 2 && warp < 2 + nw) {
        if (mode & 1) {
            const int q = warp & 3, part = (warp - 2) >> 2, nparts = (nw + 3) / 4;
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 256;
            unsigned int acc = 0;
            for (int it = 0; it < iters; ++it) {
                for (int c = part * X; c < 256; c += nparts * X) {
                    uint32_t r[X];
                    if (X == 16) tmem_ld16(t_addr + c, r); else tmem_ld32(t_addr + c, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < X; ++j) acc ^= r[j];
                }
            }
            long long t1 = clock64();
            if (lane == 0 && warp == 2) out[blockIdx.x * 4 + 1] = t1 - t0;
            if (acc == 0x12345678u) sink[0] = acc;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    long long* out; unsigned int* sink;
    cudaMallocManaged(&out, 148 * 4 * sizeof(long long)); cudaMalloc(&sink, 4);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k_probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_probe<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int x = 16; x <= 32; x *= 2)
        for (int nw = 4; nw <= 16; nw *= 2)
            for (int mode = 1; mode <= 3; ++mode) {
                if (mode == 2 && (nw != 4)) continue;
                for (int i = 0; i < 148 * 4; ++i) out[i] = 0;
                if (x == 16) k_probe<16><<<148, 576, smem>>>(mode, nw, iters, out, sink);
                else k_probe<32><<<148, 576, smem>>>(mode, nw, iters, out, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                printf("ld.x%d  warps %2d  mode %d:  cycles per accumulator read (128 x 256 fp32) %8.1f   cycles per MMA %7.1f\n", x, nw, mode,
                       (double)out[1] / iters, (double)out[0] / iters / 16);
            }
    return 0;
}
