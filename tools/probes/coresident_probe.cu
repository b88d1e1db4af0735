// Which resource keeps a big-shared-memory CTA off an SM that already runs a small persistent block?
// A: 148 blocks x TA threads, SA bytes of dynamic smem, spins until a flag (or 3 ms).  B: 148 blocks x TB threads, SB bytes,
// RB registers (forced by a dummy array), records when its first thread starts.  B starts "early" iff it co-resides with A.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); return v; }
__global__ void kA(volatile int* flag, unsigned long long* t) {
    extern __shared__ char sm[];
    if (threadIdx.x == 0) { sm[0] = 1; if (blockIdx.x == 0) t[0] = gt(); }
    unsigned long long t0 = gt();
    while (*flag == 0 && gt() - t0 < 3000000ull) { __nanosleep(200); }
    if (threadIdx.x == 0 && blockIdx.x == 0) t[1] = gt();
}
#ifndef RB
#define RB 64
#endif
template <int R>
__global__ void __launch_bounds__(576, 1) kB(unsigned long long* t, float* out) {
    extern __shared__ char sm[];
    if (threadIdx.x == 0) { sm[0] = 1; atomicMin(&t[2], gt()); }
    float acc[R];
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = threadIdx.x * i;
#pragma unroll 1
    for (int k = 0; k < 200; ++k)
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i] = acc[i] * 1.0001f + acc[(i + 1) % R];
    float s = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) atomicMax(&t[3], gt());
}
int main(int argc, char** argv) {
    int SA = argc > 1 ? atoi(argv[1]) : 31104, SB = argc > 2 ? atoi(argv[2]) : 197888, TA = argc > 3 ? atoi(argv[3]) : 160;
    int carve = argc > 4 ? atoi(argv[4]) : -1;
    unsigned long long* t; int* flag; float* out;
    cudaMallocManaged(&t, 64); cudaMalloc(&flag, 4); cudaMalloc(&out, 148 * 576 * 4);
    cudaStream_t s1, s2, s3; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking);
    cudaFuncSetAttribute(kB<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB);
    cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, SA);
    if (carve >= 0) cudaFuncSetAttribute(kA, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kB<RB>);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(flag, 0, 4); t[2] = ~0ull; t[3] = 0; cudaDeviceSynchronize();
        kA<<<148, TA, SA, s1>>>(flag, t);
        cudaEvent_t e; cudaEventCreate(&e);
        kB<RB><<<148, 576, SB, s2>>>(t, out);
        cudaEventRecord(e, s2);
        cudaStreamWaitEvent(s3, e, 0);
        cudaMemsetAsync(flag, 1, 4, s3);        // A is released only after B has finished: B early <=> co-resident
        cudaError_t err = cudaDeviceSynchronize();
        printf("SA=%d SB=%d TA=%d carve=%d regsB=%d : A %.1f..%.1f us, B start %.1f end %.1f us (%s) %s\n", SA, SB, TA, carve, fa.numRegs,
               0.0, (t[1] - t[0]) / 1e3, ((double)t[2] - (double)t[0]) / 1e3, ((double)t[3] - (double)t[0]) / 1e3,
               t[3] < t[1] ? "CO-RESIDENT" : "B waited for A", cudaGetErrorString(err));
    }
    return 0;
}
