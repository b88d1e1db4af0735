// Optimizer and parameter utilities: dense TF1-form Adam (HBM-bound, the floor of the train step),
// Xavier-uniform init, fp32 -> bf16 shadow cast, l2 / loss reductions.
//
// [TF1] tf.train.AdamOptimizer (models/DAEs.py:102, :198) == ApplyAdam functor:
//     m   += (g - m) * (1 - beta1)
//     v   += (g*g - v) * (1 - beta2)
//     var -= (m * alpha) / (sqrt(v) + eps),   alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
// dense on every element of every trainable variable, every step (rows with g == 0 still move).
// Every operation is an explicitly rounded fp32 op (no FMA contraction) so the update is bit-exact
// against the NumPy oracle given the same gradient.
#include <cuda_runtime.h>

#include "adam.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace dae {

// n4 float4 groups; row_len4 = row_len/4 (groups per row) when row_touched != nullptr
__global__ void __launch_bounds__(256)
k_adam_vec4(float4* __restrict__ w, float4* __restrict__ m, float4* __restrict__ v, const float4* __restrict__ g,
            uint2* __restrict__ wb, const unsigned char* __restrict__ touched, unsigned int n4,
            unsigned int row_len4, const AdamConst c) {
    const unsigned int stride = gridDim.x * blockDim.x;                 // n4 < 2^31: 32-bit index math
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (touched == nullptr || touched[i / row_len4] != 0) gv = __ldcs(g + i);
        float4 wv = __ldcs(w + i), mv = __ldcs(m + i), vv = __ldcs(v + i);
        adam_one(wv.x, mv.x, vv.x, gv.x, c);
        adam_one(wv.y, mv.y, vv.y, gv.y, c);
        adam_one(wv.z, mv.z, vv.z, gv.z, c);
        adam_one(wv.w, mv.w, vv.w, gv.w, c);
        __stcs(w + i, wv);
        __stcs(m + i, mv);
        __stcs(v + i, vv);
        if (wb != nullptr) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(wv.x, wv.y), hi = __floats2bfloat162_rn(wv.z, wv.w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned int*>(&lo);
            pk.y = *reinterpret_cast<unsigned int*>(&hi);
            wb[i] = pk;      // default policy: the shadow is re-read by the next step's decode
        }
    }
}

// Same update on a row-major matrix [rows, row_len] (the catalogue rows this rank owns), gradient = dense part +
// sparse rows, with the bf16 tensor-core operand copy of the rows refreshed in the same pass.
__global__ void __launch_bounds__(256)
k_adam_rows_vec4(float4* __restrict__ w, float4* __restrict__ m, float4* __restrict__ v, const float4* __restrict__ g,
                 const float4* __restrict__ g_sparse, __nv_bfloat16* __restrict__ shadow,
                 const unsigned char* __restrict__ touched, unsigned int n4,
                 unsigned int row_len4, int row_shift, const AdamConst c) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const unsigned int lrow = row_shift >= 0 ? (i >> row_shift) : i / row_len4;   // H/4 is a power of two for H = 64, 128, 256
        // gradient = dense part (dW_dec) + sparse rows (dW_enc, read only where the step touched the row)
        float4 gv = g != nullptr ? __ldcs(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (g_sparse != nullptr && touched[lrow] != 0) {
            const float4 gs = __ldcs(g_sparse + i);
            gv.x = __fadd_rn(gv.x, gs.x); gv.y = __fadd_rn(gv.y, gs.y); gv.z = __fadd_rn(gv.z, gs.z); gv.w = __fadd_rn(gv.w, gs.w);
        }
        float4 wv = __ldcs(w + i), mv = __ldcs(m + i), vv = __ldcs(v + i);
        adam_one(wv.x, mv.x, vv.x, gv.x, c);
        adam_one(wv.y, mv.y, vv.y, gv.y, c);
        adam_one(wv.z, mv.z, vv.z, gv.z, c);
        adam_one(wv.w, mv.w, vv.w, gv.w, c);
        __stcs(w + i, wv);
        __stcs(m + i, mv);
        __stcs(v + i, vv);
        if (shadow != nullptr) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(wv.x, wv.y), hi = __floats2bfloat162_rn(wv.z, wv.w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned int*>(&lo);
            pk.y = *reinterpret_cast<unsigned int*>(&hi);
            reinterpret_cast<uint2*>(shadow)[i] = pk;      // default policy: the operand copy is re-read by the next step's decode
        }
    }
}

void launch_adam_rows(const AdamArgs& a, const float* g_sparse, __nv_bfloat16* shadow, cudaStream_t st) {
    AdamConst c{a.alpha, a.one_minus_b1, a.one_minus_b2, a.eps, a.lambda};
    const long long n4 = a.n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = 148LL * 16;
    if (blocks > cap) blocks = cap;
    const unsigned int rl4 = (unsigned int)(a.row_len / 4);
    int shift = -1;
    if (rl4 != 0 && (rl4 & (rl4 - 1)) == 0) { shift = 0; while ((1u << shift) < rl4) ++shift; }
    k_adam_rows_vec4<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(a.w), reinterpret_cast<float4*>(a.m),
                                                  reinterpret_cast<float4*>(a.v), reinterpret_cast<const float4*>(a.g),
                                                  reinterpret_cast<const float4*>(g_sparse), shadow, a.row_touched, (unsigned int)n4, rl4, shift, c);
}

__global__ void k_adam_scalar(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                              const float* __restrict__ g, __nv_bfloat16* __restrict__ wb,
                              const unsigned char* __restrict__ touched, long long begin, long long n, int row_len,
                              const AdamConst c) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gv = 0.f;
        if (touched == nullptr || touched[i / row_len] != 0) gv = g[i];
        float wv = w[i], mv = m[i], vv = v[i];
        adam_one(wv, mv, vv, gv, c);
        w[i] = wv; m[i] = mv; v[i] = vv;
        if (wb != nullptr) wb[i] = __float2bfloat16(wv);
    }
}

void launch_adam(const AdamArgs& a, cudaStream_t st) {
    AdamConst c{a.alpha, a.one_minus_b1, a.one_minus_b2, a.eps, a.lambda};
    const bool vec_ok = (a.row_len % 4 == 0 || a.row_touched == nullptr) &&
                        ((reinterpret_cast<uintptr_t>(a.w) | reinterpret_cast<uintptr_t>(a.m) |
                          reinterpret_cast<uintptr_t>(a.v) | reinterpret_cast<uintptr_t>(a.g)) % 16 == 0);
    long long n4 = vec_ok ? a.n / 4 : 0;
    if (n4 > 0) {
        long long blocks = (n4 + 255) / 256;
        const long long cap = 148LL * 16;
        if (blocks > cap) blocks = cap;
        k_adam_vec4<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(a.w), reinterpret_cast<float4*>(a.m),
                                                 reinterpret_cast<float4*>(a.v), reinterpret_cast<const float4*>(a.g),
                                                 reinterpret_cast<uint2*>(a.w_bf16), a.row_touched,
                                                 (unsigned int)n4, (unsigned int)(a.row_touched ? a.row_len / 4 : 1), c);
    }
    const long long done = n4 * 4;
    if (done < a.n) {
        long long blocks = (a.n - done + 255) / 256;
        if (blocks > 1184) blocks = 1184;
        k_adam_scalar<<<(int)blocks, 256, 0, st>>>(a.w, a.m, a.v, a.g, a.w_bf16, a.row_touched, done, a.n,
                                                   a.row_len > 0 ? a.row_len : 1, c);
    }
}

// tf.contrib.layers.xavier_initializer() [TF1]: U(-l, l), l = sqrt(6 / (fan_in + fan_out))   (DAEs.py:54-55)
// Keyed by the GLOBAL element index, so the values do not depend on how the rows are spread over GPUs.
__device__ __forceinline__ float xavier_value(long long gi, float limit, unsigned long long seed, unsigned stream_id) {
    const float u = philox_uniform24(seed, stream_id, 0ull, static_cast<uint32_t>(gi >> 32), static_cast<uint32_t>(gi));
    return (2.f * u - 1.f) * limit;
}
__global__ void k_xavier_local(float* __restrict__ w, long long n_local, int N, int H, float limit,
                               unsigned long long seed, unsigned stream_id, int world, int rank) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += stride) {
        const int lrow = (int)(i / H), k = (int)(i - (long long)lrow * H);
        const int grow = item_global(lrow, world, rank);
        w[i] = grow < N ? xavier_value((long long)grow * H + k, limit, seed, stream_id) : 0.f;
    }
}
void launch_xavier_init(float* w, int n_local_rows, int N, int H, float limit, unsigned long long seed,
                        unsigned stream_id, int world, int rank, cudaStream_t st) {
    k_xavier_local<<<1184, 256, 0, st>>>(w, (long long)n_local_rows * H, N, H, limit, seed, stream_id, world, rank);
}

// sum of squares, one partial per block (tf.nn.l2_loss = sum(t^2)/2 [TF1], DAEs.py:79-82)
__global__ void k_sumsq(const float* __restrict__ x, long long n, float* __restrict__ partial) {
    __shared__ double s_red[8];
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = x[i];
        acc += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
        partial[blockIdx.x] = (float)t;
    }
}
void launch_sumsq(const float* x, long long n, float* partial, int nblocks, cudaStream_t st) {
    k_sumsq<<<nblocks, 256, 0, st>>>(x, n, partial);
}

// cost = inv_batch * sum(loss partials) + lambda * 0.5 * sum(sumsq partials)      (DAEs.py:100)
__global__ void k_reduce_loss(const float* __restrict__ partial, int n, const float* __restrict__ sq, int n_sq,
                              float lambda_half, float inv_batch, float* __restrict__ out) {
    __shared__ double s_red[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partial[i] * (double)inv_batch;
    for (int i = threadIdx.x; i < n_sq; i += blockDim.x) acc += (double)sq[i] * (double)lambda_half;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
        out[0] = (float)t;
    }
}
void launch_reduce_loss2(const float* partial, int n, const float* sumsq_partial, int n_sq, float lambda,
                         float inv_batch, float* loss_out, cudaStream_t st) {
    k_reduce_loss<<<1, 256, 0, st>>>(partial, n, sumsq_partial, n_sq, 0.5f * lambda, inv_batch, loss_out);
}

// zero every gradient row whose flag is set, and the flag (rows may have been touched by any rank).
// One warp scans 32 flags with a coalesced read, then zeroes each flagged row cooperatively.
__global__ void k_clear_flagged(int N, int H, float* __restrict__ g_enc, unsigned char* __restrict__ touched) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp_global * 32; base < N; base += nwarps * 32) {
        const int row = base + lane;
        const bool f = row < N && touched[row] != 0;
        unsigned int m = __ballot_sync(0xffffffffu, f);
        if (f) touched[row] = 0;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            float4* dst = reinterpret_cast<float4*>(g_enc + (size_t)(base + b) * H);
            for (int k = lane; k < (H >> 2); k += 32) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
void launch_clear_flagged(int N, int H, float* g_enc, unsigned char* touched, cudaStream_t st) {
    k_clear_flagged<<<(N + 255) / 256, 256, 0, st>>>(N, H, g_enc, touched);
}

// Force the module / functions to load now: with CUDA's lazy loading the FIRST launch of a kernel may
// synchronise the context, which would deadlock against a cross-GPU flag barrier already spinning.
void preload_optim() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_adam_vec4);
    cudaFuncGetAttributes(&a, k_adam_rows_vec4);
    cudaFuncGetAttributes(&a, k_adam_scalar);
    cudaFuncGetAttributes(&a, k_xavier_local);
    cudaFuncGetAttributes(&a, k_sumsq);
    cudaFuncGetAttributes(&a, k_reduce_loss);
    cudaFuncGetAttributes(&a, k_clear_flagged);
    (void)cudaGetLastError();
}

}  // namespace dae
