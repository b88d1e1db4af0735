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
#include "umma.cuh"

namespace dae {

// geometry of the background streamer's work units (k_adam_bg below)
constexpr int kBgStageElems = 512;                   // floats per array per stage: 2 rows at H = 256
constexpr int kBgChunkStages = 8;                    // stages per claimed chunk (16 rows at H = 256)
constexpr unsigned int kBgStopBit = 0x40000000u;     // raised in ctl[0] by launch_bg_stop

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
// `bg_counter` (optional): rows >= n_rows - min(*bg_counter, bg_chunks) * bg_chunk_rows were claimed by the background
// streamer (k_adam_bg) earlier in the step, which has already updated their UNTOUCHED rows (g == 0): those are skipped
// here, so that every element still receives exactly one dense Adam update per step.
__global__ void __launch_bounds__(256)
k_adam_rows_vec4(float4* __restrict__ w, float4* __restrict__ m, float4* __restrict__ v, const float4* __restrict__ g,
                 const float4* __restrict__ g_sparse, __nv_bfloat16* __restrict__ shadow,
                 const unsigned char* __restrict__ touched, unsigned int n4,
                 unsigned int row_len4, int row_shift, const AdamConst c, const unsigned int* __restrict__ bg_counter,
                 unsigned int bg_n_chunks, unsigned int bg_rows_per_chunk, unsigned int n_rows, int touch_mode) {
    const unsigned int stride = gridDim.x * blockDim.x;
    unsigned int bg_row0 = 0xffffffffu;                 // first row the background streamer has taken care of
    if (bg_counter != nullptr) {
        const unsigned int done = min(__ldg(bg_counter) & (kBgStopBit - 1u), bg_n_chunks);
        bg_row0 = n_rows - done * bg_rows_per_chunk;
    }
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const unsigned int lrow = row_shift >= 0 ? (i >> row_shift) : i / row_len4;   // H/4 is a power of two for H = 64, 128, 256
        const bool hit = touched != nullptr && touched[lrow] != 0;
        if (lrow >= bg_row0 && !hit) continue;
        // touch_mode 1: only the rows no playlist lists (g == 0: independent of the step's backward), 2: only the listed rows
        if ((touch_mode == 1 && hit) || (touch_mode == 2 && !hit)) continue;
        // gradient = dense part (dW_dec) + sparse rows (dW_enc, read only where the step touched the row)
        float4 gv = g != nullptr ? __ldcs(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (g_sparse != nullptr && hit) {
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

int bg_chunk_rows(int H) { return kBgChunkStages * (kBgStageElems / H); }
int bg_chunks(int n_rows, int H) { return n_rows / bg_chunk_rows(H); }

void launch_adam_rows(const AdamArgs& a, const float* g_sparse, __nv_bfloat16* shadow, cudaStream_t st,
                      const BgAdam* bg) {
    AdamConst c{a.alpha, a.one_minus_b1, a.one_minus_b2, a.eps, a.lambda};
    const long long n4 = a.n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = 148LL * 16;
    if (blocks > cap) blocks = cap;
    const unsigned int rl4 = (unsigned int)(a.row_len / 4);
    int shift = -1;
    if (rl4 != 0 && (rl4 & (rl4 - 1)) == 0) { shift = 0; while ((1u << shift) < rl4) ++shift; }
    const unsigned int n_rows = (unsigned int)(a.n / a.row_len);
    k_adam_rows_vec4<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(a.w), reinterpret_cast<float4*>(a.m),
                                                  reinterpret_cast<float4*>(a.v), reinterpret_cast<const float4*>(a.g),
                                                  reinterpret_cast<const float4*>(g_sparse), shadow, a.row_touched, (unsigned int)n4, rl4, shift, c,
                                                  bg ? bg->ctl : nullptr, bg ? (unsigned int)bg_chunks(n_rows, a.row_len) : 0u,
                                                  (unsigned int)bg_chunk_rows(a.row_len), n_rows, a.touch_mode);
}

// ------------------------------------------------------------------------------------------
// Background encoder Adam (the dense TF1 Adam is the HBM floor of the step).  The encoder rows that no
// playlist of the batch touches (94 % of them) have g == 0: their update does not depend on this step's backward at
// all.  k_adam_bg streams them while the compute-bound front of the step (encode gather, G1 decode + loss, dh) leaves
// HBM half idle.  It is built to CO-RESIDE with those kernels on every SM: 160 threads, <= 40 registers and a 30 KB
// shared-memory ring (G1 keeps 194 KB, k_dh 193 KB of the SM's 227 KB), through which ONE thread moves rows with 1-D
// bulk copies (cp.async.bulk + mbarrier complete_tx in, bulk async-groups out) -- bytes in flight come from the copy
// engine, not from registers.  Work is claimed in chunks of kBgChunkStages stages from the TOP of the row range
// (the popularity-ranked tail of the catalogue: rarely touched) with an atomic counter, until the host-enqueued stop
// flag is raised (the decoder's fused dW + Adam kernel is about to need the SMs and the bandwidth).  The later dense
// pass (k_adam_rows_vec4 with bg_counter) does everything that is left: the unclaimed prefix and every touched row.
// Each element gets exactly one adam_one() per step, the same rounded operations on either path -> bit-identical
// to the single dense pass whatever the split point.
// ------------------------------------------------------------------------------------------
constexpr int kBgStages = 5;
constexpr int kBgStageBytes = 3 * kBgStageElems * 4; // w, m, v: 6 KB
// Register budget (measured with tools/probes/coresident_probe.cu): the 18 warps of the G1 CTA (96 registers) put 5 warps
// = 15 360 of an SM sub-partition's 16 384 registers on two of the four sub-partitions, so a co-resident warp there may
// own at most 1 024 registers = 32 per thread -- and at most one such warp per sub-partition.  Three warps of <= 32
// registers fit wherever the hardware puts them; the first version (5 warps x 40 registers) kept G1 off the SM until
// the streamer had exited.
constexpr int kBgThreads = 32 + 64;                  // I/O warp + 2 update warps (one row of 256 floats each)
constexpr int kBgSmem = kBgStages * kBgStageBytes + 256;
constexpr uint32_t kBgExit = 0x80000000u;

// ctl: [0] chunks claimed (atomic; kBgStopBit set: no further claims), [2 + smid] blocks resident on that SM (a second
// block on an SM exits: two rings would not leave room for the co-resident tensor-core CTA)
__global__ void __maxnreg__(32)
k_adam_bg(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, const unsigned char* __restrict__ touched,
          int n_rows, int H, int n_chunks, unsigned int* __restrict__ ctl, const AdamConst c) {
    extern __shared__ __align__(128) uint8_t bg_smem[];
    uint8_t* ring = bg_smem;
    uint64_t* ld_full = reinterpret_cast<uint64_t*>(bg_smem + kBgStages * kBgStageBytes);   // [kBgStages]
    uint64_t* done = ld_full + kBgStages;                                                      // [kBgStages]
    uint32_t* mask_s = reinterpret_cast<uint32_t*>(done + kBgStages);                          // [kBgStages]
    int* row0_s = reinterpret_cast<int*>(mask_s + kBgStages);                                  // [kBgStages]
    __shared__ int s_dup;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_dup = atomicAdd(ctl + 2 + (smid & 255u), 1u) != 0u;
        for (int s = 0; s < kBgStages; ++s) { mbar_init(&ld_full[s], 1); mbar_init(&done[s], 2); }
        fence_barrier_init();
    }
    __syncthreads();
    if (s_dup) return;
    const int rows_per_stage = kBgStageElems / H;
    const uint32_t row_bytes = static_cast<uint32_t>(H * 4);
    const int chunk_rows = kBgChunkStages * rows_per_stage;

    if (warp == 0) {
        if (lane != 0) return;
        // ================= I/O thread: claims chunks, fills the ring =================
        // A stage is refilled as soon as the update warps have READ it: they store their results straight to global
        // memory from registers, so a stage's turnaround is one HBM load latency plus the update -- with 30 KB of ring
        // that is what sets the streamer's bandwidth.  The claim of the NEXT chunk is issued one chunk ahead (the
        // atomic's round trip would otherwise stall this thread once per chunk).
        const uint64_t pol = policy_evict_first();
        unsigned int next_c = atomicAdd(ctl, 1u);
        int chunk_row0 = 0, sic = kBgChunkStages;       // stage-in-chunk: == kBgChunkStages -> move to the claimed chunk
        for (int i = 0;; ++i) {
            const int s = i % kBgStages;
            if (i >= kBgStages) mbar_wait(&done[s], static_cast<uint32_t>(((i / kBgStages) - 1) & 1));   // previous use read
            if (sic == kBgChunkStages) {
                if (next_c >= (unsigned int)n_chunks) {                  // stop bit raised, or nothing left: undo the claim,
                    atomicSub(ctl, 1u);                                   // tell the update warps
                    mask_s[s] = kBgExit;
                    mbar_arrive(&ld_full[s]);
                    break;
                }
                chunk_row0 = n_rows - ((int)next_c + 1) * chunk_rows;     // chunks are claimed from the top of the range
                next_c = atomicAdd(ctl, 1u);
                sic = 0;
            }
            const int row0 = chunk_row0 + sic * rows_per_stage;
            ++sic;
            uint32_t mask = 0;
            for (int j = 0; j < rows_per_stage; ++j) mask |= (touched[row0 + j] == 0 ? 1u : 0u) << j;
            mask_s[s] = mask; row0_s[s] = row0;
            uint8_t* dst = ring + s * kBgStageBytes;
            if (mask == (1u << rows_per_stage) - 1u) {                    // the common case: one contiguous copy per array
                const uint32_t bytes = row_bytes * rows_per_stage;
                const size_t off = (size_t)row0 * H;
                mbar_expect_tx(&ld_full[s], 3u * bytes);
                bulk_load_hint(dst, w + off, bytes, &ld_full[s], pol);
                bulk_load_hint(dst + kBgStageElems * 4, m + off, bytes, &ld_full[s], pol);
                bulk_load_hint(dst + 2 * kBgStageElems * 4, v + off, bytes, &ld_full[s], pol);
            } else if (mask != 0) {
                mbar_expect_tx(&ld_full[s], 3u * row_bytes * __popc(mask));
                for (int j = 0; j < rows_per_stage; ++j) {
                    if (!((mask >> j) & 1u)) continue;
                    const size_t off = (size_t)(row0 + j) * H;
                    bulk_load_hint(dst + j * row_bytes, w + off, row_bytes, &ld_full[s], pol);
                    bulk_load_hint(dst + kBgStageElems * 4 + j * row_bytes, m + off, row_bytes, &ld_full[s], pol);
                    bulk_load_hint(dst + 2 * kBgStageElems * 4 + j * row_bytes, v + off, row_bytes, &ld_full[s], pol);
                }
            } else {
                mbar_arrive(&ld_full[s]);                                  // every row of the stage is touched: nothing to do
            }
        }
    } else {
        // ================= update warps: thread t owns floats [4t, 4t+4) and [4t+256, 4t+260) of the stage's w, m, v ====
        const int t = threadIdx.x - 32;                  // 0..63
        for (int i = 0;; ++i) {
            const int s = i % kBgStages;
            mbar_wait(&ld_full[s], static_cast<uint32_t>((i / kBgStages) & 1));
            const uint32_t mask = mask_s[s];
            if (mask == kBgExit) break;
            const size_t base4 = ((size_t)row0_s[s] * H) / 4;                   // first float4 of the stage's rows
            const float4* ps = reinterpret_cast<const float4*>(ring + s * kBgStageBytes);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int e4 = t + 64 * half;                                     // float4 index inside the stage (0..127)
                const bool live = (mask >> ((4 * e4) / H)) & 1u;
                float4 wv, mv, vv;
                if (live) { wv = ps[e4]; mv = ps[e4 + kBgStageElems / 4]; vv = ps[e4 + 2 * (kBgStageElems / 4)]; }
                if (half == 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&done[s]);    // the stage has been read (mask / row0 included): it may be refilled
                }
                if (live) {
                    adam_one(wv.x, mv.x, vv.x, 0.f, c);
                    adam_one(wv.y, mv.y, vv.y, 0.f, c);
                    adam_one(wv.z, mv.z, vv.z, 0.f, c);
                    adam_one(wv.w, mv.w, vv.w, 0.f, c);
                    __stcs(reinterpret_cast<float4*>(w) + base4 + e4, wv);
                    __stcs(reinterpret_cast<float4*>(m) + base4 + e4, mv);
                    __stcs(reinterpret_cast<float4*>(v) + base4 + e4, vv);
                }
            }
        }
    }
}

void launch_adam_bg(const AdamArgs& a, const BgAdam& bg, cudaStream_t st) {
    AdamConst c{a.alpha, a.one_minus_b1, a.one_minus_b2, a.eps, a.lambda};
    const int n_rows = (int)(a.n / a.row_len);
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    k_adam_bg<<<sms, kBgThreads, kBgSmem, st>>>(a.w, a.m, a.v, a.row_touched, n_rows, a.row_len,
                                                  bg_chunks(n_rows, a.row_len), bg.ctl, c);
}

__global__ void k_or_u32(unsigned int* p, unsigned int v) { atomicOr(p, v); }
void launch_bg_stop(const BgAdam& bg, cudaStream_t st) { k_or_u32<<<1, 1, 0, st>>>(bg.ctl, kBgStopBit); }

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
__global__ void k_clear_flagged(int N, int H, float* __restrict__ g_enc, unsigned char* __restrict__ touched,
                                int* __restrict__ touch_cnt) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp_global * 32; base < N; base += nwarps * 32) {
        const int row = base + lane;
        const bool f = row < N && touched[row] != 0;
        unsigned int m = __ballot_sync(0xffffffffu, f);
        if (f) { touched[row] = 0; touch_cnt[row] = 0; }
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            float4* dst = reinterpret_cast<float4*>(g_enc + (size_t)(base + b) * H);
            for (int k = lane; k < (H >> 2); k += 32) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
void launch_clear_flagged(int N, int H, float* g_enc, unsigned char* touched, int* touch_cnt, cudaStream_t st) {
    k_clear_flagged<<<(N + 255) / 256, 256, 0, st>>>(N, H, g_enc, touched, touch_cnt);
}

// One warp per listed row: dense TF1 Adam with the row's sparse gradient (the pass over the few thousand rows a batch
// lists; every other row has already been updated with g == 0)
__global__ void __launch_bounds__(256)
k_adam_listed(float4* __restrict__ w, float4* __restrict__ m, float4* __restrict__ v, const float4* __restrict__ g,
              const int* __restrict__ list, int row_len4, const AdamConst c) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = list[0];
    for (int i = warp_global; i < n; i += nwarps) {
        const size_t base = (size_t)list[1 + i] * row_len4;
        for (int k = lane; k < row_len4; k += 32) {
            const float4 gv = __ldcs(g + base + k);
            float4 wv = __ldcs(w + base + k), mv = __ldcs(m + base + k), vv = __ldcs(v + base + k);
            adam_one(wv.x, mv.x, vv.x, gv.x, c);
            adam_one(wv.y, mv.y, vv.y, gv.y, c);
            adam_one(wv.z, mv.z, vv.z, gv.z, c);
            adam_one(wv.w, mv.w, vv.w, gv.w, c);
            __stcs(w + base + k, wv);
            __stcs(m + base + k, mv);
            __stcs(v + base + k, vv);
        }
    }
}
void launch_adam_listed(const AdamArgs& a, const float* g_sparse, const int* list, cudaStream_t st) {
    AdamConst c{a.alpha, a.one_minus_b1, a.one_minus_b2, a.eps, a.lambda};
    k_adam_listed<<<148 * 4, 256, 0, st>>>(reinterpret_cast<float4*>(a.w), reinterpret_cast<float4*>(a.m), reinterpret_cast<float4*>(a.v),
                                           reinterpret_cast<const float4*>(g_sparse), list, a.row_len / 4, c);
}
__global__ void k_clear_listed(int H4, float4* __restrict__ g_enc, unsigned char* __restrict__ touched, int* __restrict__ touch_cnt,
                               const int* __restrict__ list) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = list[0];
    for (int i = warp_global; i < n; i += nwarps) {
        const int row = list[1 + i];
        for (int k = lane; k < H4; k += 32) g_enc[(size_t)row * H4 + k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane == 0) { touched[row] = 0; touch_cnt[row] = 0; }
    }
}
void launch_clear_listed(int H, float* g_enc, unsigned char* touched, int* touch_cnt, const int* list, cudaStream_t st) {
    k_clear_listed<<<148 * 2, 256, 0, st>>>(H / 4, reinterpret_cast<float4*>(g_enc), touched, touch_cnt, list);
}

// Force the module / functions to load now: with CUDA's lazy loading the FIRST launch of a kernel may
// synchronise the context, which would deadlock against a cross-GPU flag barrier already spinning.
void preload_optim() {
    cudaFuncAttributes a;
    PRELOAD_KERNEL(k_adam_vec4);
    PRELOAD_KERNEL(k_adam_rows_vec4);
    PRELOAD_KERNEL(k_adam_scalar);
    PRELOAD_KERNEL(k_xavier_local);
    PRELOAD_KERNEL(k_sumsq);
    PRELOAD_KERNEL(k_reduce_loss);
    PRELOAD_KERNEL(k_clear_flagged);
    PRELOAD_KERNEL(k_adam_listed);
    PRELOAD_KERNEL(k_clear_listed);
    // co-resident with CTAs that own ~195 KB of shared memory: never ask for an L1-heavy carveout
    cudaFuncSetAttribute(k_adam_bg, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    PRELOAD_KERNEL(k_adam_bg);
    PRELOAD_KERNEL(k_or_u32);
    (void)cudaGetLastError();
}

void set_trap_log_optim(unsigned int* host_mapped) {
    cudaMemcpyToSymbol(g_trap_log, &host_mapped, sizeof(host_mapped));
}

}  // namespace dae
