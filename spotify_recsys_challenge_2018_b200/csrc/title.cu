// Title branch, CUDA-core part: the character CNN of models/title_models/Char_CNN.py:16-75
// (embedding -> parallel VALID convolutions over time -> bias -> ReLU -> max over time -> concat ->
// dropout) forward and backward, the mixing weights of DAE_title (models/DAEs.py:159-162), and the
// truncated-normal initialiser.  All of it is tiny next to the catalogue-sized output layer
// (~1.2 GFLOP per 256 titles): plain fp32 FMA, weights served from L2, no tensor cores.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "kernels.h"
#include "philox.cuh"

namespace dae {

// ------------------------------------------------------------------------------------------
// forward: one CTA per title, one thread per (width, filter) feature
// ------------------------------------------------------------------------------------------
// conv_W: widths back to back, each [w][E][F] (the reference's [fs, E, 1, F], Char_CNN.py:43); conv_b: [n_widths][F].
__global__ void __launch_bounds__(512)
k_charcnn_fwd(const long long* __restrict__ titles, const float* __restrict__ emb, const float* __restrict__ conv_W,
              const float* __restrict__ conv_b, const CnnShape s, float* __restrict__ feat,
              unsigned char* __restrict__ argpos, __nv_bfloat16* __restrict__ feat_d, __nv_bfloat16* __restrict__ feat_dT,
              int B, int bpad, float kp_t, unsigned long long seed, unsigned long long step, int row_offset) {
    extern __shared__ float s_x[];                 // [L][E] embedded title, pad rows zero
    const int b = blockIdx.x;
    const int j = threadIdx.x;                     // feature index in [0, D)
    const int D = s.F * s.n_widths;
    if (b >= B) {                                  // padding rows of the tensor-core operand
        for (int c = j; c < kTitleFpad; c += blockDim.x) {
            feat_d[(size_t)b * kTitleFpad + c] = __float2bfloat16(0.f);
            feat_dT[(size_t)c * bpad + b] = __float2bfloat16(0.f);
        }
        return;
    }
    for (int i = j; i < s.L * s.E; i += blockDim.x) {
        const int pos = i / s.E, e = i - pos * s.E;
        const long long id = titles[(size_t)b * s.L + pos];
        s_x[i] = (id >= 0 && id < s.C) ? emb[(size_t)id * s.E + e] : 0.f;      // pad id -1 -> zero vector
    }
    __syncthreads();
    float best = 0.f;
    int best_pos = 0;
    if (j < D) {
        const int wi = j / s.F, f = j - wi * s.F;
        const int w = s.width[wi];
        const float* W = conv_W + s.w_off[wi] + f;                             // [k][e][f]: threads read consecutive f
        const int P = s.L - w + 1;
        float acc[kTitleMaxLen];
#pragma unroll
        for (int p = 0; p < kTitleMaxLen; ++p) acc[p] = 0.f;
        for (int k = 0; k < w; ++k) {
            for (int e = 0; e < s.E; ++e) {
                const float wv = __ldg(W + (size_t)(k * s.E + e) * s.F);
#pragma unroll
                for (int p = 0; p < kTitleMaxLen; ++p)
                    if (p < P) acc[p] = fmaf(s_x[(p + k) * s.E + e], wv, acc[p]);
            }
        }
        const float bias = conv_b[wi * s.F + f];
        best = fmaxf(acc[0] + bias, 0.f);                                      // bias_add, relu, reduce_max (Char_CNN.py:50-58)
#pragma unroll
        for (int p = 1; p < kTitleMaxLen; ++p) {
            if (p < P) {
                const float v = fmaxf(acc[p] + bias, 0.f);
                if (v > best) { best = v; best_pos = p; }
            }
        }
        feat[(size_t)b * D + j] = best;
        argpos[(size_t)b * D + j] = static_cast<unsigned char>(best_pos);
    }
    // dropout (Char_CNN.py:67) and the bf16 operand copies [bpad, 512] / [512, bpad]; columns >= D stay zero
    for (int c = j; c < kTitleFpad; c += blockDim.x) {
        float v = 0.f;
        if (c < D && c == j) {
            const bool keep = philox_keep(seed, kStreamTitle, step, static_cast<uint32_t>(b + row_offset),
                                          static_cast<uint32_t>(c), kp_t);
            v = keep ? __fdiv_rn(best, kp_t) : 0.f;
        }
        const __nv_bfloat16 hb = __float2bfloat16(v);
        feat_d[(size_t)b * kTitleFpad + c] = hb;
        feat_dT[(size_t)c * bpad + b] = hb;
    }
}

// The same forward, register-tiled (one launch per filter width): CTA = 4 titles, thread = (filter f, position half).
// A weight is loaded ONCE per thread (one step ahead of its use) and feeds 4 titles x NP positions; the embedded titles sit
// in shared memory padded to float4 rows, so one 16-byte broadcast load feeds four FMAs.  NP = ceil(P / 2) is a template
// parameter and the second half starts at P - NP (one position may be computed by both halves): no predicated-off
// FMAs.  Same (k, e) summation order as the kernel above: bit-identical features.  92 M -> ~30 M warp instructions per
// 256 titles of the shipped shape (the kernel is issue-bound).
constexpr int kCnnG = 4;           // titles per CTA
template <int NP>
__device__ __forceinline__ void
charcnn_fwd_body(const long long* __restrict__ titles, const float* __restrict__ emb, const float* __restrict__ conv_W,
                 const float* __restrict__ conv_b, const CnnShape& s, int wi, float* __restrict__ feat,
                 unsigned char* __restrict__ argpos, __nv_bfloat16* __restrict__ feat_d, __nv_bfloat16* __restrict__ feat_dT,
                 int B, int bpad, float kp_t, unsigned long long seed, unsigned long long step, int row_offset, float* s_x) {
    const int Epad = (s.E + 3) & ~3;
    const int b0 = blockIdx.x * kCnnG;
    const int w = s.width[wi], P = s.L - w + 1, D = s.F * s.n_widths;
    const int half = threadIdx.x >= s.F ? 1 : 0, f = threadIdx.x - half * s.F;     // threads [0, F) / [F, 2F)
    const bool active = threadIdx.x < 2 * s.F;
    float* s_best = s_x + kCnnG * s.L * Epad;                 // [kCnnG][F] best of the second half
    int* s_bpos = reinterpret_cast<int*>(s_best + kCnnG * s.F);
    for (int i = threadIdx.x; i < kCnnG * s.L * Epad; i += blockDim.x) {
        const int g = i / (s.L * Epad), r = i - g * s.L * Epad;
        const int pos = r / Epad, e = r - pos * Epad;
        float v = 0.f;
        if (b0 + g < B && e < s.E) {
            const long long id = titles[(size_t)(b0 + g) * s.L + pos];
            if (id >= 0 && id < s.C) v = emb[(size_t)id * s.E + e];              // pad id -1 -> zero vector
        }
        s_x[i] = v;
    }
    __syncthreads();
    const int p0 = half ? P - NP : 0;                         // 2 * NP >= P: the halves may share one position
    float acc[kCnnG][NP];
#pragma unroll
    for (int g = 0; g < kCnnG; ++g)
#pragma unroll
        for (int pi = 0; pi < NP; ++pi) acc[g][pi] = 0.f;
    if (active) {
        const float* W = conv_W + s.w_off[wi] + f;            // [k][e][f]: threads read consecutive f
        // the weights of step (k, e4) are loaded one step ahead: their L2 latency hides under the previous step's FMAs
        const int n_e4 = Epad >> 2, n_steps = w * n_e4;
        float wn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wn[u] = u < s.E ? __ldg(W + (size_t)u * s.F) : 0.f;
        for (int stp = 0; stp < n_steps; ++stp) {
            const int k = stp / n_e4, e4 = (stp - k * n_e4) << 2;
            float wv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) wv[u] = wn[u];
            if (stp + 1 < n_steps) {
                const int k1 = (stp + 1) / n_e4, e1 = ((stp + 1) - k1 * n_e4) << 2;
#pragma unroll
                for (int u = 0; u < 4; ++u) wn[u] = e1 + u < s.E ? __ldg(W + (size_t)(k1 * s.E + e1 + u) * s.F) : 0.f;
            }
            const float* xb = s_x + (p0 + k) * Epad + e4;
#pragma unroll
            for (int pi = 0; pi < NP; ++pi) {
#pragma unroll
                for (int g = 0; g < kCnnG; ++g) {
                    const float4 x = *reinterpret_cast<const float4*>(xb + (g * s.L + pi) * Epad);
                    acc[g][pi] = fmaf(x.x, wv[0], acc[g][pi]);
                    acc[g][pi] = fmaf(x.y, wv[1], acc[g][pi]);
                    acc[g][pi] = fmaf(x.z, wv[2], acc[g][pi]);
                    acc[g][pi] = fmaf(x.w, wv[3], acc[g][pi]);
                }
            }
        }
    }
    float best[kCnnG];
    int bpos[kCnnG];
    const float bias = active ? conv_b[wi * s.F + f] : 0.f;
#pragma unroll
    for (int g = 0; g < kCnnG; ++g) {
        best[g] = -1.f;                                       // relu output >= 0: the first position always wins over -1
        bpos[g] = p0;
#pragma unroll
        for (int pi = 0; pi < NP; ++pi) {
            const float v = fmaxf(acc[g][pi] + bias, 0.f);                       // bias_add, relu, reduce_max (Char_CNN.py:50-58)
            if (v > best[g]) { best[g] = v; bpos[g] = p0 + pi; }
        }
        if (active && half) { s_best[g * s.F + f] = best[g]; s_bpos[g * s.F + f] = bpos[g]; }
    }
    __syncthreads();
    if (active && !half) {
        const int j = wi * s.F + f;
#pragma unroll
        for (int g = 0; g < kCnnG; ++g) {
            const int b = b0 + g;
            if (b >= B) continue;
            const float v1 = s_best[g * s.F + f];
            if (v1 > best[g]) { best[g] = v1; bpos[g] = s_bpos[g * s.F + f]; }   // ties keep the earlier position, like a scan
            feat[(size_t)b * D + j] = best[g];
            argpos[(size_t)b * D + j] = static_cast<unsigned char>(bpos[g]);
            // dropout (Char_CNN.py:67) and the bf16 operand copies [bpad, 512] / [512, bpad]
            const bool keep = philox_keep(seed, kStreamTitle, step, static_cast<uint32_t>(b + row_offset),
                                          static_cast<uint32_t>(j), kp_t);
            const __nv_bfloat16 hb = __float2bfloat16(keep ? __fdiv_rn(best[g], kp_t) : 0.f);
            feat_d[(size_t)b * kTitleFpad + j] = hb;
            feat_dT[(size_t)j * bpad + b] = hb;
        }
    }
}

// ONE launch for all widths (blockIdx.y = width, widest first: its CTAs take longest): 4 x 64 CTAs fill the GPU, where one
// launch per width left more than half of the SMs idle.  NP is dispatched per block (uniform) to the templated body.
__global__ void __launch_bounds__(256)
k_charcnn_fwd_tiled(const long long* __restrict__ titles, const float* __restrict__ emb, const float* __restrict__ conv_W,
                    const float* __restrict__ conv_b, const CnnShape s, float* __restrict__ feat,
                    unsigned char* __restrict__ argpos, __nv_bfloat16* __restrict__ feat_d, __nv_bfloat16* __restrict__ feat_dT,
                    int B, int bpad, float kp_t, unsigned long long seed, unsigned long long step, int row_offset) {
    extern __shared__ __align__(16) float s_x[];   // [kCnnG][L][Epad]: embedded titles, pad ids zero; then the halves' maxima
    const int wi = s.n_widths - 1 - blockIdx.y;
    const int np = (s.L - s.width[wi] + 2) / 2;
#define CNN_CASE(n) case n: charcnn_fwd_body<n>(titles, emb, conv_W, conv_b, s, wi, feat, argpos, feat_d, feat_dT, B, bpad, kp_t, seed, step, row_offset, s_x); break;
    switch (np) {
        CNN_CASE(1) CNN_CASE(2) CNN_CASE(3) CNN_CASE(4) CNN_CASE(5) CNN_CASE(6) CNN_CASE(7) CNN_CASE(8) CNN_CASE(9)
        CNN_CASE(10) CNN_CASE(11) CNN_CASE(12) CNN_CASE(13)
        default: break;
    }
#undef CNN_CASE
}

void launch_charcnn_fwd(const CnnFwdArgs& a, cudaStream_t st) {
    const int D = a.shape.F * a.shape.n_widths;
    bool tiled = a.shape.F <= 128;
    for (int i = 0; i < a.shape.n_widths; ++i) {
        const int np = (a.shape.L - a.shape.width[i] + 2) / 2;
        tiled = tiled && np >= 1 && np <= 13;
    }
    if (tiled) {
        // padding rows / columns of the operand copies are zero from allocation and never written (api_title.cu re-zeroes
        // them when the batch size changes); widest filters first: their CTAs take longest
        const int Epad = (a.shape.E + 3) & ~3;
        const size_t smem = sizeof(float) * (kCnnG * a.shape.L * Epad + kCnnG * a.shape.F) + sizeof(int) * kCnnG * a.shape.F;
        k_charcnn_fwd_tiled<<<dim3((a.B + kCnnG - 1) / kCnnG, a.shape.n_widths), 256, smem, st>>>(
            a.titles, a.emb, a.conv_W, a.conv_b, a.shape, a.feat, a.argpos, a.feat_d, a.feat_dT, a.B, a.bpad, a.kp_t, a.seed,
            a.step, a.row_offset);
        return;
    }
    const int threads = ((D > kTitleFpad ? D : kTitleFpad) + 31) / 32 * 32;
    k_charcnn_fwd<<<a.bpad, threads, sizeof(float) * a.shape.L * a.shape.E, st>>>(
        a.titles, a.emb, a.conv_W, a.conv_b, a.shape, a.feat, a.argpos, a.feat_d, a.feat_dT, a.B, a.bpad, a.kp_t, a.seed,
        a.step, a.row_offset);
}

// x_count = s * kp_in ; w_t = u / (u + x_count + 1e-10) ; w_p = x_count / (u + x_count + 1e-10)     DAEs.py:159-162
__global__ void k_mix_weights(const float* __restrict__ rowsum, const float* __restrict__ titles_use, float kp_in,
                              int B, int bpad, float* __restrict__ w_t, float* __restrict__ w_p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bpad) return;
    float wt = 0.f, wp = 0.f;
    if (b < B) {
        const float xc = rowsum[b] * kp_in, u = titles_use[b];
        const float deno = (u + xc) + kEpsLog;
        wt = __fdiv_rn(u, deno);
        wp = __fdiv_rn(xc, deno);
    }
    w_t[b] = wt;
    w_p[b] = wp;
}
void launch_mix_weights(const float* rowsum, const float* titles_use, float kp_in, int B, int bpad, float* w_t,
                        float* w_p, cudaStream_t st) {
    k_mix_weights<<<(bpad + 127) / 128, 128, 0, st>>>(rowsum, titles_use, kp_in, B, bpad, w_t, w_p);
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
// d[b, j] = (sum of the split-K partials of dfeat_d[b, j]) * keep / kp_t * (feat > 0)      (dropout, ReLU, max)
__global__ void k_title_dfeat(const float* __restrict__ partial, int nsplit, int bpad, const float* __restrict__ feat,
                              int B, int D, float kp_t, unsigned long long seed, unsigned long long step,
                              int row_offset, float* __restrict__ d) {
    const int b = blockIdx.x, j = threadIdx.x;
    if (j >= D) return;
    const int half = j / 256, c = j - half * 256;                             // two 256-column launches of the dh kernel
    const float* p = partial + (size_t)half * nsplit * bpad * 256 + (size_t)b * 256 + c;
    float s0 = 0.f, s1 = 0.f;
    int sidx = 0;
    for (; sidx + 2 <= nsplit; sidx += 2) {                                    // fixed order: deterministic
        s0 += p[(size_t)sidx * bpad * 256];
        s1 += p[(size_t)(sidx + 1) * bpad * 256];
    }
    for (; sidx < nsplit; ++sidx) s0 += p[(size_t)sidx * bpad * 256];
    const bool keep = philox_keep(seed, kStreamTitle, step, static_cast<uint32_t>(b + row_offset), static_cast<uint32_t>(j), kp_t);
    const float f = feat[(size_t)b * D + j];
    d[(size_t)b * D + j] = (keep && f > 0.f) ? (s0 + s1) * __fdiv_rn(1.f, kp_t) : 0.f;
}

// dW[k][e][f] = sum_b x[b, arg[b,f] + k, e] * d[b, f];  db[f] = sum_b d[b, f].  One CTA per feature: no atomics.
// Everything the inner loop touches is staged in shared memory first (the feature's gradient column, the character ids
// under its arg-max windows, the embedding table): the first version chased titles -> embedding through global memory
// once per title and thread (124 us per step); summation order over b unchanged.
__global__ void __launch_bounds__(512)
k_charcnn_bwd_w(const long long* __restrict__ titles, const float* __restrict__ emb, const float* __restrict__ d,
                const unsigned char* __restrict__ argpos, const CnnShape s, int B, float* __restrict__ g_W,
                float* __restrict__ g_b) {
    extern __shared__ __align__(16) float s_mem[];
    float* s_emb = s_mem;                                       // [C][E]
    float* s_d = s_emb + s.C * s.E;                             // [B]
    short* s_id = reinterpret_cast<short*>(s_d + B);            // [B][w]
    const int j = blockIdx.x;
    const int D = s.F * s.n_widths;
    const int wi = j / s.F, f = j - wi * s.F;
    const int w = s.width[wi];
    const int t = threadIdx.x;
    for (int i = t; i < s.C * s.E; i += blockDim.x) s_emb[i] = emb[i];
    for (int b = t; b < B; b += blockDim.x) s_d[b] = d[(size_t)b * D + j];
    for (int i = t; i < B * w; i += blockDim.x) {
        const int b = i / w, k = i - b * w;
        const long long id = titles[(size_t)b * s.L + argpos[(size_t)b * D + j] + k];
        s_id[i] = (id >= 0 && id < s.C) ? (short)id : (short)-1;
    }
    __syncthreads();
    const int k = t / s.E, e = t - k * s.E;                     // (k, e) pair
    float acc = 0.f, accb = 0.f;
    for (int b = 0; b < B; ++b) {
        const float dv = s_d[b];
        if (dv == 0.f) continue;                                // uniform over the CTA
        accb += dv;
        if (k < w) {
            const int id = s_id[b * w + k];
            if (id >= 0) acc = fmaf(s_emb[id * s.E + e], dv, acc);
        }
    }
    if (k < w) g_W[s.w_off[wi] + (size_t)(k * s.E + e) * s.F + f] = acc;
    if (t == 0) g_b[wi * s.F + f] = accb;
}

// conv_W [k][e][f] per width -> conv_WT [f][k][e] (e contiguous): the layout the embedding backward reads coalesced
__global__ void k_conv_transpose(const float* __restrict__ W, float* __restrict__ WT, const CnnShape s) {
    const int wi = blockIdx.y;
    const int w = s.width[wi], n = w * s.E * s.F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int f = i / (w * s.E), ke = i - f * (w * s.E);
        WT[s.w_off[wi] + i] = W[s.w_off[wi] + (size_t)ke * s.F + f];
    }
}
void launch_conv_transpose(const float* W, float* WT, const CnnShape& s, cudaStream_t st) {
    k_conv_transpose<<<dim3(32, s.n_widths), 256, 0, st>>>(W, WT, s);
}

// dx[b, pos, e] = sum_{j, k : arg[b,j] + k == pos} W[k][e][f] * d[b, j]: the gradient at the embedded characters of title b.
// One CTA per title, 8 feature groups x 64 embedding lanes: group jg walks features jg, jg + 8, ... and accumulates into
// ITS OWN [L][E] slab of shared memory (thread (jg, e) is the only writer of its column: no conflicts, no atomics); the
// slabs are added in group order.  Weights come from the transposed copy (coalesced over e).
constexpr int kEmbGroups = 8;
__global__ void __launch_bounds__(512)
k_charcnn_bwd_emb(const float* __restrict__ conv_WT, const float* __restrict__ d, const unsigned char* __restrict__ argpos,
                  const CnnShape s, float* __restrict__ dx) {
    extern __shared__ __align__(16) float s_mem[];
    const int LE = s.L * s.E;
    float* s_dx = s_mem;                                        // [kEmbGroups][L][E]
    float* s_d = s_dx + kEmbGroups * LE;                        // [D]
    const int b = blockIdx.x;
    const int D = s.F * s.n_widths;
    for (int i = threadIdx.x; i < kEmbGroups * LE; i += blockDim.x) s_dx[i] = 0.f;
    for (int j = threadIdx.x; j < D; j += blockDim.x) s_d[j] = d[(size_t)b * D + j];
    __syncthreads();
    const int e = threadIdx.x & 63, jg = threadIdx.x >> 6;
    if (e < s.E) {
        float* my = s_dx + jg * LE + e;
        for (int j = jg; j < D; j += kEmbGroups) {
            const float dv = s_d[j];
            if (dv == 0.f) continue;                            // uniform over the group's two warps
            const int wi = j / s.F, f = j - wi * s.F;
            const int w = s.width[wi];
            const int p0 = argpos[(size_t)b * D + j];
            const float* WT = conv_WT + s.w_off[wi] + (size_t)f * w * s.E + e;
            for (int k = 0; k < w; ++k) my[(p0 + k) * s.E] = fmaf(__ldg(WT + k * s.E), dv, my[(p0 + k) * s.E]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LE; i += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < kEmbGroups; ++g) v += s_dx[g * LE + i];
        dx[(size_t)b * LE + i] = v;
    }
}

// g_emb[c, e] = sum over (b, pos) with title[b, pos] == c of dx[b, pos, e]: no atomics, bit-reproducible.  One CTA per
// character: the ids are staged in shared memory, each of the 8 warps takes one contiguous eighth of the (b, pos) range
// (ballot over 32 positions at a time, matches added in ascending order, lane e and e + 32, ... own the embedding columns),
// and the 8 partial rows are added in warp order.
__global__ void __launch_bounds__(256)
k_charcnn_emb_reduce(const long long* __restrict__ titles, const float* __restrict__ dx, int n_pos, int E,
                     float* __restrict__ g_emb) {
    extern __shared__ int s_ids[];                 // [n_pos] then [8][128] partial sums
    float* s_part = reinterpret_cast<float*>(s_ids + ((n_pos + 31) & ~31));
    const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n_pos; i += blockDim.x) s_ids[i] = (int)titles[i];
    __syncthreads();
    const int seg = (((n_pos + 7) / 8) + 31) & ~31;
    const int lo = warp * seg, hi = min(n_pos, lo + seg);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};           // E <= 128
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        unsigned hits = __ballot_sync(0xffffffffu, i < hi && s_ids[i] == c);
        while (hits) {
            const int j = i0 + __ffs(hits) - 1;
            hits &= hits - 1;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (lane + 32 * u < E) acc[u] += dx[(size_t)j * E + lane + 32 * u];
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) s_part[warp * 128 + lane + 32 * u] = acc[u];
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) v += s_part[w8 * 128 + e];
        g_emb[(size_t)c * E + e] = v;
    }
}

void launch_charcnn_bwd(const CnnBwdArgs& a, cudaStream_t st) {
    const int D = a.shape.F * a.shape.n_widths;
    k_title_dfeat<<<a.B, (D + 31) / 32 * 32, 0, st>>>(a.dh_partial, a.nsplit, a.bpad, a.feat, a.B, D, a.kp_t, a.seed, a.step,
                                                      a.row_offset, a.d);
    int maxw = 0;
    for (int i = 0; i < a.shape.n_widths; ++i) maxw = a.shape.width[i] > maxw ? a.shape.width[i] : maxw;
    const size_t smem_w = sizeof(float) * (a.shape.C * a.shape.E + a.B) + sizeof(short) * (size_t)a.B * maxw + 16;
    k_charcnn_bwd_w<<<D, (maxw * a.shape.E + 31) / 32 * 32, smem_w, st>>>(a.titles, a.emb, a.d, a.argpos, a.shape, a.B, a.g_conv_W,
                                                                          a.g_conv_b);
    const size_t smem_e = sizeof(float) * (kEmbGroups * a.shape.L * a.shape.E + D);
    k_charcnn_bwd_emb<<<a.B, 512, smem_e, st>>>(a.conv_WT, a.d, a.argpos, a.shape, a.dx);
    const int n_pos = a.B * a.shape.L;
    k_charcnn_emb_reduce<<<a.shape.C, 256, sizeof(int) * ((n_pos + 31) & ~31) + sizeof(float) * 8 * 128, st>>>(
        a.titles, a.dx, n_pos, a.shape.E, a.g_emb);
}

// ------------------------------------------------------------------------------------------
// tf.contrib.layers.xavier_initializer(uniform=False) [TF1]: truncated normal, stddev sqrt(2.6 / (fan_in + fan_out)),
// samples beyond two standard deviations redrawn (Char_CNN.py:19, :45-47, :71-73).  Philox -> Box-Muller.
// ------------------------------------------------------------------------------------------
// Columns [col0, col0 + ncols) of a [rows, row_len] variable, stored at w[r * ld + (c - col0)]; keyed by the element's
// index in the WHOLE variable (r * row_len + c): the values do not depend on how the variable is split into blocks.
__global__ void k_trunc_normal(float* __restrict__ w, long long n, int row_len, int col0, int ncols, int ld, float stddev,
                               unsigned long long seed, unsigned stream_id) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const long long r = e / ncols, c = col0 + (e - r * ncols);
        const long long i = r * row_len + c;
        float z = 0.f;
        for (unsigned long long attempt = 0; attempt < 64; ++attempt) {
            const float u1 = philox_uniform24(seed, stream_id, 2 * attempt, static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(i));
            const float u2 = philox_uniform24(seed, stream_id, 2 * attempt + 1, static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(i));
            z = sqrtf(-2.f * logf(u1 + 5.9604645e-8f)) * cospif(2.f * u2);
            if (fabsf(z) <= 2.f) break;
            z = 0.f;
        }
        w[r * ld + (c - col0)] = z * stddev;
    }
}
void launch_trunc_normal(float* w, long long rows, int row_len, int ld, float stddev, unsigned long long seed,
                         unsigned stream_id, cudaStream_t st, int col0, int ncols) {
    if (ncols <= 0) { col0 = 0; ncols = row_len; }
    k_trunc_normal<<<1184, 256, 0, st>>>(w, rows * ncols, row_len, col0, ncols, ld, stddev, seed, stream_id);
}

// fp32 [rows, ld] -> bf16 [rows, ld] (operand copy of the output layer)
__global__ void k_cast_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16(src[i]);
}
void launch_cast_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st) {
    k_cast_bf16<<<1184, 256, 0, st>>>(src, dst, n);
}
// dense fp32 block [rows, cols] -> columns [col0, col0 + cols) of the bf16 operand [rows, ld]
__global__ void k_cast_block_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n, int cols, int ld, int col0) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long r = i / cols;
        dst[r * ld + col0 + (i - r * cols)] = __float2bfloat16(src[i]);
    }
}
void launch_cast_block_bf16(const float* src, __nv_bfloat16* dst, long long rows, int cols, int ld, int col0, cudaStream_t st) {
    if (cols > 0) k_cast_block_bf16<<<1184, 256, 0, st>>>(src, dst, rows * cols, cols, ld, col0);
}

// [D, N] (the reference's Output_W layout, Char_CNN.py:72) <-> item-major [N, ld] with zero padding columns
__global__ void k_transpose_pad(const float* __restrict__ src, float* __restrict__ dst, int D, int N, int ld, int to_item_major) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    if (to_item_major) {
        const int n = n0 + threadIdx.x, d = d0 + threadIdx.y;
        for (int yy = 0; yy < 32; yy += 8) {
            const int dd = d + yy;
            tile[threadIdx.y + yy][threadIdx.x] = (dd < D && n < N) ? src[(size_t)dd * N + n] : 0.f;
        }
        __syncthreads();
        for (int yy = 0; yy < 32; yy += 8) {
            const int nn = n0 + threadIdx.y + yy, dd = d0 + threadIdx.x;
            if (nn < N && dd < ld) dst[(size_t)nn * ld + dd] = tile[threadIdx.x][threadIdx.y + yy];
        }
    } else {
        for (int yy = 0; yy < 32; yy += 8) {
            const int nn = n0 + threadIdx.y + yy, dd = d0 + threadIdx.x;
            tile[threadIdx.y + yy][threadIdx.x] = (nn < N && dd < D) ? src[(size_t)nn * ld + dd] : 0.f;
        }
        __syncthreads();
        for (int yy = 0; yy < 32; yy += 8) {
            const int dd = d0 + threadIdx.y + yy, nn = n0 + threadIdx.x;
            if (dd < D && nn < N) dst[(size_t)dd * N + nn] = tile[threadIdx.x][threadIdx.y + yy];
        }
    }
}
void launch_transpose_pad(const float* src, float* dst, int D, int N, int ld, int to_item_major, cudaStream_t st) {
    const int dcols = to_item_major ? ld : D;
    k_transpose_pad<<<dim3((N + 31) / 32, (dcols + 31) / 32), dim3(32, 8), 0, st>>>(src, dst, D, N, ld, to_item_major);
}

void preload_title_cnn() {
    cudaFuncAttributes a;
    PRELOAD_KERNEL(k_charcnn_fwd);
    PRELOAD_KERNEL(k_charcnn_fwd_tiled);
    PRELOAD_KERNEL(k_conv_transpose);
    cudaFuncSetAttribute(k_charcnn_bwd_emb, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    PRELOAD_KERNEL(k_mix_weights);
    PRELOAD_KERNEL(k_title_dfeat);
    PRELOAD_KERNEL(k_charcnn_bwd_w);
    PRELOAD_KERNEL(k_charcnn_bwd_emb);
    PRELOAD_KERNEL(k_charcnn_emb_reduce);
    PRELOAD_KERNEL(k_trunc_normal);
    PRELOAD_KERNEL(k_cast_bf16);
    PRELOAD_KERNEL(k_cast_block_bf16);
    PRELOAD_KERNEL(k_transpose_pad);
    (void)cudaGetLastError();
}

}  // namespace dae
