// Seed-masked top-K ranking (K6).
//
// Restates utils/metrics.py:58-68 (single_eval) and main_runner/main_challenge.py:26-36
// (cand_generate): argsort(-scores[:T]) -> drop every seed id -> first K.  The reference's argsort
// is unstable, so ties are undefined upstream; the canonical order here is (score desc, index asc).
//
// One CTA per playlist row.  Exact 3-pass radix select (12+12+8 bits of the order-preserving key)
// finds the (K + #seeds)-th largest score, the survivors (<= 2048) are sorted in shared memory,
// seeds are removed and the first K are written.  Nothing of size T is ever sorted.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "kernels.h"

namespace dae {

constexpr int kTopkThreads = 1024;
constexpr int kCandMax = 2048;

__device__ __forceinline__ uint32_t score_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // monotone: larger score -> larger key
}
__device__ __forceinline__ float key_score(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}

// sigmoid_out: the row holds logits and is ranked by p = sigmoid(z) in fp32 -- the SAME expression as the decode
// epilogue (gemm_sm100.cu), so that the fused decode + top-K orders exactly like the dense path / the reference
// (distinct logits may round to the same p; ties then go by index)
__device__ __forceinline__ float load_score(const float* __restrict__ x, int i, int sigmoid_out) {
    const float v = x[i];
    return sigmoid_out ? __fdividef(1.f, 1.f + __expf(-v)) : v;
}

// The final list of the fused decode + top-K is ordered by p = sigmoid(z) in fp32 (ties: lower id first), its filter
// compares LOGITS: distinct logits collapse onto one p once 1 - p nears the fp32 resolution (and onto exactly 1.0f from
// z ~ 16.6 on), so an item just below the K-th logit can tie in p with kept items and win on its id.  The filter threshold
// derived from a K-th largest logit t therefore keeps everything whose p could equal the threshold item's:
// |dz| p (1 - p) < 2 ulp(p)  ->  dz < ~2.4e-7 (1 + e^z), 4x margin for the approximate exp / division; from z = 15 on
// simply everything >= 15 (a list that overflows falls back to the dense path).
__device__ __forceinline__ float filter_threshold(float t) {
    if (t > 15.f) return 15.f;
    if (t > -CUDART_INF_F) return t - (1e-6f * (1.f + __expf(t)) + 1e-6f * fabsf(t));
    return t;
}

// histogram increment with warp aggregation (scores cluster in a few bins -> avoid same-address atomics)
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t digit, bool active) {
    const uint32_t amask = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const uint32_t peers = __match_any_sync(amask, digit);
    if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[digit], __popc(peers));
}

// Whole block: largest bin b with sum_{d>=b} hist[d] >= want.  Returns via smem: s_out[0]=b, s_out[1]=sum_{d>b}.
__device__ void select_bin(const uint32_t* hist, int nbins, uint32_t want, uint32_t* s_warp, uint32_t* s_out) {
    const int t = threadIdx.x;
    const int per = nbins / kTopkThreads > 0 ? nbins / kTopkThreads : 1;
    const int nthreads_used = nbins / per;
    // thread t owns bins [hi - per + 1, hi] counted from the top
    uint32_t mine = 0;
    const int top = nbins - 1 - t * per;
    if (t < nthreads_used)
        for (int j = 0; j < per; ++j) mine += hist[top - j];
    // inclusive scan over threads (thread 0 = highest bins)
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < (t >> 5); ++w) base += s_warp[w];
    incl += base;
    const uint32_t excl = incl - mine;
    if (t < nthreads_used && excl < want && incl >= want) {
        uint32_t run = excl;
        for (int j = 0; j < per; ++j) {
            const uint32_t h = hist[top - j];
            if (run + h >= want) { s_out[0] = top - j; s_out[1] = run; break; }
            run += h;
        }
    }
    __syncthreads();
}

// Whole block: the want-th largest of n <= 2 * kTopkThreads keys held in shared memory at keys[i * stride] (radix passes of
// 12 + 12 + 8 bits); 0 when want > n.  passes == 2 stops after 24 bits and returns them with the low 8 bits clear: a lower
// bound of the want-th largest (within 2^-15 relative), which is all a pre-threshold needs.
__device__ uint32_t select_kth_smem(const uint32_t* keys, int stride, int n, uint32_t want, int passes, uint32_t* s_hist,
                                    uint32_t* s_warp, uint32_t* s_out) {
    const int t = threadIdx.x;
    if (want > (uint32_t)n) return 0u;
    uint32_t prefix = 0;
#pragma unroll 1
    for (int pass = 0; pass < passes; ++pass) {
        const int sh = pass == 0 ? 20 : (pass == 1 ? 8 : 0);
        const int nb = pass == 2 ? 256 : 4096;
        for (int i = t; i < nb; i += kTopkThreads) s_hist[i] = 0;
        __syncthreads();
        for (int i = t; i < 2 * kTopkThreads; i += kTopkThreads) {
            bool act = i < n;
            const uint32_t key = act ? keys[(size_t)i * stride] : 0u;
            if (pass == 1) act = act && ((key >> 20) == prefix);
            if (pass == 2) act = act && ((key >> 8) == prefix);
            hist_add(s_hist, (key >> sh) & (uint32_t)(nb - 1), act);
        }
        __syncthreads();
        select_bin(s_hist, nb, want, s_warp, s_out);
        want -= s_out[1];
        prefix = (pass == 0) ? s_out[0] : ((prefix << (pass == 1 ? 12 : 8)) | s_out[0]);
        __syncthreads();
    }
    return passes == 2 ? prefix << 8 : prefix;
}

// Descending bitonic sort of s[0..npow) (npow a power of two, 32 <= npow <= 2 * kTopkThreads), whole block.  Thread t keeps
// elements t (and t + 1024: TWO) in registers: compare-exchange distances below 32 are warp shuffles, the distance 1024 is
// inside the thread, the others go through shared memory with ONE barrier each (two buffers alternate; s2: npow entries of
// scratch).  The network is fully unrolled: directions and distances are immediates, ~9 instructions per step.
template <bool TWO>
__device__ __forceinline__ void bitonic_desc_t(unsigned long long* s, unsigned long long* s2, int npow) {
    const int t = threadIdx.x;
    unsigned long long v0 = s[t], v1 = TWO ? s[t + kTopkThreads] : 0ull;
    auto pick = [](unsigned long long a, unsigned long long b, bool take_max) { return ((a < b) == take_max) ? b : a; };
    int cur = 1;
#pragma unroll
    for (int lk = 1; lk <= (TWO ? 11 : 10); ++lk) {
        const int k = 1 << lk;
        if (k > npow) break;                              // block-uniform
#pragma unroll
        for (int lj = lk - 1; lj >= 0; --lj) {
            const int j = 1 << lj;
            if (TWO && j == kTopkThreads) {               // elements t and t + 1024 of the same thread (k == 2048: descending)
                const unsigned long long a = v0, b = v1;
                v0 = pick(a, b, true);
                v1 = pick(b, a, false);
                continue;
            }
            unsigned long long p0, p1 = 0ull;
            if (j >= 32) {
                unsigned long long* buf = cur ? s2 : s;
                cur ^= 1;
                buf[t] = v0;
                if (TWO) buf[t + kTopkThreads] = v1;
                __syncthreads();
                p0 = buf[t ^ j];
                if (TWO) p1 = buf[(t ^ j) + kTopkThreads];
            } else {
                p0 = __shfl_xor_sync(0xffffffffu, v0, j);
                if (TWO) p1 = __shfl_xor_sync(0xffffffffu, v1, j);
            }
            // element i: descending block when (i & k) == 0, lower partner when (i & j) == 0; it keeps the max iff both agree
            const bool max0 = (((t >> lk) ^ (t >> lj)) & 1) == 0;
            v0 = pick(v0, p0, max0);
            if (TWO) v1 = pick(v1, p1, lk == 10 ? !max0 : max0);      // i = t + 1024: only bit 10 differs (k == 1024 flips)
        }
    }
    __syncthreads();                                      // the last shared-memory step may still be read by slower warps
    s[t] = v0;
    if (TWO) s[t + kTopkThreads] = v1;
    __syncthreads();
}
__device__ __noinline__ void bitonic_desc(unsigned long long* s, unsigned long long* s2, int npow) {
    if (npow > kTopkThreads) bitonic_desc_t<true>(s, s2, npow);
    else bitonic_desc_t<false>(s, s2, npow);
}

__global__ void __launch_bounds__(kTopkThreads, 2)     // 2 CTAs / SM: the kernel is latency-bound (block scans between passes)
k_topk(const float* __restrict__ scores, long long ld, int T, int K, const int* __restrict__ seed_ptr,
       const int* __restrict__ seed_idx, int idx_base, int* __restrict__ out_idx, float* __restrict__ out_score,
       const int* __restrict__ remap, const int* __restrict__ row_n, int sigmoid_out, float* __restrict__ thr_out, int thr_div) {
    __shared__ __align__(16) uint32_t s_hist[4096];
    __shared__ __align__(16) unsigned long long s_cand[kCandMax];
    __shared__ int s_seed[kCandMax];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_out[2];
    __shared__ uint32_t s_cnt[2];
    const int row = blockIdx.x;
    const int t = threadIdx.x;
    const float* x = scores + (size_t)row * ld;
    // candidate-list mode: the row holds row_n[row] (item, logit) pairs in arbitrary order; entry i is item rm[i]
    const int* rm = remap != nullptr ? remap + (size_t)row * ld : nullptr;
    if (row_n != nullptr) T = min(row_n[row], (int)ld);

    int nseed = 0, sb = 0;
    if (seed_ptr != nullptr) {
        sb = seed_ptr[row];
        nseed = min(seed_ptr[row + 1] - sb, kCandMax - K > 0 ? kCandMax - K : 0);
    }
    // threshold-only mode: this shard's share of the row's k + #seeds (every shard finds that many above ITS threshold, so
    // at least k + #seeds items of the catalogue exceed the minimum of the shards' thresholds)
    const int Kreq = thr_out != nullptr ? (K + nseed + thr_div - 1) / thr_div : K + nseed;
    uint32_t want = (uint32_t)min(Kreq, T);
    if (want > (uint32_t)kCandMax) want = kCandMax;
    const uint32_t want0 = want;
    const int ncand = (int)want0;

    // ---- fast path: ONE scan for a pre-threshold, one to gather what passes it, then a sort of that small set ----------
    // The row is cut into 2048 groups (thread t: elements t + 1024 j, even j -> group t, odd j -> group t + 1024).  The
    // want-th largest of the group MAXIMA is a lower bound of the want-th largest element (each of those want groups holds
    // an element >= it), and close to it while want << 2048: everything >= it (typically 1.0-1.4 x want entries) is
    // gathered into shared memory and sorted by (score desc, id asc) -- the first `want` entries are exactly what the three
    // radix passes + ordered tie collection below produce.  More than 2048 gathered entries (heavy ties, want near 2048):
    // the general path below redoes the row.
    bool sorted = false;
    if (want0 > 0 && want0 <= (uint32_t)kTopkThreads) {
        uint32_t t0 = 0u;                                   // pre-threshold key; 0 = keep everything (T <= 2048)
        if (T > kCandMax) {
            uint32_t m0 = 0u, m1 = 0u;
            int i = t;
            for (; i + 7 * kTopkThreads < T; i += 8 * kTopkThreads) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = x[i + u * kTopkThreads];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t key = score_key(sigmoid_out ? __fdividef(1.f, 1.f + __expf(-v[u])) : v[u]);
                    if (u & 1) m1 = max(m1, key); else m0 = max(m0, key);
                }
            }
            for (int u = 0; i < T; i += kTopkThreads, ++u) {
                const uint32_t key = score_key(load_score(x, i, sigmoid_out));
                if (u & 1) m1 = max(m1, key); else m0 = max(m0, key);
            }
            uint32_t* s_max = reinterpret_cast<uint32_t*>(s_seed);      // the seeds are loaded after the select
            s_max[t] = m0;
            s_max[t + kTopkThreads] = m1;
            __syncthreads();
            t0 = select_kth_smem(s_max, 1, 2 * kTopkThreads, want0, 2, s_hist, s_warp, s_out);
        }
        if (t == 0) s_cnt[0] = 0;
        __syncthreads();
        const int Tround4 = (T + 4 * kTopkThreads - 1) / (4 * kTopkThreads) * (4 * kTopkThreads);
        for (int i0 = t; i0 < Tround4; i0 += 4 * kTopkThreads) {
            uint32_t key[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kTopkThreads;
                key[u] = i < T ? score_key(load_score(x, i, sigmoid_out)) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kTopkThreads;
                const bool hit = i < T && key[u] >= t0;
                const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                if (bal) {
                    const int lane = t & 31, leader = __ffs(bal) - 1;
                    uint32_t base = 0;
                    if (lane == leader) base = atomicAdd(&s_cnt[0], (uint32_t)__popc(bal));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
                    if (hit && slot < (uint32_t)kCandMax)
                        s_cand[slot] = ((unsigned long long)key[u] << 32) | (0xFFFFFFFFu - (uint32_t)(rm ? rm[i] : i));
                }
            }
        }
        __syncthreads();
        const uint32_t c = s_cnt[0];
        if (c <= (uint32_t)kCandMax) {
            if (thr_out != nullptr) {      // threshold only: the K-th largest of the gathered keys (the high words), no sort
                const uint32_t kth = T >= Kreq ? select_kth_smem(reinterpret_cast<const uint32_t*>(s_cand) + 1, 2, (int)c, want0, 3,
                                                                 s_hist, s_warp, s_out) : 0u;
                if (t == 0) thr_out[row] = filter_threshold(T >= Kreq ? key_score(kth) : -CUDART_INF_F);
                return;
            }
            int npow = 32;
            while (npow < (int)c) npow <<= 1;
            for (int i = (int)c + t; i < npow; i += kTopkThreads) s_cand[i] = 0ull;
            __syncthreads();
            bitonic_desc(s_cand, reinterpret_cast<unsigned long long*>(s_hist), npow);
            sorted = true;
        }
        __syncthreads();
    }
    if (!sorted) {

    // ---- pass 1..3: radix select of the want-th largest key -------------------------------
    uint32_t prefix = 0;      // selected high bits so far
    uint32_t above_total = 0; // elements strictly above the final threshold
    const int shifts[3] = {20, 8, 0};
    const int nbins[3] = {4096, 4096, 256};
    const uint32_t masks[3] = {0xFFFu, 0xFFFu, 0xFFu};
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = t; i < nbins[pass]; i += kTopkThreads) s_hist[i] = 0;
        __syncthreads();
        const int sh = shifts[pass];
        const int Tround = (T + kTopkThreads - 1) / kTopkThreads * kTopkThreads;
        for (int i = t; i < Tround; i += kTopkThreads) {
            bool act = i < T;
            uint32_t key = 0;
            if (act) key = score_key(load_score(x, i, sigmoid_out));
            if (pass == 1) act = act && ((key >> 20) == prefix);
            if (pass == 2) act = act && ((key >> 8) == prefix);
            hist_add(s_hist, (key >> sh) & masks[pass], act);
        }
        __syncthreads();
        select_bin(s_hist, nbins[pass], want, s_warp, s_out);
        const uint32_t bin = s_out[0], above = s_out[1];
        above_total += above;
        want -= above;
        prefix = (pass == 0) ? bin : ((prefix << (pass == 1 ? 12 : 8)) | bin);
        __syncthreads();
    }
    const uint32_t thr = prefix;          // full 32-bit key of the want0-th largest element
    const uint32_t need_eq = want;        // how many elements equal to thr are needed (>= 1)
    if (thr_out != nullptr) {
        // threshold-only mode (the intermediate passes of the fused decode + top-K): the K-th largest logit is all that
        // is needed -- no collection, no sort (half of this kernel's time).  Fewer than K entries: keep everything.
        if (t == 0) thr_out[row] = filter_threshold(T >= Kreq ? key_score(thr) : -CUDART_INF_F);
        return;
    }
    // Ties AT the cut go to the lowest item ids.  In a dense row the position is the id, and the ordered pick below does
    // it.  In candidate-list mode the list order is arbitrary: when more entries tie with the threshold than are needed
    // (scores near saturation), find the id cut-off X = the need_eq-th smallest id among them by bisection (31 counting
    // sweeps over <= 16 K entries; only in that rare case) and accept exactly the ties with id <= X.
    uint32_t id_cut = 0xFFFFFFFFu;
    const uint32_t eq_total = s_hist[thr & 0xFFu];      // the last pass's histogram: entries whose full key == thr
    if (rm != nullptr && eq_total > need_eq) {          // block-uniform
        uint32_t lo = 0, hi = 0x7FFFFFFFu;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            __syncthreads();
            if (t == 0) s_cnt[0] = 0;
            __syncthreads();
            uint32_t c = 0;
            for (int i = t; i < T; i += kTopkThreads)
                c += (score_key(load_score(x, i, sigmoid_out)) == thr && (uint32_t)rm[i] <= mid) ? 1u : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if ((t & 31) == 0 && c) atomicAdd(&s_cnt[0], c);
            __syncthreads();
            if (s_cnt[0] >= need_eq) hi = mid; else lo = mid + 1;
        }
        id_cut = lo;
        __syncthreads();
    }

    // ---- collect: all keys > thr (any order) + the need_eq lowest-index keys == thr ---------
    if (t == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
    __syncthreads();
    const int Tround = (T + kTopkThreads - 1) / kTopkThreads * kTopkThreads;
    uint32_t eq_seen = 0;                 // block-uniform running count of == thr elements
    for (int i = t; i < Tround; i += kTopkThreads) {
        uint32_t key = 0;
        const bool in = i < T;
        if (in) key = score_key(load_score(x, i, sigmoid_out));
        if (in && key > thr) {
            const uint32_t slot = atomicAdd(&s_cnt[0], 1u);
            if (slot < (uint32_t)kCandMax)
                s_cand[slot] = ((unsigned long long)key << 32) | (0xFFFFFFFFu - (uint32_t)(rm ? rm[i] : i));
        }
        if (eq_total == need_eq) {        // every entry equal to the threshold is needed (the common case: a unique K-th
            if (in && key == thr) {       // score): no order to respect among them, no block scans
                const uint32_t slot = (want0 - need_eq) + atomicAdd(&s_cnt[1], 1u);
                s_cand[slot] = ((unsigned long long)key << 32) | (0xFFFFFFFFu - (uint32_t)(rm ? rm[i] : i));
            }
        } else if (eq_seen < need_eq) {   // ordered pick among ties: chunk-ordered block scan
            const bool eq = in && key == thr && (rm == nullptr || (uint32_t)rm[i] <= id_cut);
            const uint32_t bal = __ballot_sync(0xffffffffu, eq);
            const uint32_t wrank = __popc(bal & ((1u << (t & 31)) - 1u));
            if ((t & 31) == 0) s_warp[t >> 5] = __popc(bal);
            __syncthreads();
            uint32_t base = 0, total = 0;
            for (int w = 0; w < 32; ++w) {
                const uint32_t c = s_warp[w];
                if (w < (t >> 5)) base += c;
                total += c;
            }
            const uint32_t rank = eq_seen + base + wrank;
            if (eq && rank < need_eq) {
                const uint32_t slot = (want0 - need_eq) + rank;   // ties go after the strictly-greater block
                s_cand[slot] = ((unsigned long long)key << 32) | (0xFFFFFFFFu - (uint32_t)(rm ? rm[i] : i));
            }
            eq_seen += total;
            __syncthreads();
        }
    }
    __syncthreads();
    // slots [0, want0-need_eq) hold keys > thr (count == above_total == want0-need_eq), then the ties
    (void)above_total;
    int npow = 32;
    while (npow < ncand) npow <<= 1;
    for (int i = ncand + t; i < npow; i += kTopkThreads) s_cand[i] = 0ull;
    __syncthreads();
    bitonic_desc(s_cand, reinterpret_cast<unsigned long long*>(s_hist), npow);   // on (key, ~idx): score desc, index asc
    }   // !sorted
    for (int i = t; i < nseed; i += kTopkThreads) s_seed[i] = seed_idx[sb + i] - idx_base;
    __syncthreads();
    // ---- drop seeds, ordered compaction, first K --------------------------------------------------
    // each thread owns 2 consecutive candidates (kCandMax / kTopkThreads)
    uint32_t keepf[2] = {0, 0};
    for (int u = 0; u < 2; ++u) {
        const int i = t * 2 + u;
        if (i < ncand) {
            const int idx = (int)(0xFFFFFFFFu - (uint32_t)(s_cand[i] & 0xFFFFFFFFull));
            bool is_seed = false;
            for (int s = 0; s < nseed; ++s) is_seed |= (s_seed[s] == idx);
            keepf[u] = is_seed ? 0u : 1u;
        }
    }
    uint32_t mine = keepf[0] + keepf[1];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    __syncthreads();
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    uint32_t base = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < (t >> 5)) base += c;
        total += c;
    }
    uint32_t pos = base + incl - mine;
    for (int u = 0; u < 2; ++u) {
        if (keepf[u]) {
            if (pos < (uint32_t)K) {
                const unsigned long long c = s_cand[t * 2 + u];
                out_idx[(size_t)row * K + pos] = (int)(0xFFFFFFFFu - (uint32_t)(c & 0xFFFFFFFFull)) + idx_base;
                out_score[(size_t)row * K + pos] = key_score((uint32_t)(c >> 32));
            }
            ++pos;
        }
    }
    for (int i = (int)total + t; i < K; i += kTopkThreads) {   // fewer than K candidates: pad
        out_idx[(size_t)row * K + i] = -1;
        out_score[(size_t)row * K + i] = -CUDART_INF_F;
    }
}

// Per-playlist filter threshold for the next pass of the fused decode + top-K: the kp-th largest logit found so far
// (a lower bound of the kp-th largest of any superset), -inf while fewer than kp candidates exist, +inf for the
// padding rows of the batch tile (they emit nothing).
__global__ void k_thr_from_topk(const float* __restrict__ score, const int* __restrict__ idx, int kp, int batch, int rows,
                                float* __restrict__ thr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float t = CUDART_INF_F;
    if (r < batch) {
        t = (score != nullptr && idx[(size_t)r * kp + kp - 1] >= 0) ? score[(size_t)r * kp + kp - 1] : -CUDART_INF_F;
        t = filter_threshold(t);
    }
    thr[r] = t;
}
void launch_thr_from_topk(const float* score, const int* idx, int kp, int batch, int rows, float* thr, cudaStream_t st) {
    k_thr_from_topk<<<(rows + 255) / 256, 256, 0, st>>>(score, idx, kp, batch, rows, thr);
}

// ------------------------------------------------------------------------------------------
// Challenge metrics of a ranked list (utils/metrics.py:20-49), one CTA per playlist:
//   r-precision  |set(answer) & set(cand[:len(answer)])| / len(answer)      metrics.py:25-27  (answers may hold -1:
//                never matched, still counted in the denominator)
//   ndcg         dcg / idcg with the reference's own IDCG: 1 + sum_{j=2..1+h} 1/log2(j), h = hits found after rank 0
//                (metrics.py:29-42)
//   clicks       first hit index // 10, 51 when there is none                metrics.py:44-49
// cand rows are -1 padded at the end (a list may be shorter than k).  Integer work is exact; the two quotients are
// formed in double precision like the Python code, the sums in rank order by one thread (<= k terms).
// ------------------------------------------------------------------------------------------
constexpr int kMetricsMaxAns = 2048;
__global__ void __launch_bounds__(128)
k_metrics(const int* __restrict__ cand, long long ld, int k, const int* __restrict__ ans_ptr, const int* __restrict__ ans_idx,
          double* __restrict__ out) {
    __shared__ int s_ans[kMetricsMaxAns];
    __shared__ unsigned int s_hit[32];                 // k <= 1024 ranks, one bit each
    __shared__ unsigned int s_rep[32];                 // hit whose id already occurred at an earlier rank (set semantics of r-precision)
    const int row = blockIdx.x, t = threadIdx.x;
    const int a0 = ans_ptr[row], A = ans_ptr[row + 1] - a0;
    const int An = A < kMetricsMaxAns ? A : kMetricsMaxAns;
    for (int i = t; i < An; i += blockDim.x) s_ans[i] = ans_idx[a0 + i];
    if (t < 32) { s_hit[t] = 0u; s_rep[t] = 0u; }
    __syncthreads();
    for (int i = t; i < k; i += blockDim.x) {
        const int c = cand[(size_t)row * ld + i];
        if (c < 0) continue;                           // padding (and -1 answers can never match a candidate)
        bool hit = false;
        for (int j = 0; j < An; ++j) hit |= (s_ans[j] == c);
        if (hit) {
            atomicOr(&s_hit[i >> 5], 1u << (i & 31));
            // r-precision intersects SETS (metrics.py:25): a repeated id counts once.  (Ranked lists never repeat an id;
            // the reference function is defined for any list.)
            bool rep = false;
            for (int j = 0; j < i; ++j) rep |= (cand[(size_t)row * ld + j] == c);
            if (rep) atomicOr(&s_rep[i >> 5], 1u << (i & 31));
        }
    }
    __syncthreads();
    if (t != 0) return;
    int n_cand = 0;
    while (n_cand < k && cand[(size_t)row * ld + n_cand] >= 0) ++n_cand;
    int r_hits = 0, first = -1, later = 0;
    double dcg = 0.0;
    for (int w = 0; w < (k + 31) / 32; ++w) {
        unsigned int m = s_hit[w];
        while (m) {
            const int i = w * 32 + __ffs(m) - 1;
            m &= m - 1;
            if (i < A && !((s_rep[w] >> (i & 31)) & 1u)) ++r_hits;
            if (first < 0) first = i;
            if (i == 0) dcg = 1.0;
            else { dcg += 1.0 / (log((double)(i + 1)) / log(2.0)); ++later; }
        }
    }
    double idcg = 1.0;
    for (int j = 2; j < 2 + later; ++j) idcg += 1.0 / (log((double)j) / log(2.0));
    out[(size_t)row * 3 + 0] = A > 0 ? (double)r_hits / (double)A : 0.0;
    out[(size_t)row * 3 + 1] = n_cand > 0 ? dcg / idcg : 0.0;
    out[(size_t)row * 3 + 2] = first >= 0 ? (double)(first / 10) : 51.0;
}
void launch_metrics(const int* cand, long long ld, int B, int k, const int* ans_ptr, const int* ans_idx, double* out,
                    cudaStream_t st) {
    k_metrics<<<B, 128, 0, st>>>(cand, ld, k, ans_ptr, ans_idx, out);
}

void launch_topk(const TopkArgs& a, cudaStream_t st) {
    k_topk<<<a.B, kTopkThreads, 0, st>>>(a.scores, a.ld, a.T, a.k, a.seed_ptr, a.seed_idx, a.idx_base, a.out_idx,
                                         a.out_score, a.remap, a.row_n, a.sigmoid_out, a.thr_out, a.thr_div > 0 ? a.thr_div : 1);
}

// Force the module / functions to load now: with CUDA's lazy loading the FIRST launch of a kernel may
// synchronise the context, which would deadlock against a cross-GPU flag barrier already spinning.
void preload_topk() {
    cudaFuncAttributes a;
    PRELOAD_KERNEL(k_topk);
    PRELOAD_KERNEL(k_thr_from_topk);
    PRELOAD_KERNEL(k_metrics);
    (void)cudaGetLastError();
}

}  // namespace dae
