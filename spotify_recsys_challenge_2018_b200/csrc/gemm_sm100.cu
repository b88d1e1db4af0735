// Tensor-core kernels of the DAE step for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators.
//
//   G1    k_itemtile<TRAIN|PREDICT|FILTER>   Z[item, b] = W_dec[item,:] . h_d[b,:]      (DAEs.py:75 / :143)
//           TRAIN   epilogue: +b_dec, sigmoid, weighted BCE (DAEs.py:98-99), d cost/dz (bf16, item-major),
//                             db_dec = sum_b dz, per-CTA loss partial.  The [B,N] score matrix is never written.
//           PREDICT epilogue: +b_dec, sigmoid (, title mix DAEs.py:180) -> y_pred[b, item] fp32 (or the raw logits)
//           FILTER  epilogue: logits >= the playlist's threshold -> its candidate list (fused decode + top-K)
//   G2    k_dw              dW_dec^T tile = h_d^T . dz^T -> g_dec (parity tests, title head, two-kernel path)
//   G2+K5 k_dw_adam_fused   the same tile kept in tensor memory, dense TF1 Adam applied from there; w / m / v staged
//                           through shared memory with 1-D bulk copies (the default decoder update)
//   G3    k_dh              dh_d[b,:] = sum_item dz[item,b] W_dec[item,:]  (split-K over items,
//                           both operands MN-major straight from their item-major layouts)
//
// Shape of every kernel: persistent, 1 CTA / SM,
//   warp 0  TMA producer (one elected lane)        -> smem ring, mbarrier full/empty
//   warp 1  TMEM allocator + MMA issuer (one lane) -> 2 x 256-column fp32 accumulators, tcgen05.commit
//   (warp 2 of k_dw_adam_fused: the optimizer-state I/O thread)
//   4 (k_dh) or 16 epilogue warps: tcgen05.ld 32 lanes x 16/32 columns per warp; lane == item row (G1), hidden unit
//   (G2) or batch row (G3)
// The small operand (h_d, 128 KB at B=256,H=256) stays resident in shared memory for the CTA's lifetime where it
// fits; only the catalogue-sized operand streams through the ring, so W / dz are read from HBM exactly once per kernel.
//
// Synchronisation rule that every kernel here obeys (learnt the hard way, see k_dw_adam_fused): an mbarrier parity wait
// is only meaningful if the waiter has observed the PREVIOUS phase of that barrier -- every waiter must see
// consecutive phases of every barrier it waits on.  tests/test_sync_protocol.py models it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "tmap.h"
#include "umma.cuh"

namespace dae {

// ------------------------------------------------------------------------------------------
// G1 / G2: item-tile kernels
// ------------------------------------------------------------------------------------------
enum { MODE_TRAIN = 0, MODE_PREDICT = 1, MODE_FILTER = 2 };

constexpr int kStages = 6;
constexpr int kStagesStreamB = 4;                    // G2, K > 256: stages of (16 KB item chunk + 32 KB second-operand chunk)
constexpr int kEpiWarps = 16;                        // 4 warps per TMEM lane quadrant, 32-column chunks dealt round-robin
constexpr int kItemThreads = 64 + 32 * kEpiWarps;
constexpr int kABytes = kTileItems * 128;           // one K-chunk of the streamed operand: 128 rows x 128 B
constexpr int kBChunkBytes = 256 * 128;             // one K-chunk of the resident operand: <=256 rows x 128 B
constexpr int kSmemB = 4 * kBChunkBytes;            // 131072
constexpr int kSmemA = kStages * kABytes;           // 98304
constexpr int kSmemBars = 512;
constexpr int kMaxRing = 12;                          // barrier slots of the streamed-operand ring (inference modes size it at run time)
constexpr int kSmemThr = kMaxBpad * 4;              // FILTER: per-playlist thresholds of the CTA's batch tile
// FILTER gives the last TWO ring stages (32 KB) to its epilogue warps, 2 KB each: a queue of accumulator rows that passed the
// scan's pre-test (kQRec records of 16 logits-to-be + bias + tag, 80 B apart) and two staging halves of kWHalf candidates
// (one fills while the other one's list appends are in flight)
constexpr int kFilterStgStages = 2;
constexpr int kWarpStgBytes = kFilterStgStages * kABytes / 16;       // 2048
constexpr int kQRec = 12;
constexpr int kQWords = 20;
constexpr int kWHalf = 64;
constexpr int kStgOff = 1024;                       // staging halves at +1024 / +1536 of the warp's 2 KB block (queue: 960 B at +0)
static_assert(kQRec * kQWords * 4 <= kStgOff && kStgOff + 2 * kWHalf * 8 <= kWarpStgBytes, "per-warp FILTER buffers");
constexpr int kSmemItemTile = kSmemB + kSmemA + kSmemBars + kSmemThr + 1024;  // + alignment slack
// TRAIN runs a 4-stage ring: the epilogue paces the kernel (one tile = 4 chunks is ahead of the MMA at any time), and
// the 32 KB it gives up is what lets the background optimizer streamer (optim.cu: k_adam_bg, 31 KB) share the SM
constexpr int kStagesTrain = 4;
constexpr int kSmemItemTileTrain = kSmemB + kStagesTrain * kABytes + kSmemBars + 1024;

struct ItemTileDev {
    int n_rows;       // rows of the streamed operand (TRAIN: this rank's item rows; PREDICT: catalogue columns kept)
    int n_global;     // catalogue size
    int tiles;        // 128-row tiles to process
    int kchunks;      // H / 64
    int n_cols;       // UMMA N: rows of one batch tile
    int batch;        // valid rows per batch tile
    int world, rank;  // TRAIN: local row -> catalogue id (tile-cyclic ownership)
    const float* bias;
    const uint32_t* ybits;
    int ywords;
    __nv_bfloat16* dzT;
    int ld_dz;
    float* db_dec;
    float* loss_partial;
    float inv_batch;
    float* out;
    long long ld_out;
    int n_out;
    const float* mix_wp;
    const float* mix_wt;
    const float* title_score;
    // FILTER: logits above the playlist's threshold are appended to its candidate list
    const float* thr;        // [n_batch_tiles * n_cols]
    float* cand_val;         // [n_batch_tiles * n_cols, cand_cap] logits z = h_d . W_dec[item] + b_dec[item]
    int* cand_idx;           // [.., cand_cap] catalogue ids
    int* cand_cnt;           // [..] candidates found (may exceed cand_cap: the caller checks)
    int cand_cap;
    int item0;               // catalogue id of row 0 of the streamed operand
    int raw_logits;          // PREDICT: write z instead of sigmoid(z)
    unsigned long long* trace;   // debug: %globaltimer when the first / last CTA of the grid started
    int pair;                // PREDICT / FILTER: launched as clusters of two batch tiles (grid.y) that share every W chunk
    int ring;                // PREDICT / FILTER: 16 KB stages behind the resident operand (FILTER gives the last two to its epilogue)
    int tune;                // FILTER switches (dae_model_set_debug bits 16..), TIMING EXPERIMENTS ONLY (wrong results):
                             // bit 1 = the epilogue releases every accumulator unread, bit 2 = the scan never queues a row
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// G1 TRAIN epilogue on 16 accumulator values of one item row (16 batch columns): p = sigmoid(z), the weighted BCE
// term (DAEs.py:98-99), d cost / d z packed to bf16.  `yw` bit j = y[b0 + j, item].  The hot loop keeps
// the cheap gradient form ratio = (1-p | p), which equals p(1-p)/((p | 1-p) + 1e-10) to fp32 rounding unless the
// denominator is tiny; `minden` tells the caller when train_chunk_exact has to redo the chunk.  MASKED: columns
// >= n_valid are padding of the batch tile.
constexpr int kCw = 16;     // accumulator columns an epilogue warp handles per tcgen05.ld (keeps the live set under 96 registers)
template <bool MASKED>
__device__ __forceinline__ void train_chunk(const uint32_t (&r)[kCw], uint32_t yw, float c1, float c2, float wl_pos,
                                            float wl_neg, float c_pos, float c_neg, int n_valid, float& loss,
                                            float& db, float& minden, uint32_t (&packed)[kCw / 2]) {
    float md = 1.f;
#pragma unroll
    for (int j = 0; j < kCw; j += 2) {
        float dzv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float e = ex2_approx(fmaf(__uint_as_float(r[j + u]), c1, c2));
            const float pr = rcp_approx(1.f + e);
            const float omp = 1.f - pr;
            const bool yb = (yw >> (j + u)) & 1u;
            const float den = (yb ? pr : omp) + kEpsLog;
            float wl = yb ? wl_pos : wl_neg;
            float cz = yb ? c_pos : c_neg;
            if (MASKED) {
                const bool live = (j + u) < n_valid;
                wl = live ? wl : 0.f;
                cz = live ? cz : 0.f;
            }
            md = fminf(md, den);
            loss = fmaf(wl, lg2_approx(den), loss);
            const float dz = (yb ? omp : pr) * cz;
            db += dz;
            dzv[u] = dz;
        }
        packed[j >> 1] = pack_bf16x2(dzv[0], dzv[1]);
    }
    minden = md;
}

// the exact gradient form near saturation (p == 1.0f in fp32 -> p(1-p) == 0 -> dz == 0, as TF computes it)
__device__ __forceinline__ void train_chunk_exact(const uint32_t (&r)[kCw], uint32_t yw, float c1, float c2, float c_pos,
                                               float c_neg, int n_valid, float& db, uint32_t (&packed)[kCw / 2]) {
#pragma unroll
    for (int j = 0; j < kCw; j += 2) {        // static register indices: the arrays must stay in registers
        float dzv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float e = ex2_approx(fmaf(__uint_as_float(r[j + u]), c1, c2));
            const float pr = rcp_approx(1.f + e);
            const float omp = 1.f - pr;
            const bool yb = (yw >> (j + u)) & 1u;
            const float den = (yb ? pr : omp) + kEpsLog;
            const float cz = (j + u) < n_valid ? (yb ? c_pos : c_neg) : 0.f;
            const float fast = (yb ? omp : pr) * cz;
            const float exact = den < 2e-3f ? __fdividef(pr * omp, den) * cz : fast;
            db += exact - fast;
            dzv[u] = exact;
        }
        packed[j >> 1] = pack_bf16x2(dzv[0], dzv[1]);
    }
}

// min of three (sm_100 FMNMX3)
__device__ __forceinline__ float min3(float a, float b, float c) { float y; asm("min.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
// max of three (FMNMX3) and the packed two-lane fp32 add of sm_100 (FADD2: one issue slot for two accumulator columns)
__device__ __forceinline__ float max3(float a, float b, float c) { float y; asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ float2 add2(uint32_t a0, uint32_t a1, float b0, float b1) {
    unsigned long long a, b, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(d));
    return r;
}

// G1 TRAIN epilogue, the common case: 16 full columns treated as y = 0 (targets are ~5e-4 dense; the few y = 1
// cells are patched afterwards from tensor memory).  Per element: t = -z log2e (bias folded, FFMA), e = 2^t (MUFU),
// u = (1 + e) / c_neg (FFMA), dz = 1 / u = p * c_neg (MUFU: the reciprocal IS the gradient), 1 - p = 1 - dz / c_neg
// (FFMA), and the loss term log2(1 - p) is taken once per EIGHT elements on their product (1 - p >= 2e-3 or the
// caller redoes the chunk exactly, so the product cannot underflow): 2.125 MUFU + ~6.5 other instructions per
// element instead of 3 + 19.  Returns sum(log2(1 - p)), sum(dz), min(1 - p).
__device__ __forceinline__ void train_chunk_y0(const uint32_t (&r)[kCw], float c1, float c2, float ic, float& lg_sum,
                                               float& dz_sum, float& minden, uint32_t (&packed)[kCw / 2]) {
    float md = 1.f, ls = 0.f, sd = 0.f;
    const float nic = -ic;
#pragma unroll
    for (int j0 = 0; j0 < kCw; j0 += 8) {
        float prod = 1.f;
#pragma unroll
        for (int j = j0; j < j0 + 8; j += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), c1, c2));
            const float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), c1, c2));
            const float d0 = rcp_approx(fmaf(e0, ic, ic));
            const float d1 = rcp_approx(fmaf(e1, ic, ic));
            const float o0 = fmaf(d0, nic, 1.f);
            const float o1 = fmaf(d1, nic, 1.f);
            md = min3(md, o0, o1);
            prod *= o0;
            prod *= o1;
            sd += d0;
            sd += d1;
            packed[j >> 1] = pack_bf16x2(d0, d1);
        }
        ls += lg2_approx(prod);
    }
    lg_sum = ls; dz_sum = sd; minden = md;
}

// ---- FILTER epilogue helpers (fused decode + top-K) --------------------------------------------------------------
// The list appends of one staging half (n <= 64 entries of (column | item << 8, logit)) in two steps: flush_begin issues
// the returned atomics that reserve the entries' positions in their playlists' lists and does NOT wait for them;
// flush_end, at the next swap of the halves, stores the entries (re-read from the half, which nobody touches in between)
// at those positions.  A flush is a round trip to L2 (1-2 us): nothing waits it out.
// (`o` = the kernel's parameter block: it lives in the constant bank, so the scan keeps no output pointers in registers.)
__device__ __forceinline__ void flush_begin(const uint2* half, int n, const ItemTileDev& o, int& g0, int& g1) {
    const int lane = threadIdx.x & 31;
    const int b0 = blockIdx.y * o.n_cols;          // first playlist of the CTA's batch tile
    __syncwarp();                                  // the half was written by other lanes
    g0 = o.cand_cap; g1 = o.cand_cap;
    if (lane < n) g0 = atomicAdd(o.cand_cnt + b0 + (int)(half[lane].x & 255u), 1);
    if (lane + 32 < n) g1 = atomicAdd(o.cand_cnt + b0 + (int)(half[lane + 32].x & 255u), 1);
}
__device__ __noinline__ void flush_end(const uint2* half, int n, const ItemTileDev& o, int g0, int g1) {
    const int lane = threadIdx.x & 31;
    const int b0 = blockIdx.y * o.n_cols;
    if (lane < n && g0 < o.cand_cap) {
        const uint2 e = half[lane];
        const size_t at = (size_t)(b0 + (int)(e.x & 255u)) * o.cand_cap + g0;
        o.cand_val[at] = __uint_as_float(e.y);
        o.cand_idx[at] = o.item0 + (int)(e.x >> 8);
    }
    if (lane + 32 < n && g1 < o.cand_cap) {
        const uint2 e = half[lane + 32];
        const size_t at = (size_t)(b0 + (int)(e.x & 255u)) * o.cand_cap + g1;
        o.cand_val[at] = __uint_as_float(e.y);
        o.cand_idx[at] = o.item0 + (int)(e.x >> 8);
    }
    __syncwarp();
}

// Queued accumulator rows -> staged candidates, the whole warp on two records at a time (lane = cell k of record
// lane >> 4): the scan's own test per cell, (acc_k - thr'_k) + bias >= -tol with the same rounding (thr' = the loosened
// threshold in shared memory: whatever passes z_k >= thr_k passes this, and the few extra candidates just below a
// threshold are harmless -- the select is exact on whatever the lists hold), hits appended to the staging half at
// positions from the ballot.  Stops when the half cannot take a pair's hits: returns entries staged | records done << 16.
__device__ __noinline__ uint32_t process_queue(const float* qb, int qi, int qn, const float* nthr, uint2* seg, int wn) {
    const int lane = threadIdx.x & 31, k = lane & 15, sub = lane >> 4;
    const unsigned ltmask = (1u << lane) - 1u;
#pragma unroll 1
    for (; qi < qn; qi += 2) {
        const int rec = qi + sub;
        bool pass = false;
        uint32_t tag = 0;
        float z = 0.f;
        if (rec < qn) {
            const float* R = qb + rec * kQWords;
            const float a = R[k], bz = R[16];
            tag = __float_as_uint(R[17]);
            const float nt = nthr[(tag & 255u) + k];
            pass = __fadd_rn(__fadd_rn(a, nt), bz) >= -4e-7f * fabsf(bz);
            z = a + bz;                                          // the logit, as the dense path forms it
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        const int n = __popc(bal);
        if (wn + n > kWHalf) break;                              // warp-uniform: the caller swaps halves and comes back
        if (pass) seg[wn + __popc(bal & ltmask)] = make_uint2(tag + k, __float_as_uint(z));
        wn += n;
    }
    return (uint32_t)wn | ((uint32_t)(qi < qn ? qi : qn) << 16);
}

// blockIdx.y = batch tile: TRAIN decodes this rank's item rows against every rank's rows of the global batch
// (tile bt = rank bt's playlists), PREDICT tiles an inference batch of more than 256 rows.
template <int MODE>
__global__ void __launch_bounds__(kItemThreads, 1)
k_itemtile(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ItemTileDev p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
    // Shared memory: the resident operand (h_d of the batch tile: kchunks x [n_cols rows x 128 B]), then the ring of W
    // chunks, then barriers at a fixed offset.  TRAIN: 4 stages behind a 128 KB resident area (kStagesTrain).  Inference:
    // the ring takes ALL the space the resident operand leaves -- the decode is paced by bytes in flight (TMA latency x
    // rate), and a 128-row batch tile (64 KB resident) leaves 10 stages where a 256-row one leaves 6; FILTER gives the last
    // stage, 16 KB, to the candidate staging buffer.
    const int bchunk = MODE == MODE_TRAIN ? kBChunkBytes : p.n_cols * 128;
    const int RING = MODE == MODE_TRAIN ? kStagesTrain : p.ring;
    const int NST = MODE == MODE_FILTER ? RING - kFilterStgStages : RING;
    constexpr int kRingBytes = (MODE == MODE_TRAIN ? kStagesTrain : kStages) * kABytes;
    uint8_t* sB = smem;
    uint8_t* sA = smem + (MODE == MODE_TRAIN ? kSmemB : ((p.kchunks * bchunk + 1023) & ~1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemB + kRingBytes);
    uint64_t* full = bars;                 // [kMaxRing]
    uint64_t* empty = bars + kMaxRing;     // [kMaxRing]
    uint64_t* bfull = bars + 2 * kMaxRing; // [1]
    uint64_t* tfull = bfull + 1;           // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* loss_smem = reinterpret_cast<float*>(tmem_slot + 1);  // [kEpiWarps]
    float* thr_smem = reinterpret_cast<float*>(smem + kSmemB + kRingBytes + kSmemBars);   // [n_cols] (FILTER)
    uint8_t* stg = sA + NST * kABytes;                                            // FILTER: [kEpiWarps][kWarpStgBytes]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int bt = blockIdx.y;
    // Pair mode (cluster dims (1, 2, 1)): the two CTAs hold different batch tiles and walk the same item tiles in lock
    // step.  Each W chunk is read from L2 ONCE -- by the CTA whose rank equals the chunk's parity -- and multicast into both
    // CTAs' rings; every MMA commit releases the stage in both.  The decode is L2 -> SM bandwidth bound otherwise (16 batch
    // tiles x 1 GB of W through L2 per pass: ncu 2.6 TB/s, tensor pipe 33 % busy).
    const bool pair = MODE != MODE_TRAIN && p.pair != 0;
    const int crank = bt & 1;
    if (MODE == MODE_TRAIN && p.trace != nullptr && threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMin(p.trace, now);
        atomicMax(p.trace + 1, now);
    }
    if (MODE == MODE_FILTER) {
        // the scan works on max_k(acc_k - thr_k) + bias: shared memory holds MINUS the threshold, loosened by more than the
        // rounding of either form of the comparison -- (acc - thr) + bias against fl(acc + bias) >= thr: half an ulp of
        // |thr| and of |bias| (the scan adds 4e-7 |bias|), ~6e-8 relative each -- so that nothing with z >= thr is dropped;
        // what passes in addition (a few items within ~1e-6 below a threshold) only makes a list longer
        for (int i = threadIdx.x; i < p.n_cols; i += blockDim.x) {
            const float t = p.thr[bt * p.n_cols + i];
            thr_smem[i] = fabsf(t) <= 3.0e38f ? -(t - (2e-6f + 4e-7f * fabsf(t))) : -t;
        }
    }

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], pair ? 2 : 1);      // a stage is free once BOTH CTAs of the pair have consumed it
        }
        mbar_init(bfull, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();                    // the peer's barriers exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const uint64_t pol_stream = policy_evict_first();
            const uint64_t pol_keep = policy_evict_last();
            mbar_expect_tx(bfull, static_cast<uint32_t>(p.kchunks * p.n_cols * 128));
            for (int kc = 0; kc < p.kchunks; ++kc)
                tma_load_2d_hint(sB + kc * bchunk, &tmB, bfull, kc * 64, bt * p.n_cols, pol_keep);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_expect_tx(&full[stage], kABytes);
                    // several batch tiles re-read the same item tile: keep it in L2 then, stream it otherwise
                    if (!pair)
                        tma_load_2d_hint(sA + stage * kABytes, &tmA, &full[stage], kc * 64, tile * kTileItems,
                                         gridDim.y > 1 ? pol_keep : pol_stream);
                    else if ((kc & 1) == crank)
                        tma_load_2d_mcast(sA + stage * kABytes, &tmA, &full[stage], kc * 64, tile * kTileItems, 0x3, pol_keep);
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The WHOLE warp walks the loop (converged) and one elected lane issues: inside an `if (lane == 0)` region the
        // compiler wraps every tcgen05 instruction in an elect / branch loop and rebuilds both descriptors from their byte
        // addresses, ~110 instructions per K chunk on one dependent chain -- the issue loop then took LONGER than the MMAs
        // it feeds (ncu: half of this warp's samples).  Descriptors advance by adding to their low word (start address
        // >> 4: no carry out of the 14-bit field below 256 KB).
        {
            const uint32_t idesc = umma_idesc_bf16(kTileItems, static_cast<uint32_t>(p.n_cols), 0, 0);
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024 B, version 1, SWIZZLE_128B
            const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);      // LBO field = 1 (unused for K-major)
            const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_step = static_cast<uint32_t>(bchunk) >> 4;
            mbar_wait(bfull, 0);
            tc_fence_after();
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(stage) * (kABytes >> 4);
                    const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(kc) * b_step;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_bf16(d_tmem, umma_desc_pack(a_lo + 2 * ks, kDescHi), umma_desc_pack(b_lo + 2 * ks, kDescHi), idesc,
                                      (kc | ks) != 0 ? 1u : 0u);
                        if (pair) umma_commit_mcast(&empty[stage], 0x3);
                        else umma_commit(&empty[stage]);
                        if (kc == p.kchunks - 1) umma_commit(&tfull[acc]);      // same thread as the MMAs it tracks
                    }
                    __syncwarp();
                    if (++stage == NST) { stage = 0; phase ^= 1u; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3;                    // TMEM lane quadrant this warp may read
        const int part = (warp - 2) >> 2;          // chunks part, part + 4, ... of the accumulator columns are this warp's
        const int row_in_tile = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        float loss_acc = 0.f;
        const int nchunks = p.n_cols >> 5;
        const int yword0 = bt * (p.n_cols >> 5);   // this batch tile's words inside a row of the target bitmask
        // e^-z = 2^(acc * c1 + c2) with the bias folded in; -w ln(den) = wl * lg2(den); dz = ratio * (c_pos | c_neg)
        const float c1 = -1.4426950408889634f;
        const float wl_pos = -0.6931471805599453f, wl_neg = kNegWeight * wl_pos;
        const float c_pos = -p.inv_batch, c_neg = kNegWeight * p.inv_batch;
        const float ic = 1.f / c_neg;
        // FILTER: this warp's staging segment, entries staged (warp-uniform), where they go
        // FILTER: the warp's queue of accumulator rows (qn records) and its staging halves: `seg` is being filled (wn
        // entries), the other one (address ^ 512) holds the pn entries whose list appends are in flight (positions g0 / g1)
        float* const qb = reinterpret_cast<float*>(stg + (warp - 2) * kWarpStgBytes);
        uint2* seg = reinterpret_cast<uint2*>(stg + (warp - 2) * kWarpStgBytes + kStgOff);
        auto other_half = [](uint2* h) { return reinterpret_cast<uint2*>(reinterpret_cast<uintptr_t>(h) ^ (uintptr_t)(kWHalf * 8)); };
        int wn = 0, pn = 0, g0 = 0, g1 = 0, qn = 0;
        // every queued row through the cell test into the staging half; a full half is handed to the L2 (appends begun, not
        // awaited) and the other one -- its appends of the previous swap long done -- takes over
        auto drain = [&]() {
            __syncwarp();                                   // the queue was written by other lanes
            int qi = 0;
            while (qi < qn) {
                const uint32_t rv = process_queue(qb, qi, qn, thr_smem, seg, wn);
                wn = (int)(rv & 0xffffu); qi = (int)(rv >> 16);
                if (qi < qn) {
                    if (pn > 0) flush_end(other_half(seg), pn, p, g0, g1);
                    flush_begin(seg, wn, p, g0, g1);
                    pn = wn; wn = 0;
                    seg = other_half(seg);
                }
            }
            qn = 0;
            __syncwarp();                                   // every lane has read its records before any lane queues new ones
        };
        // per-tile scalars of this lane's item row (bias, target words of the warp's two chunks) are fetched ONE TILE
        // AHEAD: they come from HBM (~1 us) and would otherwise stall the warp at the top of every tile
        float bz_nx = 0.f;
        uint32_t ya_nx = 0, yb_nx = 0;
        auto prefetch_row = [&](int tile) {
            const int row = tile * kTileItems + row_in_tile;
            const int item = MODE == MODE_TRAIN ? item_global(row, p.world, p.rank) : row;
            bz_nx = 0.f; ya_nx = 0; yb_nx = 0;
            if (tile < p.tiles && row < p.n_rows && item < p.n_global) {
                bz_nx = __ldg(p.bias + item);
                if (MODE == MODE_TRAIN) {
                    const uint32_t* yrow = p.ybits + (size_t)row * p.ywords + yword0;
                    if (part < nchunks) ya_nx = __ldg(yrow + part);
                    if (part + 4 < nchunks) yb_nx = __ldg(yrow + part + 4);
                }
            }
        };
        prefetch_row(blockIdx.x);
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int row = tile * kTileItems + row_in_tile;                                   // row of the streamed operand
            const int item = MODE == MODE_TRAIN ? item_global(row, p.world, p.rank) : row;     // catalogue id
            const bool item_ok = row < p.n_rows && item < p.n_global;
            const float bz = bz_nx;
            const uint32_t yw_a = ya_nx, yw_b = yb_nx;
            prefetch_row(tile + gridDim.x);
            float db = 0.f, lt = 0.f;
            const float c2 = bz * c1;
            const float bzf = item_ok ? bz : __int_as_float(0x7fc00000);   // FILTER: rows past the range compare false (NaN)
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + lane_addr + static_cast<uint32_t>(acc * 256);
            if (MODE == MODE_FILTER) {
            // fused decode + top-K, filter stage: keep the logits >= the playlist's threshold (a lower bound of
            // its (K + #seeds)-th largest logit, so no member of the top-K can be lost); ~0.03 % of the cells pass in
            // the full-range pass, ~4 % in the middle one -- and they cluster on ROWS: a popular item passes for
            // most playlists of the tile, so one lane quadrant's four warps find hits in every half chunk of a
            // tile while the other twelve find none, on nearly every tile.  Every tile needs all 16 warps before
            // the tensor core may reuse the accumulator, so the scan does the minimum while it holds it: a PRE-TEST
            // on the lane's 16 cells at once, m = max_k (acc_k - thr'_k) (one FADD2 + one FMNMX3 per two cells,
            // thr' = the threshold loosened beyond the rounding of both forms), (m + bias) >= -tol, one vote per
            // half chunk; a lane that passes copies its 16 accumulator values to the warp's queue (5 stores).  The
            // per-cell test, the staging and the list appends happen AFTER the accumulator is released (`drain`),
            // where a hot warp's extra work is absorbed by the slack of the two-accumulator pipeline instead of
            // stalling the CTA (ncu + timing experiments: 4.4 ms with the hit path inside the scan, 2.4 without any).
            // The accumulator columns of the NEXT half chunk are requested before the current one is examined (two register
            // sets): the tensor-memory read latency (~100 cycles with 16 warps reading) is off the warp's serial path.
            const int nh = (p.tune & 2) ? 0 : (part < nchunks ? 2 * ((nchunks - part + 3) >> 2) : 0);   // this warp's half chunks
            auto col_of = [&](int j) { return (part + 4 * (j >> 1)) * 32 + (j & 1) * kCw; };
            auto examine = [&](const uint32_t (&r)[kCw], int col0) {
                const float4* nth4 = reinterpret_cast<const float4*>(thr_smem + col0);
                float mxa = -CUDART_INF_F, mxb = -CUDART_INF_F;
#pragma unroll
                for (int j4 = 0; j4 < kCw / 4; ++j4) {
                    const float4 nt = nth4[j4];                      // same address for every lane: broadcast
                    const float2 d0 = add2(r[4 * j4 + 0], r[4 * j4 + 1], nt.x, nt.y);
                    const float2 d1 = add2(r[4 * j4 + 2], r[4 * j4 + 3], nt.z, nt.w);
                    mxa = max3(mxa, d0.x, d0.y);
                    mxb = max3(mxb, d1.x, d1.y);
                }
                bool mine = fmaxf(mxa, mxb) + bzf >= -4e-7f * fabsf(bzf);   // tolerance: rounding of (acc - thr) ~ ulp(bias)
                unsigned rem = __ballot_sync(0xffffffffu, mine);
                if (p.tune & 4) rem = 0;
#pragma unroll 1
                while (rem != 0) {                                   // (one round unless the queue fills up)
                    if (qn == kQRec) drain();
                    const int room = kQRec - qn;
                    const int rank = __popc(rem & ((1u << lane) - 1u));
                    if (mine && rank < room) {
                        float* R = qb + (qn + rank) * kQWords;
#pragma unroll
                        for (int j4 = 0; j4 < kCw / 4; ++j4)
                            *reinterpret_cast<uint4*>(R + 4 * j4) = make_uint4(r[4 * j4], r[4 * j4 + 1], r[4 * j4 + 2], r[4 * j4 + 3]);
                        *reinterpret_cast<float2*>(R + 16) = make_float2(bzf, __uint_as_float((uint32_t)col0 | ((uint32_t)item << 8)));
                        mine = false;
                    }
                    const int cnt = __popc(rem);
                    qn += cnt < room ? cnt : room;
                    rem = __ballot_sync(0xffffffffu, mine);
                }
            };
            uint32_t ra[kCw], rb[kCw];
            if (nh > 0) tmem_ld16(t_addr + col_of(0), ra);
#pragma unroll 1
            for (int j = 0; j < nh; j += 2) {
                tmem_ld_wait();
                if (j + 1 < nh) tmem_ld16(t_addr + col_of(j + 1), rb);
                examine(ra, col_of(j));
                if (j + 1 < nh) {
                    tmem_ld_wait();
                    if (j + 2 < nh) tmem_ld16(t_addr + col_of(j + 2), ra);
                    examine(rb, col_of(j + 1));
                }
            }
            } else
#pragma unroll 1
            for (int c = part; c < nchunks; c += 4) {
                if (MODE == MODE_TRAIN) {
                    const uint32_t w32 = c == part ? yw_a : yw_b;
#pragma unroll 1
                    for (int hh = 0; hh < 2; ++hh) {
                        const int col0 = c * 32 + hh * kCw;            // first batch column of this half chunk
                        uint32_t r[kCw];
                        tmem_ld16(t_addr + col0, r);
                        tmem_ld_wait();
                        const uint32_t w = (w32 >> (hh * kCw)) & 0xffffu;
                        uint32_t packed[kCw / 2];
                        uint32_t pend = 0;             // y = 1 cells of this lane still to be patched
                        bool general = col0 + kCw > p.batch;           // partial half chunk (batch padding): warp-uniform
                        if (!general) {
                            float ls, sd, md;
                            train_chunk_y0(r, c1, c2, ic, ls, sd, md, packed);
                            if (md < 2e-3f) general = true;            // some 1 - p is within reach of the 1e-10 inside the logs
                            else { lt = fmaf(wl_neg, ls, lt); db += sd; pend = w; }
                        }
                        if (general) {
                            float minden;
                            if (col0 + kCw <= p.batch) train_chunk<false>(r, w, c1, c2, wl_pos, wl_neg, c_pos, c_neg, 0, lt, db, minden, packed);
                            else train_chunk<true>(r, w, c1, c2, wl_pos, wl_neg, c_pos, c_neg, p.batch - col0, lt, db, minden, packed);
                            if (minden < 2e-3f)     // exact gradient form near saturation
                                train_chunk_exact(r, w, c1, c2, c_pos, c_neg, p.batch - col0, db, packed);
                        }
                        __nv_bfloat16* dst = p.dzT + (size_t)row * p.ld_dz + bt * p.n_cols + col0;     // 32 B: one full sector
                        if (item_ok) st_global_v8(dst, packed);
                        // patch the y = 1 cells: re-read the accumulator column (warp-uniform address), redo the cell with
                        // its target, replace its bf16 gradient (same thread, later store wins) and correct the sums
                        uint32_t um = __reduce_or_sync(0xffffffffu, item_ok ? pend : 0u);
                        while (um != 0) {
                            const int k = __ffs(um) - 1;
                            um &= um - 1;
                            const float zacc = __uint_as_float(tmem_ld1(t_addr + col0 + k));
                            tmem_ld_wait();
                            if (item_ok && ((pend >> k) & 1u)) {
                                const float e = ex2_approx(fmaf(zacc, c1, c2));
                                const float pr = rcp_approx(1.f + e);
                                const float omp = 1.f - pr;
                                const float den = pr + kEpsLog;
                                lt = fmaf(wl_pos, lg2_approx(den), lt);
                                lt = fmaf(-wl_neg, lg2_approx(omp), lt);
                                const float dz1 = (den < 2e-3f ? __fdividef(pr * omp, den) : omp) * c_pos;
                                db += dz1 - pr * c_neg;
                                dst[k] = __float2bfloat16_rn(dz1);
                            }
                        }
                    }
                } else {
                    uint32_t r[32];
                    tmem_ld32(t_addr + c * 32, r);
                    tmem_ld_wait();
                    const bool col_ok = item < p.n_out;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int b = bt * p.n_cols + c * 32 + j;
                        const float z = __uint_as_float(r[j]) + bz;
                        float pr = p.raw_logits ? z : __fdividef(1.f, 1.f + __expf(-z));
                        if (b < p.batch && col_ok && item_ok) {
                            const size_t o = (size_t)b * p.ld_out + item;
                            if (p.mix_wp != nullptr) {
                                const float ts = p.title_score ? __ldg(p.title_score + o) : 0.f;
                                pr = ts * __ldg(p.mix_wt + b) + pr * __ldg(p.mix_wp + b);
                            }
                            p.out[o] = pr;   // lanes = consecutive items -> 128 B per warp store
                        }
                    }
                }
            }
            if (MODE == MODE_TRAIN) {
                if (item_ok) {
                    // this warp's share of the row's bias gradient (its column chunks of this batch tile): stored, not
                    // atomically added -- k_sum_db_parts adds the 4 x batch-tile shares in a fixed order (deterministic)
                    p.db_dec[(size_t)(bt * 4 + part) * p.n_rows + row] = db;
                    loss_acc += lt;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
            if (MODE == MODE_FILTER) {
                if (qn > 0) drain();
            }
        }
        if (MODE == MODE_FILTER) {
            if (pn > 0) flush_end(other_half(seg), pn, p, g0, g1);
            if (wn > 0) {
                flush_begin(seg, wn, p, g0, g1);
                flush_end(seg, wn, p, g0, g1);
            }
        }
        if (MODE == MODE_TRAIN) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
            if (lane == 0) loss_smem[warp - 2] = loss_acc;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();      // no CTA leaves while its peer may still signal its barriers
    tc_fence_after();
    if (MODE == MODE_TRAIN) {
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < kEpiWarps; ++i) t += loss_smem[i];
            p.loss_partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// db_dec[row] = sum of the epilogue warps' shares (4 column parts x batch tiles), fixed order; rows of tiles no CTA
// processed (beyond the catalogue) are zero
__global__ void k_sum_db_parts(const float* __restrict__ parts, int n_parts, int n_rows, int rows_done, float* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    float t = 0.f;
    if (r < rows_done)
        for (int k = 0; k < n_parts; ++k) t += parts[(size_t)k * n_rows + r];
    out[r] = t;
}

static int sm_count() {
    static int n = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}

int decode_grid(int N, int n_batch_tiles) {
    const int tiles = (N + kTileItems - 1) / kTileItems;
    int gx = sm_count() / (n_batch_tiles > 0 ? n_batch_tiles : 1);
    if (gx < 1) gx = 1;
    return tiles < gx ? tiles : gx;
}

static int g_itemtile_pair = 0;       // dae_model_set_debug bit 15 sets it (multicast batch-tile pairs: measured, no gain)
void set_itemtile_pair(int on) { g_itemtile_pair = on; }
static int g_itemtile_tune = 0;       // dae_model_set_debug bits 16.. (A/B switches of the FILTER epilogue)
void set_itemtile_tune(int bits) { g_itemtile_tune = bits; }

template <int MODE>
static void launch_itemtile(const CUtensorMap& tmA, const CUtensorMap& tmB, ItemTileDev p, dim3 grid, cudaStream_t st) {
    const size_t smem = MODE == MODE_TRAIN ? kSmemItemTileTrain : kSmemItemTile;
    // an even number of batch tiles (PREDICT / FILTER over a large batch): clusters of two batch tiles share the W stream
    p.pair = (MODE != MODE_TRAIN && g_itemtile_pair && grid.y >= 2 && grid.y % 2 == 0) ? 1 : 0;
    p.ring = (kSmemB + kSmemA - ((p.kchunks * p.n_cols * 128 + 1023) & ~1023)) / kABytes;
    p.tune = g_itemtile_tune;
    if (p.ring > kMaxRing) p.ring = kMaxRing;
    if (!p.pair) {
        k_itemtile<MODE><<<grid, kItemThreads, smem, st>>>(tmA, tmB, p);
        return;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(kItemThreads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 2; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_itemtile<MODE>, tmA, tmB, p);
}

void launch_decode_train(const DecodeArgs& a, cudaStream_t st) {
    const int nbt = a.n_batch_tiles > 0 ? a.n_batch_tiles : 1;
    ItemTileDev p{};
    p.world = a.pt.world > 0 ? a.pt.world : 1;
    p.rank = a.pt.rank;
    p.n_rows = a.n_local;
    p.n_global = a.N;
    // local tiles that hold at least one valid catalogue row: global tile = local * world + rank
    const int tiles_total = (a.N + kTileItems - 1) / kTileItems;
    p.tiles = tiles_total > p.rank ? (tiles_total - p.rank + p.world - 1) / p.world : 0;
    p.kchunks = a.H / 64;
    p.n_cols = a.bpad;
    p.batch = a.batch;
    p.bias = a.bias;
    p.ybits = a.ybits; p.ywords = a.ywords;
    p.dzT = a.dzT; p.ld_dz = nbt * a.bpad;
    p.db_dec = a.db_parts; p.loss_partial = a.loss_partial; p.inv_batch = a.inv_batch;
    p.trace = a.trace;
    const int gx = decode_grid(p.tiles * kTileItems, nbt);
    if (p.tiles == 0) {                                             // a rank may own no tile of a tiny catalogue
        cudaMemsetAsync(a.loss_partial, 0, sizeof(float) * 2 * 148, st);
        cudaMemsetAsync(a.db_dec, 0, sizeof(float) * a.n_local, st);
        return;
    }
    cudaMemsetAsync(a.loss_partial, 0, sizeof(float) * 2 * 148, st);
    const CUtensorMap tmA = make_map_bf16(a.W, a.H, a.n_local, kTileItems);
    const CUtensorMap tmB = make_map_bf16(a.h_d, a.H, (uint64_t)a.bpad * nbt, a.bpad);
    launch_itemtile<MODE_TRAIN>(tmA, tmB, p, dim3(gx, nbt, 1), st);
    k_sum_db_parts<<<(a.n_local + 255) / 256, 256, 0, st>>>(a.db_parts, 4 * nbt, a.n_local, p.tiles * kTileItems, a.db_dec);
}

void launch_decode_predict(const DecodeArgs& a, cudaStream_t st) {
    const int nbt = a.n_batch_tiles > 0 ? a.n_batch_tiles : 1;
    ItemTileDev p{};
    p.world = 1; p.rank = 0;
    p.n_rows = a.n_out < a.N - a.item0 ? a.n_out : a.N - a.item0;   // only the columns that are kept (tracks) are scored
    p.n_global = a.N - a.item0;
    p.tiles = (p.n_rows + kTileItems - 1) / kTileItems;
    p.kchunks = a.H / 64;
    p.n_cols = a.bpad;
    p.batch = a.batch;
    p.bias = a.bias + a.item0;                      // rows [item0, item0 + n_out) of the catalogue -> output columns [0, n_out)
    p.out = a.out; p.ld_out = a.ld_out; p.n_out = a.n_out;
    p.mix_wp = a.mix_wp; p.mix_wt = a.mix_wt; p.title_score = a.title_score;
    p.raw_logits = a.raw_logits;
    const CUtensorMap tmA = make_map_bf16(a.W + (size_t)a.item0 * a.H, a.H, a.N - a.item0, kTileItems);
    const CUtensorMap tmB = make_map_bf16(a.h_d, a.H, (uint64_t)a.bpad * nbt, a.bpad);
    launch_itemtile<MODE_PREDICT>(tmA, tmB, p, dim3(decode_grid(p.n_rows, nbt), nbt, 1), st);
}

// G1f: decode of the catalogue rows [a.item0, a.item0 + a.n_out) against every batch tile, logits above the playlist's
// threshold appended to its candidate list (fused decode + top-K: the [B, T] score matrix never exists)
void launch_decode_filter(const DecodeArgs& a, cudaStream_t st) {
    const int nbt = a.n_batch_tiles > 0 ? a.n_batch_tiles : 1;
    ItemTileDev p{};
    p.world = 1; p.rank = 0;
    p.n_rows = a.n_out;
    p.n_global = a.n_out;
    p.tiles = (p.n_rows + kTileItems - 1) / kTileItems;
    p.kchunks = a.H / 64;
    p.n_cols = a.bpad;
    p.batch = a.batch;
    p.bias = a.bias + a.item0;
    p.thr = a.thr; p.cand_val = a.cand_val; p.cand_idx = a.cand_idx; p.cand_cnt = a.cand_cnt; p.cand_cap = a.cand_cap;
    p.item0 = a.item0;
    if (p.tiles == 0) return;
    const CUtensorMap tmA = make_map_bf16(a.W + (size_t)a.item0 * a.H, a.H, a.n_out, kTileItems);
    const CUtensorMap tmB = make_map_bf16(a.h_d, a.H, (uint64_t)a.bpad * nbt, a.bpad);
    launch_itemtile<MODE_FILTER>(tmA, tmB, p, dim3(decode_grid(p.n_rows, nbt), nbt, 1), st);
}

// ------------------------------------------------------------------------------------------
// G2: dW_dec^T tile = h_d^T . dz^T  with the dense TF1 Adam update fused into the epilogue
// ------------------------------------------------------------------------------------------
// Orientation: the accumulator holds dW^T -- TMEM lane = hidden unit h, TMEM column = item -- so that
// for one item the 32 lanes of an epilogue warp touch 32 CONSECUTIVE floats of that item's row of
// W / m / v (one full 128-byte line per warp access; the item-major layout is the reference's
// [n_input, n_hidden], DAEs.py:54).  A = h_d^T [H, K] (K-major; resident in smem, or streamed when
// K = ranks x batch tile > 256), B = dz tile [128 items, K] (K-major, streamed).  M = 128 per UMMA:
// H = 256 takes two M-halves (TMEM columns [0,128) and [128,256) of the tile's buffer), H = 64 runs
// with zero rows (TMA out-of-bounds fill).  192 + 16 epilogue warps: the epilogue is pure HBM
// streaming (26 B / parameter) and needs bytes in flight, not registers.
constexpr int kDwEpiWarps = 16;
constexpr int kDwThreads = 64 + 32 * kDwEpiWarps;

struct DwDev {
    int tiles;        // local item tiles that hold at least one valid catalogue row
    int kchunks;      // K / 64
    int H;            // hidden units (valid TMEM lanes over the M-halves)
    int mhalves;      // ceil(H / 128)
    int stream_h;     // h_d^T chunk rides in the ring with the dz chunk (K > 256)
    int n_global;     // catalogue size
    float* g;         // raw dW_dec [local rows, H]
    PeerTable pt;
    int ld, col0;            // row stride and first column of the row-major outputs
};

__global__ void __launch_bounds__(kDwThreads, 1)
k_dw(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmH,
          const __grid_constant__ DwDev p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
    uint8_t* sH = smem;                    // resident h_d^T: kchunks x [256 rows x 128 B]
    uint8_t* sA = smem + kSmemB;           // ring of dz chunks [128 items x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemB + kSmemA);
    uint64_t* full = bars;                 // [kStages]
    uint64_t* empty = bars + kStages;      // [kStages]
    uint64_t* hfull = bars + 2 * kStages;  // [1]
    uint64_t* tfull = hfull + 1;           // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool sh = p.stream_h != 0;
    const int nstages = sh ? kStagesStreamB : kStages;
    const uint32_t stage_bytes = sh ? (kABytes + kBChunkBytes) : kABytes;
    const int hbox_rows = p.mhalves * 128;                       // rows per h_d^T box (rows >= H are zero-filled)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDz);
        tma_prefetch_desc(&tmH);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(hfull, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], kDwEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const uint64_t pol_stream = policy_evict_first();
            const uint64_t pol_keep = policy_evict_last();
            if (!sh) {
                mbar_expect_tx(hfull, static_cast<uint32_t>(p.kchunks * hbox_rows * 128));
                for (int kc = 0; kc < p.kchunks; ++kc)
                    tma_load_2d_hint(sH + kc * kBChunkBytes, &tmH, hfull, kc * 64, 0, pol_keep);
            }
            uint8_t* ring = sh ? smem : sA;
            const uint32_t tx = static_cast<uint32_t>(kABytes + (sh ? hbox_rows * 128 : 0));
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_expect_tx(&full[stage], tx);
                    tma_load_2d_hint(ring + stage * stage_bytes, &tmDz, &full[stage], kc * 64, tile * kTileItems,
                                     pol_stream);
                    if (sh) tma_load_2d_hint(ring + stage * stage_bytes + kABytes, &tmH, &full[stage], kc * 64, 0, pol_keep);
                    if (++stage == nstages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, kTileItems, 0, 0);
            if (!sh) {
                mbar_wait(hfull, 0);
                tc_fence_after();
            }
            uint8_t* ring = sh ? smem : sA;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t dz_addr = smem_u32(ring + stage * stage_bytes);
                    const uint32_t h_addr = sh ? dz_addr + kABytes : smem_u32(sH + kc * kBChunkBytes);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t bd = umma_smem_desc(dz_addr + ks * 32, 16, 1024);
                        for (int mh = 0; mh < p.mhalves; ++mh) {
                            const uint64_t ad = umma_smem_desc(h_addr + mh * (128 * 128) + ks * 32, 16, 1024);
                            umma_bf16(d_tmem + static_cast<uint32_t>(mh * kTileItems), ad, bd, idesc, (kc | ks) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == nstages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3;                        // TMEM lane quadrant this warp may read
        const int part = (warp - 2) >> 2;              // 0..3: which quarter of the tile's (M-half, item) columns
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int cols_per = p.mhalves * kTileItems / 4;   // 64 (H > 128) or 32
        const int col_lo = part * cols_per;
        const int mh = col_lo / kTileItems;
        const int it_lo = col_lo % kTileItems;
        const int h = mh * 128 + q * 32 + lane;
        const bool h_ok = h < p.H;
        const int world = p.pt.world;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + lane_addr + static_cast<uint32_t>(acc * 256 + col_lo);
#pragma unroll 1
            for (int cc = 0; cc < cols_per; cc += 16) {
                uint32_t r[16];
                __syncwarp();                                                        // tcgen05.ld is warp-collective
                tmem_ld16(t_addr + cc, r);
                tmem_ld_wait();
                const int item0 = tile * kTileItems + it_lo + cc;                    // local row of column cc
                const int gitem0 = item_global(item0, world, p.pt.rank);             // 16 consecutive catalogue ids
                const int nvalid = h_ok ? min(16, p.n_global - gitem0) : 0;
                const size_t off0 = (size_t)item0 * p.ld + p.col0 + h;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < nvalid) p.g[off0 + (size_t)j * p.ld] = __uint_as_float(r[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// G2 + K5: dW_dec^T tile in tensor memory, dense TF1 Adam applied to it from there (the
// gradient never exists in HBM: 2 (dz) + 24 (w, m, v read + write) + 2 (bf16 operand) = 28 B / parameter
// instead of 2 + 4 (dW write) + 30 for the two-kernel path)
// ------------------------------------------------------------------------------------------
// Same W^T orientation as k_dw (TMEM lane = hidden unit, column = item).  The optimizer state never passes
// through registers as global loads: one I/O thread streams 8-item row groups of w / m / v (8 x H fp32 = 8 KB
// each, contiguous in the item-major layout) into a 6- (or 4-) stage shared-memory ring with 1-D bulk copies
// (cp.async.bulk, mbarrier complete_tx), the 16 epilogue warps (two groups of 8, alternating row groups)
// read their accumulator columns with tcgen05.ld, update the stage IN PLACE (conflict-free: lanes =
// consecutive h), and the I/O thread writes the stage back with bulk stores (bulk async-groups; a stage is
// re-loaded once its store has finished reading it).  Bytes in flight come from the copy engine, not from
// warps stalled on loads, so the epilogue keeps HBM busy with 16 warps.  The bf16 operand copy of the
// updated rows goes out as 64-byte warp stores.  MMA operands: both the dz chunk and the h_d^T chunk ride
// a 2-stage ring (h_d^T is 128 KB and L2-resident); the MMA of tile t+1 overlaps the epilogue of tile t
// through the two TMEM accumulators, and is ~10x shorter than it.
constexpr int kFuItems = 8;                                    // rows of w / m / v per staging stage
constexpr int kFuSub = kTileItems / kFuItems;                  // 16 row groups per tile
// Ring depths: K <= 512: 2 MMA operand stages + 5 staging stages; K > 512 (4+ ranks, longer contraction): 3 + 3.
// Row group i is handled by epilogue group i % 2 and uses stage i % NS.  With NS odd the two groups ALTERNATE on a stage,
// so every stage has one "rows have landed" barrier PER GROUP (ld_full[stage][group]): the warps of a group then wait for
// consecutive phases of their own barrier.  (Round-1 bug: with one barrier per stage a warp could be two phases ahead of
// it; a parity wait cannot tell "phase k completed" from "phase k+2 completed", the warp passed early, updated a stage
// whose rows had not landed and the pipeline derailed -- a rare mbarrier time-out on one GPU, within a few steps with 4+
// ranks or other kernels sharing the SMs.  Even depths (each stage owned by one group) are also correct but measured
// 9 % slower: 0.381 vs 0.349 ms.)
constexpr int kFuStages = 5;                                   // max staging stages (8 rows x w, m, v each)
constexpr int kFuMmaStages = 3;                                // max MMA operand stages (16 KB dz chunk + 32 KB h_d^T chunk)
constexpr int kFuMmaStageBytes = kABytes + kBChunkBytes;       // 16 KB dz chunk + 32 KB h_d^T chunk
constexpr int kFuStageBytes = 3 * kFuItems * 256 * 4;          // 24 KB at H = 256
constexpr int kFuEpiWarps = 16;
constexpr int kFuThreads = 128 + 32 * kFuEpiWarps;             // producer, MMA, I/O, (idle), 16 epilogue warps
constexpr int kFuDataBytes = 2 * kFuMmaStageBytes + 5 * kFuStageBytes;              // == 3 * 48 KB + 3 * 24 KB
static_assert(kFuDataBytes == 3 * kFuMmaStageBytes + 3 * kFuStageBytes, "both layouts fill the same bytes");
constexpr int kSmemFused = kFuDataBytes + 256 + 1024;       // data, barriers (25 x 8 B), alignment slack

struct FusedDev {
    int tiles, kchunks, H, mhalves, n_global, world, rank;
    int n_mma, n_stg;                      // ring depths (kFuMmaStages, kFuStages)
    float* w; float* m; float* v;          // [local rows, H]
    const float* g_extra;                  // tied: sparse-row dW_enc, added where touched
    const unsigned char* touched;
    __nv_bfloat16* shadow;                 // [local rows, shadow_ld] (+ shadow_col0) or nullptr
    int shadow_ld, shadow_col0;
    AdamConst adam;
};

__global__ void __launch_bounds__(kFuThreads, 1)
k_dw_adam_fused(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmH,
                const __grid_constant__ FusedDev p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
    const int NM = p.n_mma, NS = p.n_stg;
    uint8_t* sS = smem + NM * kFuMmaStageBytes;                      // optimizer-state staging ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFuDataBytes);
    uint64_t* full = bars;                         // [2] MMA operand stage loaded
    uint64_t* empty = full + kFuMmaStages;         // [2] MMA operand stage consumed
    uint64_t* tfull = empty + kFuMmaStages;        // [2] accumulator complete
    uint64_t* tempty = tfull + 2;                  // [2] accumulator drained
    uint64_t* ld_full = tempty + 2;                // [5][2] w / m / v rows of the stage have landed, per consuming group
    uint64_t* done = ld_full + 2 * kFuStages;      // [5] the stage holds the updated rows
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + kFuStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int hbox_rows = p.mhalves * 128;
    const int n_my = p.tiles > (int)blockIdx.x ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDz);
        tma_prefetch_desc(&tmH);
        for (int s = 0; s < kFuMmaStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], kFuEpiWarps);
        }
        for (int s = 0; s < kFuStages; ++s) {
            mbar_init(&ld_full[2 * s], 1);
            mbar_init(&ld_full[2 * s + 1], 1);
            mbar_init(&done[s], kFuEpiWarps / 2);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: MMA operands =================
        if (lane == 0) {
            const uint64_t pol_stream = policy_evict_first();
            const uint64_t pol_keep = policy_evict_last();
            const uint32_t tx = static_cast<uint32_t>(kABytes + hbox_rows * 128);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < n_my; ++t) {
                const int tile = blockIdx.x + t * gridDim.x;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1u, bars, 25);
                    mbar_expect_tx(&full[stage], tx);
                    uint8_t* dst = smem + stage * kFuMmaStageBytes;
                    tma_load_2d_hint(dst, &tmDz, &full[stage], kc * 64, tile * kTileItems, pol_stream);
                    tma_load_2d_hint(dst + kABytes, &tmH, &full[stage], kc * 64, 0, pol_keep);
                    if (++stage == NM) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, kTileItems, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < n_my; ++t) {
                const int acc = t & 1;
                mbar_wait(&tempty[acc], ((t >> 1) & 1) ^ 1u, bars, 25);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&full[stage], phase, bars, 25);
                    tc_fence_after();
                    const uint32_t dz_addr = smem_u32(smem + stage * kFuMmaStageBytes);
                    const uint32_t h_addr = dz_addr + kABytes;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t bd = umma_smem_desc(dz_addr + ks * 32, 16, 1024);
                        for (int mh = 0; mh < p.mhalves; ++mh) {
                            const uint64_t ad = umma_smem_desc(h_addr + mh * (128 * 128) + ks * 32, 16, 1024);
                            umma_bf16(d_tmem + static_cast<uint32_t>(mh * kTileItems), ad, bd, idesc, (kc | ks) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == NM) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else if (warp == 2) {
        // ================= I/O thread: optimizer state through the staging ring =================
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            const uint32_t abytes = static_cast<uint32_t>(kFuItems * p.H * 4);   // one array's rows of a stage
            const int total = n_my * kFuSub;
            auto elem_off = [&](int i) -> size_t {
                const int tile = blockIdx.x + (i / kFuSub) * gridDim.x;
                return ((size_t)tile * kTileItems + (size_t)(i % kFuSub) * kFuItems) * p.H;
            };
            auto issue_load = [&](int i) {
                const int s = i % NS;
                uint8_t* dst = sS + s * kFuStageBytes;
                const size_t off = elem_off(i);
                uint64_t* lf = &ld_full[2 * s + (i & 1)];       // the barrier of the group that consumes row group i
                mbar_expect_tx(lf, 3u * abytes);
                bulk_load_hint(dst, p.w + off, abytes, lf, pol);
                bulk_load_hint(dst + abytes, p.m + off, abytes, lf, pol);
                bulk_load_hint(dst + 2 * abytes, p.v + off, abytes, lf, pol);
            };
            for (int i = 0; i < NS && i < total; ++i) issue_load(i);
            for (int i = 0; i < total; ++i) {
                const int s = i % NS;
                mbar_wait(&done[s], static_cast<uint32_t>((i / NS) & 1), bars, 25);
                const uint8_t* src = sS + s * kFuStageBytes;
                const size_t off = elem_off(i);
                bulk_store_hint(p.w + off, src, abytes, pol);
                bulk_store_hint(p.m + off, src + abytes, abytes, pol);
                bulk_store_hint(p.v + off, src + 2 * abytes, abytes, pol);
                bulk_commit_group();
                if (i >= 1) {
                    bulk_wait_group_read<1>();           // the store of row group i-1 no longer reads its stage
                    if (i - 1 + NS < total) issue_load(i - 1 + NS);
                }
            }
            bulk_wait_group<0>();
        }
    } else if (warp >= 4) {
        // ================= epilogue warps: Adam on the accumulator columns =================
        const int ew = warp - 4;
        const int grp = ew >> 3;                       // row groups grp, grp + 2, ...
        const int q = warp & 3;                        // TMEM lane quadrant this warp may read
        const int mh = (ew & 7) >> 2;                  // M-half (hidden units 0-127 / 128-255)
        const int h = mh * 128 + q * 32 + lane;
        const bool h_ok = h < p.H;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int astride = kFuItems * p.H;            // floats between the w, m and v blocks of a stage
        for (int t = 0; t < n_my; ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            const int acc = t & 1;
            mbar_wait(&tfull[acc], static_cast<uint32_t>((t >> 1) & 1), bars, 25);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + lane_addr + static_cast<uint32_t>(acc * 256 + mh * kTileItems);
#pragma unroll 1
            for (int sub = grp; sub < kFuSub; sub += 2) {
                const int i = t * kFuSub + sub;
                const int s = i % NS;
                uint32_t r[kFuItems];
                __syncwarp();                                                    // tcgen05.ld is warp-collective
                tmem_ld8(t_addr + sub * kFuItems, r);
                // this group's barrier of the stage: with NS odd the group uses the stage every 2 * NS row groups
                mbar_wait(&ld_full[2 * s + grp], static_cast<uint32_t>((i / (2 * NS)) & 1), bars, 25);
                tmem_ld_wait();
                const int item0 = tile * kTileItems + sub * kFuItems;            // local row of column 0 of the group
                const int gitem0 = item_global(item0, p.world, p.rank);          // 8 consecutive catalogue ids
                const int nvalid = h_ok ? max(0, min(kFuItems, p.n_global - gitem0)) : 0;
                float* sw = reinterpret_cast<float*>(sS + s * kFuStageBytes) + h;
#pragma unroll
                for (int j = 0; j < kFuItems; ++j) {
                    if (j < nvalid) {
                        float* pw = sw + j * p.H;
                        float wv = pw[0], mv = pw[astride], vv = pw[2 * astride];
                        float g = __uint_as_float(r[j]);
                        if (p.g_extra != nullptr && p.touched[item0 + j] != 0)   // tied: + sparse-row dW_enc
                            g = __fadd_rn(g, __ldcs(p.g_extra + (size_t)(item0 + j) * p.H + h));
                        adam_one(wv, mv, vv, g, p.adam);
                        pw[0] = wv; pw[astride] = mv; pw[2 * astride] = vv;
                        if (p.shadow != nullptr) p.shadow[(size_t)(item0 + j) * p.shadow_ld + p.shadow_col0 + h] = __float2bfloat16_rn(wv);
                    }
                }
                fence_proxy_async_smem();              // the bulk store (async proxy) must see these writes
                __syncwarp();
                if (lane == 0) mbar_arrive(&done[s]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

void launch_dw_adam_fused(const DwArgs& a, cudaStream_t st) {
    // w / m / v must be DENSE [rows, H] (8-row groups are contiguous 1-D bulk copies).  A column block of a wider matrix
    // through 2-D tensor maps was tried for the title output layer ([N, 512]): 1 KB segments at a 2 KB pitch ran at
    // 0.56 ms per 2 GB block instead of 0.35 ms -- the title head therefore keeps its output layer as dense column blocks.
    if ((a.ld != 0 && a.ld != a.H) || a.col0 != 0 || a.H % 64 != 0 || a.H > 256) {
        fprintf(stderr, "dae_b200: launch_dw_adam_fused needs dense [rows, H <= 256] state (ld=%d col0=%d H=%d)\n", a.ld, a.col0, a.H);
        abort();
    }
    FusedDev p{};
    p.world = a.pt.world > 0 ? a.pt.world : 1;
    p.rank = a.pt.rank;
    const int tiles_total = (a.N + kTileItems - 1) / kTileItems;
    p.tiles = tiles_total > p.rank ? (tiles_total - p.rank + p.world - 1) / p.world : 0;
    if (p.tiles == 0) return;
    p.n_global = a.N;
    p.kchunks = a.K / 64;
    p.H = a.H;
    p.mhalves = (a.H + 127) / 128;
    p.w = a.w; p.m = a.m; p.v = a.v;
    p.g_extra = a.g_extra; p.touched = a.touched;
    p.shadow = a.shadow;
    p.shadow_ld = a.shadow_ld != 0 ? a.shadow_ld : a.H; p.shadow_col0 = a.shadow_col0;   // the operand copy may be a column block
    p.adam = a.adam;
    p.n_mma = a.K > 512 ? 3 : 2;
    p.n_stg = a.K > 512 ? 3 : 5;           // ODD: the per-group barrier phase in the epilogue is (i / (2 * NS)) & 1
    const CUtensorMap tmDz = make_map_bf16(a.dzT, a.K, a.n_local, kTileItems);
    const CUtensorMap tmH = make_map_bf16(a.h_dT, a.K, a.H, p.mhalves * 128);   // rows >= H: out-of-bounds zero fill
    const int grid = p.tiles < sm_count() ? p.tiles : sm_count();
    k_dw_adam_fused<<<grid, kFuThreads, kSmemFused, st>>>(tmDz, tmH, p);
}

void launch_dw(const DwArgs& a, cudaStream_t st) {
    if (a.w != nullptr) {        // Adam applied to the tile while it is in tensor memory
        launch_dw_adam_fused(a, st);
        return;
    }
    DwDev p{};
    p.pt = a.pt;
    if (p.pt.world < 1) p.pt.world = 1;
    // local tiles that hold at least one valid catalogue row: global tile = local * world + rank
    const int tiles_total = (a.N + kTileItems - 1) / kTileItems;
    p.tiles = tiles_total > p.pt.rank ? (tiles_total - p.pt.rank + p.pt.world - 1) / p.pt.world : 0;
    if (p.tiles == 0) return;
    p.n_global = a.N;
    p.kchunks = a.K / 64;
    p.H = a.H;
    p.mhalves = (a.H + 127) / 128;
    p.stream_h = a.K > 256 ? 1 : 0;
    p.g = a.g;
    p.ld = a.ld > 0 ? a.ld : a.H;
    p.col0 = a.col0;
    const CUtensorMap tmDz = make_map_bf16(a.dzT, a.K, a.n_local, kTileItems);
    const CUtensorMap tmH = make_map_bf16(a.h_dT, a.K, a.H, p.mhalves * 128);   // rows >= H: out-of-bounds zero fill
    const int grid = p.tiles < sm_count() ? p.tiles : sm_count();
    k_dw<<<grid, kDwThreads, kSmemItemTile, st>>>(tmDz, tmH, p);
}

// ------------------------------------------------------------------------------------------
// G3: dh = dz . W_dec, contraction over the catalogue (split-K), MN-major operands
// ------------------------------------------------------------------------------------------
constexpr int kDhStages = 3;
constexpr int kDhBox = 64 * 128;                       // [64 items x 64 elements] bf16 box = 8 KB
constexpr int kDhStageBytes = 8 * kDhBox;              // up to 4 boxes of dz + 4 boxes of W = 64 KB
constexpr int kSmemDh = kDhStages * kDhStageBytes + 256 + 1024;

struct DhDev {
    int kchunks_total;   // ceil(N / 64)
    int mboxes;          // bpad / 64
    int nboxes;          // H / 64
    int H, bpad;
    uint32_t lbo, sbo;
    float* partial;      // [gridDim.y][gridDim.x][bpad][H]
};

__global__ void __launch_bounds__(192, 1)
k_dh(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmW, const DhDev p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDhStages * kDhStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kDhStages;
    uint64_t* tfull = bars + 2 * kDhStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // this CTA's slice of the contraction dimension (64-item chunks), contiguous.  (A reverse wavefront over the whole
    // catalogue -- to start where G1 has just left dz and W in L2 -- was measured: 6 us SLOWER alone, no gain in the step.)
    const int per = (p.kchunks_total + gridDim.x - 1) / gridDim.x;
    const int kc0 = blockIdx.x * per;
    const int kc1 = min(kc0 + per, p.kchunks_total);
    const int nk = max(kc1 - kc0, 0);
    const int halves = p.bpad > 128 ? 2 : 1;
    const int m_rows = p.bpad > 128 ? 128 : p.bpad;   // UMMA M must be 128: bpad in {64} is handled as M=128 with zero rows

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDz);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < kDhStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    (void)m_rows;

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = policy_evict_first();
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t bytes = static_cast<uint32_t>((p.mboxes + p.nboxes) * kDhBox);
            for (int k = 0; k < nk; ++k) {
                mbar_wait(&empty[stage], phase ^ 1u);
                mbar_expect_tx(&full[stage], bytes);
                uint8_t* sa = smem + stage * kDhStageBytes;
                uint8_t* sb = sa + 4 * kDhBox;
                const int item0 = (kc0 + k) * 64;
                for (int m = 0; m < p.mboxes; ++m)
                    tma_load_2d_hint(sa + m * kDhBox, &tmDz, &full[stage], blockIdx.y * p.bpad + m * 64, item0, pol);
                for (int n = 0; n < p.nboxes; ++n) tma_load_2d_hint(sb + n * kDhBox, &tmW, &full[stage], n * 64, item0, pol);
                if (++stage == kDhStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.H), 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < nk; ++k) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * kDhStageBytes);
                const uint32_t b_addr = a_addr + 4 * kDhBox;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {          // 16 items per UMMA
                    const uint64_t bd = umma_smem_desc(b_addr + ks * 2048, p.lbo, p.sbo);
                    for (int hf = 0; hf < halves; ++hf) {
                        const uint64_t ad = umma_smem_desc(a_addr + hf * 2 * kDhBox + ks * 2048, p.lbo, p.sbo);
                        umma_bf16(tmem_base + static_cast<uint32_t>(hf * 256), ad, bd, idesc, (k | ks) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty[stage]);
                if (++stage == kDhStages) { stage = 0; phase ^= 1u; }
            }
            umma_commit(tfull);
        }
    } else {
        const int q = warp & 3;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        float* out = p.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * p.bpad * p.H;
        if (nk > 0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
        }
        for (int hf = 0; hf < halves; ++hf) {
            const int brow = hf * 128 + q * 32 + lane;
            for (int c = 0; c < (p.H >> 5); ++c) {
                uint32_t r[32];
                if (nk > 0) {
                    tmem_ld32(tmem_base + lane_addr + static_cast<uint32_t>(hf * 256 + c * 32), r);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                if (brow < p.bpad) {
                    float* dst = out + (size_t)brow * p.H + c * 32;
#pragma unroll
                    for (int v = 0; v < 4; ++v) st_global_v8(dst + 8 * v, r + 8 * v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int dh_nsplit(int N, int n_batch_tiles) {
    const int chunks = (N + 63) / 64;
    int sms = sm_count() / (n_batch_tiles > 0 ? n_batch_tiles : 1);
    if (sms < 1) sms = 1;
    return chunks < sms ? chunks : sms;
}

void launch_dh(const DhArgs& a, cudaStream_t st) {
    const int nbt = a.n_batch_tiles > 0 ? a.n_batch_tiles : 1;
    const int ld_dz = a.ld_dz > 0 ? a.ld_dz : a.bpad;
    const CUtensorMap tmDz = make_map_bf16(a.dzT, ld_dz, a.N, 64);
    const CUtensorMap tmW = make_map_bf16(a.W, a.H, a.N, 64, a.ldW);
    DhDev p{};
    p.kchunks_total = (a.N + 63) / 64;
    p.mboxes = a.bpad / 64;
    p.nboxes = a.H / 64;
    p.H = a.H;
    p.bpad = a.bpad;
    p.lbo = a.lbo > 0 ? (uint32_t)a.lbo : (uint32_t)kDhBox;   // next 64 elements along M/N: the next box
    p.sbo = a.sbo > 0 ? (uint32_t)a.sbo : 1024u;              // next 8 items along K
    p.partial = a.partial;
    k_dh<<<dim3(a.nsplit, nbt), 192, kSmemDh, st>>>(tmDz, tmW, p);
}

void preload_gemm() {
    cudaFuncAttributes a;
    cudaFuncSetAttribute(k_itemtile<MODE_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemItemTileTrain);
    cudaFuncSetAttribute(k_itemtile<MODE_PREDICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemItemTile);
    cudaFuncSetAttribute(k_itemtile<MODE_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemItemTile);
    cudaFuncSetAttribute(k_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemItemTile);
    cudaFuncSetAttribute(k_dw_adam_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFused);
    cudaFuncSetAttribute(k_dh, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDh);
    PRELOAD_KERNEL(k_itemtile<MODE_TRAIN>);
    PRELOAD_KERNEL(k_itemtile<MODE_PREDICT>);
    PRELOAD_KERNEL(k_itemtile<MODE_FILTER>);
    PRELOAD_KERNEL(k_dw);
    PRELOAD_KERNEL(k_dw_adam_fused);
    PRELOAD_KERNEL(k_dh);
    PRELOAD_KERNEL(k_sum_db_parts);
    (void)cudaGetLastError();
}

void set_trap_log_gemm(unsigned int* host_mapped) {
    cudaMemcpyToSymbol(g_trap_log, &host_mapped, sizeof(host_mapped));
}

}  // namespace dae
