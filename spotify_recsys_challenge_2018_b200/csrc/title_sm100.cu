// Title branch, tensor-core part (sm_100a): the mixed decode of DAE_title (models/DAEs.py:175-181)
//
//     y_pred[b, item] = sigmoid(feat_d[b,:] . W_out[item,:] + b_out[item]) * w_t[b]      (Char_CNN.py:75)
//                     + sigmoid(h_d[b,:]    . W_dec[item,:] + b_dec[item]) * w_p[b]      (DAEs.py:178-180)
//
// Each score goes through its own sigmoid, so one item tile needs TWO accumulators: both live in
// tensor memory side by side (2 x 256 fp32 columns = all 512) and meet in one epilogue.
//   TRAIN   epilogue: weighted BCE on y_pred (DAEs.py:194-196), d cost / d z_title (bf16, item-major), db_out
//   PREDICT epilogue: y_pred -> fp32 scores [batch, item]
// The DAE is a constant here (DAEs.py:164-171): nothing flows back into it.
//
// Pipeline: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue.  Per item tile the
// producer streams H/64 chunks of (W_dec tile, h_d) then D/64 chunks of (W_out tile, feat_d) through
// a 4-stage ring of 48 KB stages; the batch-side operands come from L2, the catalogue-side ones
// from HBM exactly once.
#include <cuda.h>
#include <cuda_runtime.h>

#include "kernels.h"
#include "tmap.h"
#include "umma.cuh"

namespace dae {

enum { TMODE_TRAIN = 0, TMODE_PREDICT = 1 };

constexpr int kTStages = 4;
constexpr int kTABytes = kTileItems * 128;            // 128 items x 64 bf16
constexpr int kTBBytes = 256 * 128;                   // <= 256 batch rows x 64 bf16
constexpr int kTStageBytes = kTABytes + kTBBytes;     // 48 KB
constexpr int kTEpiWarps = 8;
constexpr int kTThreads = 64 + 32 * kTEpiWarps;
constexpr int kSmemTitle = kTStages * kTStageBytes + 2 * 256 * 4 + 256 + 1024;

struct TitleDev {
    int n_items, tiles, kd, kf, n_cols, batch;
    const float* b_dec; const float* b_out; const float* w_t; const float* w_p;
    const uint32_t* ybits; int ywords;
    __nv_bfloat16* dzT; float* db_out; float* loss_partial; float inv_batch;
    float* out; long long ld_out; int n_out;
};

template <int MODE>
__global__ void __launch_bounds__(kTThreads, 1)
k_title_tile(const __grid_constant__ CUtensorMap tmWd, const __grid_constant__ CUtensorMap tmHd,
             const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmFd,
             const __grid_constant__ TitleDev p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
    float* s_wt = reinterpret_cast<float*>(smem + kTStages * kTStageBytes);
    float* s_wp = s_wt + 256;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_wp + 256);
    uint64_t* full = bars;                  // [kTStages]
    uint64_t* empty = bars + kTStages;      // [kTStages]
    uint64_t* tfull = bars + 2 * kTStages;  // [1]
    uint64_t* tempty = tfull + 1;           // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
    float* loss_smem = reinterpret_cast<float*>(tmem_slot + 1);   // [kTEpiWarps]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nk = p.kd + p.kf;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmWd); tma_prefetch_desc(&tmHd); tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmFd);
        for (int s = 0; s < kTStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        mbar_init(tempty, kTEpiWarps);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s_wt[i] = i < p.n_cols ? p.w_t[i] : 0.f;
        s_wp[i] = i < p.n_cols ? p.w_p[i] : 0.f;
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol_stream = policy_evict_first();
            const uint64_t pol_keep = policy_evict_last();
            const uint32_t tx = static_cast<uint32_t>(kTABytes + p.n_cols * 128);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_expect_tx(&full[stage], tx);
                    uint8_t* sa = smem + stage * kTStageBytes;
                    if (kc < p.kd) {
                        tma_load_2d_hint(sa, &tmWd, &full[stage], kc * 64, tile * kTileItems, pol_stream);
                        tma_load_2d_hint(sa + kTABytes, &tmHd, &full[stage], kc * 64, 0, pol_keep);
                    } else {
                        tma_load_2d_hint(sa, &tmWo, &full[stage], (kc - p.kd) * 64, tile * kTileItems, pol_stream);
                        tma_load_2d_hint(sa + kTABytes, &tmFd, &full[stage], (kc - p.kd) * 64, 0, pol_keep);
                    }
                    if (++stage == kTStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kTileItems, static_cast<uint32_t>(p.n_cols), 0, 0);
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                mbar_wait(tempty, tphase ^ 1u);            // the epilogue has drained both accumulators
                tc_fence_after();
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * kTStageBytes);
                    const uint32_t b_addr = a_addr + kTABytes;
                    const bool title = kc >= p.kd;
                    const uint32_t d_tmem = tmem_base + (title ? 256u : 0u);
                    const int k0 = title ? kc - p.kd : kc;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ad = umma_smem_desc(a_addr + ks * 32, 16, 1024);
                        const uint64_t bd = umma_smem_desc(b_addr + ks * 32, 16, 1024);
                        umma_bf16(d_tmem, ad, bd, idesc, (k0 | ks) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == kTStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull);
                tphase ^= 1u;
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row_in_tile = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int c_lo = half * (p.n_cols >> 1), c_hi = c_lo + (p.n_cols >> 1);
        uint32_t tphase = 0;
        float loss_acc = 0.f;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int item = tile * kTileItems + row_in_tile;
            const bool item_ok = item < p.n_items;
            const float bz = item_ok ? __ldg(p.b_dec + item) : 0.f;
            const float bt = item_ok ? __ldg(p.b_out + item) : 0.f;
            const uint32_t* yrow = (MODE == TMODE_TRAIN && item_ok) ? p.ybits + (size_t)item * p.ywords : nullptr;
            float db = 0.f;
            mbar_wait(tfull, tphase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + lane_addr;
#pragma unroll 1
            for (int cc = c_lo; cc < c_hi; cc += 16) {
                uint32_t r0[16], r1[16];
                __syncwarp();
                tmem_ld16(t_addr + cc, r0);
                tmem_ld16(t_addr + 256 + cc, r1);
                tmem_ld_wait();
                uint32_t yw = 0;
                if (MODE == TMODE_TRAIN && item_ok) yw = __ldg(yrow + (cc >> 5)) >> (cc & 31);
                uint32_t packed[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float dzv[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int b = cc + j + u;
                        const float pr = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r0[j + u]) + bz)));
                        const float tr = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(r1[j + u]) + bt)));
                        const float wt = s_wt[b];
                        const float qv = tr * wt + pr * s_wp[b];                         // DAEs.py:180
                        const bool live = item_ok && b < p.batch;
                        if (MODE == TMODE_TRAIN) {
                            const bool yb = (yw >> (j + u)) & 1u;
                            const float den = (yb ? qv : 1.f - qv) + kEpsLog;            // DAEs.py:194-195
                            const float wgt = yb ? 1.f : kNegWeight;
                            loss_acc += live ? -wgt * __logf(den) : 0.f;
                            const float dq = (yb ? -1.f : kNegWeight) * __fdividef(1.f, den) * p.inv_batch;
                            float dz = dq * wt * (tr * (1.f - tr));                      // through the mix and the title sigmoid
                            dz = live ? dz : 0.f;
                            db += dz;
                            dzv[u] = dz;
                        } else {
                            if (live && item < p.n_out) p.out[(size_t)b * p.ld_out + item] = qv;   // lanes = consecutive items
                        }
                    }
                    if (MODE == TMODE_TRAIN) packed[j >> 1] = pack_bf16x2(dzv[0], dzv[1]);
                }
                if (MODE == TMODE_TRAIN && item_ok) st_global_v8(p.dzT + (size_t)item * p.n_cols + cc, packed);
            }
            if (MODE == TMODE_TRAIN && item_ok) atomicAdd(p.db_out + item, db);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
            tphase ^= 1u;
        }
        if (MODE == TMODE_TRAIN) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
            if (lane == 0) loss_smem[warp - 2] = loss_acc;
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (MODE == TMODE_TRAIN && threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kTEpiWarps; ++i) t += loss_smem[i];
        p.loss_partial[blockIdx.x] = t;
    }
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int MODE>
static void launch_title(const TitleTileArgs& a, cudaStream_t st) {
    TitleDev p{};
    p.n_items = MODE == TMODE_PREDICT ? (a.n_out < a.N ? a.n_out : a.N) : a.N;
    p.tiles = (p.n_items + kTileItems - 1) / kTileItems;
    p.kd = a.H / 64;
    p.kf = a.kf;
    p.n_cols = a.bpad;
    p.batch = a.batch;
    p.b_dec = a.b_dec; p.b_out = a.b_out; p.w_t = a.w_t; p.w_p = a.w_p;
    p.ybits = a.ybits; p.ywords = a.ywords;
    p.dzT = a.dzT; p.db_out = a.db_out; p.loss_partial = a.loss_partial; p.inv_batch = a.inv_batch;
    p.out = a.out; p.ld_out = a.ld_out; p.n_out = a.n_out;
    const CUtensorMap tmWd = make_map_bf16(a.W_dec, a.H, a.N, kTileItems);
    const CUtensorMap tmHd = make_map_bf16(a.h_d, a.H, a.bpad, a.bpad);
    const CUtensorMap tmWo = make_map_bf16(a.W_out, kTitleFpad, a.N, kTileItems);
    const CUtensorMap tmFd = make_map_bf16(a.feat_d, kTitleFpad, a.bpad, a.bpad);
    const int grid = decode_grid(p.n_items, 1);
    k_title_tile<MODE><<<grid, kTThreads, kSmemTitle, st>>>(tmWd, tmHd, tmWo, tmFd, p);
}

void launch_title_train(const TitleTileArgs& a, cudaStream_t st) {
    cudaMemsetAsync(a.db_out, 0, sizeof(float) * a.N, st);
    launch_title<TMODE_TRAIN>(a, st);
}
void launch_title_predict(const TitleTileArgs& a, cudaStream_t st) { launch_title<TMODE_PREDICT>(a, st); }

void preload_title_gemm() {
    cudaFuncAttributes a;
    cudaFuncSetAttribute(k_title_tile<TMODE_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTitle);
    cudaFuncSetAttribute(k_title_tile<TMODE_PREDICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTitle);
    PRELOAD_KERNEL(k_title_tile<TMODE_TRAIN>);
    PRELOAD_KERNEL(k_title_tile<TMODE_PREDICT>);
    (void)cudaGetLastError();
}

void set_trap_log_title_gemm(unsigned int* host_mapped) {
    cudaMemcpyToSymbol(g_trap_log, &host_mapped, sizeof(host_mapped));
}

}  // namespace dae
