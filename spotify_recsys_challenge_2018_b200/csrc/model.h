// Internal: the model object behind the C ABI (include/dae_b200.h), shared by api.cu (DAE path) and
// api_title.cu (title branch).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dae_b200.h"
#include "kernels.h"

using namespace dae;

std::string& dae_err();                      // last error of the calling thread
int fail(const char* fmt, ...);               // records the message, returns 1
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

void ensure_loaded();                         // load every kernel once per process (kernels.h: preload_*)

// per-phase device timing (dae_model_set_profiling): events around each phase of a step / call
enum Phase { PH_PREPARE = 0, PH_ENCODE, PH_YBITS, PH_DECODE_LOSS, PH_DH, PH_DA, PH_BARRIER, PH_SCATTER, PH_DW, PH_ADAM_DEC,
             PH_ADAM_ENC, PH_ADAM_BIAS, PH_REC_A, PH_REC_B, PH_REC_C, PH_REC_SELECT,
             PH_T_FWD, PH_T_DW_ADAM, PH_T_DFEAT, PH_T_CNN_BWD, PH_T_ADAM_SMALL, PH_COUNT };
struct dae_model;
void ph_begin(dae_model* m, int k, cudaStream_t s = nullptr);
void ph_end(dae_model* m, int k, cudaStream_t s = nullptr);
void ph_collect(dae_model* m);

constexpr float kBeta1 = 0.9f, kBeta2 = 0.999f, kAdamEps = 1e-8f;   // [TF1] AdamOptimizer defaults (DAEs.py:102)
constexpr int kSqBlocks = 256;

struct Slot {
    long long *x_pos = nullptr, *y_pos = nullptr;   // device
    float *x_val = nullptr, *y_val = nullptr;
    long long *hx_pos = nullptr, *hy_pos = nullptr; // pinned host mirrors
    float *hx_val = nullptr, *hy_val = nullptr;
    int nnz_x = 0, nnz_y = 0, batch = 0, y_batch = 0;
    bool has_y = false;
    CsrWork xw{}, yw{};                             // de-duplicated CSR of this slot's batch
    cudaEvent_t h2d_done = nullptr;                 // the pinned mirror may be overwritten after this
    cudaEvent_t prepared = nullptr;                 // side stream: CSR + ybits of this slot are ready
    cudaEvent_t consumed = nullptr;                 // main stream: the step has finished reading this slot
};

// Bump allocator over ONE cudaMalloc per model.  Every rank lays its arena out identically, so a
// peer's copy of any buffer is `peer base + same offset` (kernels.h: PeerTable / peer_ptr).
struct Arena {
    char* base = nullptr;
    size_t off = 0;
    template <typename T>
    T* take(size_t n) {
        off = (off + 1023) & ~size_t(1023);
        T* p = reinterpret_cast<T*>(base + off);     // base == nullptr during the measuring pass
        off += n * sizeof(T);
        return p;
    }
};

struct dae_model {
    dae_config cfg{};
    int N = 0, T = 0, H = 0, Bmax = 0, rows_alloc = 0, max_nnz = 0;
    int world = 1, rank = 0;
    int n_local = 0;                // catalogue rows held by every rank (local tiles x 128)
    bool tied = false, trainable = true, needs_y = true, own_stream = false, attached = false;
    cudaStream_t st = nullptr;      // main stream: the step
    cudaStream_t st2 = nullptr;     // side stream: H2D + COO->CSR + ybits of the NEXT batch, overlapped with the step
    cudaStream_t st3 = nullptr;     // decoder-update stream: k_dw_adam_fused overlaps the sparse / encoder tail of the step
    cudaStream_t st4 = nullptr;     // background stream: encoder Adam of the untouched rows under the compute-bound front of the step
    cudaEvent_t ev_touch = nullptr, ev_bg = nullptr, ev_pre = nullptr;
    BgAdam bg{};
    unsigned long long* trace = nullptr;   // debug bit 13: %globaltimer stamps of the step's fork / join points
    bool colsum_deferred = false;   // this step's db_enc column sums run on the bias stream (apply_adam)
    bool gather_deferred = false;   // world > 1: this step's db_dec / cost gathers run on the bias stream (apply_adam)
    bool enc_early = false;         // world >= 4: the unlisted-row encoder pass was launched at the start of the step (st4)
    bool enc_split = false;         // this step's encoder Adam of the unlisted rows already follows the decoder update on st3
    bool bg_inflight = false;       // this step launched the background streamer (the dense encoder pass must account for it)
    int row_offset0 = 0;            // global row of rank 0's first playlist in the dropout keys of the step in flight
    cudaEvent_t ev_dh = nullptr, ev_dec = nullptr, ev_a = nullptr, ev_y = nullptr, ev_da = nullptr, ev_bias = nullptr;
    bool par_step = false;          // this step forks work onto st3 (whole-step call, not profiling, debug bit 3 clear)
    bool overlap_dec = false;       // set by dae_model_train_step_staged: launch the decoder update as soon as dz is final
    bool dec_inflight = false;      // this step's decoder update is already running on st3
    int cur = 0;                    // slot used by the last step (dae_model_buffer)
    Arena arena;
    PeerTable pt{};
    void* ipc_opened[kMaxWorld] = {};
    // parameters: catalogue matrices are row-sharded (tile-cyclic), biases replicated
    float *W_enc = nullptr, *W_dec = nullptr, *b_enc = nullptr, *b_dec = nullptr;
    __nv_bfloat16* shadow = nullptr;                 // bf16 decoder operand: the rows this rank owns (training)
    __nv_bfloat16* shadow_full = nullptr;            // all N rows in catalogue order (inference); == shadow when world == 1
    bool full_stale = true;
    float *mW_enc = nullptr, *vW_enc = nullptr, *mW_dec = nullptr, *vW_dec = nullptr;
    float *mb_enc = nullptr, *vb_enc = nullptr, *mb_dec = nullptr, *vb_dec = nullptr;
    float b1_pow = kBeta1, b2_pow = kBeta2;
    long long step = 0;
    float *g_enc = nullptr;                          // sparse-row dW_enc of the rows this rank owns
    unsigned char* touched = nullptr;
    int* touched_list = nullptr;                     // [1 + rows]: rows listed by any playlist of the global batch
    int* hot_list = nullptr;                         // [1 + rows]: rows listed by >= 3 playlists (ordered gather of dW_enc)
    int* touch_cnt = nullptr;                        // playlists of the global batch that list each owned row
    float *g_b_enc = nullptr, *g_b_dec_sh = nullptr, *g_b_dec = nullptr, *g_b_dec_parts = nullptr;
    uint32_t* ybits = nullptr;                       // targets of the global batch over this rank's item rows
    float* g_dec = nullptr;                          // dW_dec of the rows this rank owns
    int debug = 0;
    bool scatter_done = false;
    Slot slots[2];
    PubInput pub{};
    int* err = nullptr;
    int* err_host = nullptr;
    unsigned int* flags = nullptr;
    unsigned int epoch = 0;
    int ywords = 8;
    float *rowsum = nullptr, *h = nullptr, *da = nullptr, *dh_partial = nullptr;
    __nv_bfloat16 *h_d = nullptr, *h_dT = nullptr, *dz_all = nullptr;
    float* dh_sum = nullptr;
    int nsplit = 0;
    float *loss_partial = nullptr, *sq_partial = nullptr, *cost_part = nullptr, *cost = nullptr, *cost_host = nullptr;
    int n_loss_partial = 0;
    float* scores = nullptr;
    size_t scores_elems = 0;
    int *topk_idx = nullptr, *seed_ptr = nullptr, *seed_idx = nullptr;
    float* topk_score = nullptr;
    size_t topk_elems = 0, seed_idx_elems = 0, seed_ptr_elems = 0;
    int *ans_ptr = nullptr, *ans_idx = nullptr;     // answers CSR of the batch being evaluated (dae_model_evaluate)
    double* metrics = nullptr;                      // [batch, 3] r-precision, ndcg, clicks
    size_t ans_ptr_elems = 0, ans_idx_elems = 0, metrics_elems = 0;
    // fused decode + top-K (large catalogues): per-playlist candidate lists, thresholds, intermediate top-K
    float* cand_val = nullptr; int* cand_idx = nullptr; int* cand_cnt = nullptr; float* cand_thr = nullptr;
    int* cand_tk_idx = nullptr; float* cand_tk_score = nullptr; int* cand_cnt_host = nullptr;
    size_t cand_rows = 0;
    int last_batch = 0, last_bpad = 0;
    long long launches = 0;
    // pipelined train loop (dae_model_train_step_async): the host runs one step ahead of the device
    int async_slot = 1;
    bool async_pending = false;
    cudaEvent_t ev_cost[2] = {nullptr, nullptr};
    float* cost_ring = nullptr;      // pinned [2]
    int* err_ring = nullptr;         // pinned [2]
    struct dae_exchange* thr_exchange = nullptr;   // item-sharded inference: filter thresholds are shared across the shards
    // optional per-phase device timing (bench.py roofline): events around each phase of a step
    bool profiling = false;
    cudaEvent_t ph_ev[2 * PH_COUNT] = {};
    bool ph_used[PH_COUNT] = {};
    double ph_ms[PH_COUNT] = {};
    long long ph_n[PH_COUNT] = {};
    std::vector<void*> host_allocs;
};

template <typename T>
static int halloc(dae_model* m, T** p, size_t n) {
    CK(cudaMallocHost(reinterpret_cast<void**>(p), n * sizeof(T)));
    m->host_allocs.push_back(*p);
    return 0;
}
#define TRY(x) do { if (int rc_ = (x)) return rc_; } while (0)


// api_dp.cu: filter thresholds shared across item shards (collective on `st`; thr == nullptr: barrier only)
struct dae_exchange;
int exchange_world(const dae_exchange* x);
int exchange_min_thresholds(dae_exchange* x, float* thr, int n, cudaStream_t st);
constexpr int kThrExchangesPerCall = 2;         // every rank runs exactly this many per recommend call, whatever its path

// api.cu internals used by the title branch
int stage_impl(dae_model* m, int32_t slot, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
               const int64_t* y_pos, const float* y_val, int64_t nnz_y, int32_t batch, bool with_y);
int check_device_flag(dae_model* m);
void run_encode(dae_model* m, int slot, int bpad, int rows_pad, float kp, float kp_in, int row_offset, bool train);
void build_ybits(dae_model* m, int slot, int B, int bpad, cudaStream_t st = nullptr);
int run_metrics(dae_model* m, const int* idx_dev, int32_t batch, int32_t k, const int32_t* ans_ptr, const int32_t* ans_idx,
                double* out_host);
