// C-ABI of libdae_b200.so (include/dae_b200.h): the model object (parameters, TF1-Adam state,
// workspaces, staging) and the orchestration of one train / predict / recommend step.
#include <cuda_runtime.h>
#include <algorithm>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "model.h"
#include "philox.cuh"

static thread_local std::string g_err;
std::string& dae_err() { return g_err; }
static unsigned int* g_trap_host = nullptr;      // mapped pinned: where a bounded device-side spin gave up (umma.cuh)
int fail(const char* fmt, ...) {
    char buf[768];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, 512, fmt, ap);
    va_end(ap);
    if (g_trap_host != nullptr && g_trap_host[0] == 0xDAE0DEADu && n > 0 && n < 512) {
        volatile unsigned int* l = g_trap_host;
        if (l[1] == 0xBA221E2u)
            snprintf(buf + n, sizeof(buf) - n, " [device trap: cross-GPU barrier timed out on rank %u waiting for rank %u, epoch %u, seen %u]",
                     l[2], l[3], l[4], l[5]);
        else
            snprintf(buf + n, sizeof(buf) - n, " [device trap: mbarrier wait timed out: blockDim %u block (%u,%u) thread %u barrier smem 0x%x parity %u; odd-phase mask 0x%x of the barriers at 0x%x]",
                     l[1], l[2], l[3], l[4], l[5], l[6], l[7], l[8]);
    }
    g_err = buf;
    return 1;
}
__global__ void k_max_int(const int* __restrict__ v, int n, int* __restrict__ out);
__global__ void k_stamp(unsigned long long* t);
void ensure_loaded() {
    static bool done = false;
    if (done) return;
    preload_sparse(); preload_optim(); preload_gemm(); preload_topk(); preload_title();
    {
        cudaFuncAttributes a;
        PRELOAD_KERNEL(k_stamp);
        PRELOAD_KERNEL(k_max_int);
    }
    if (cudaHostAlloc(reinterpret_cast<void**>(&g_trap_host), 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
        memset(g_trap_host, 0, 64);
        unsigned int* dptr = nullptr;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), g_trap_host, 0) == cudaSuccess) {
            set_trap_log_gemm(dptr); set_trap_log_title_gemm(dptr); set_trap_log_sparse(dptr); set_trap_log_optim(dptr);
        }
    } else {
        g_trap_host = nullptr;
    }
    (void)cudaGetLastError();
    done = true;
}

// debug bit 13: device-side time stamps (%globaltimer) of the step's fork / join points, one tiny kernel each, into the
// "trace" buffer -- the only way to see how the streams of a step overlap without a timeline profiler (tools/gpu_trace.py)
__global__ void k_stamp(unsigned long long* t) {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    *t = v;
}
enum { TR_START = 0, TR_ENCODE, TR_YBITS, TR_DECODE, TR_DH, TR_DEC_BEGIN, TR_DEC_END, TR_BG_BEGIN, TR_BG_END, TR_TAIL, TR_REST_BEGIN,
       TR_REST_END, TR_G1_FIRST, TR_G1_LAST, TR_STEP_END, TR_PREPARED, TR_COUNT };
// two records of 16 stamps, alternating by step parity: the last two steps of a pipelined run can be read side by side
static inline void stamp(dae_model* m, cudaStream_t s, int id) {
    if (m->debug & 8192) k_stamp<<<1, 1, 0, s>>>(m->trace + 16 * (m->step & 1) + id);
}

static void layout_csr(Arena& A, CsrWork* w, int B, int max_nnz) {
    w->cnt = A.take<int>(B);
    w->row_ptr = A.take<int>(B + 1);
    w->cursor = A.take<int>(B);
    w->keys = A.take<unsigned long long>(max_nnz);
    w->row_len = A.take<int>(B);
    w->col = A.take<int>(max_nnz);
    w->val = A.take<float>(max_nnz);
}

// Assign every device buffer of the model from the arena (called twice: measure, then place).
static void layout(dae_model* m) {
    Arena& A = m->arena;
    A.off = 0;
    const size_t LH = (size_t)m->n_local * m->H, NH = (size_t)m->N * m->H;
    const int N = m->N, H = m->H, K = m->world * kMaxBpad;
    const int rows_h = K > m->rows_alloc ? K : m->rows_alloc;            // training: the global batch; inference: batch tiles
    m->flags = A.take<unsigned int>(kMaxWorld);
    m->W_enc = A.take<float>(LH);
    m->b_enc = A.take<float>(H);
    m->b_dec = A.take<float>(N);
    m->W_dec = m->tied ? m->W_enc : A.take<float>(LH);
    m->shadow = A.take<__nv_bfloat16>(LH);                               // this rank's rows of the decoder operand
    m->shadow_full = m->world > 1 ? A.take<__nv_bfloat16>(NH) : m->shadow;   // all rows, gathered lazily for inference
    if (m->trainable) {
        m->mW_enc = A.take<float>(LH); m->vW_enc = A.take<float>(LH);
        if (m->tied) { m->mW_dec = m->mW_enc; m->vW_dec = m->vW_enc; }
        else { m->mW_dec = A.take<float>(LH); m->vW_dec = A.take<float>(LH); }
        m->mb_enc = A.take<float>(H); m->vb_enc = A.take<float>(H);
        m->mb_dec = A.take<float>(N); m->vb_dec = A.take<float>(N);
        m->g_dec = A.take<float>(LH);
        m->g_enc = A.take<float>(LH);
        m->touched = A.take<unsigned char>(m->n_local);
        m->touch_cnt = A.take<int>(m->n_local);
        m->hot_list = A.take<int>((size_t)m->n_local + 1);
        m->touched_list = A.take<int>((size_t)m->n_local + 1);
        m->g_b_enc = A.take<float>(H);
        m->g_b_dec_sh = A.take<float>(m->n_local);
        m->g_b_dec_parts = A.take<float>((size_t)4 * m->world * m->n_local);
        m->g_b_dec = m->world > 1 ? A.take<float>(N) : m->g_b_dec_sh;
        m->da = A.take<float>((size_t)K * H);
        m->dz_all = A.take<__nv_bfloat16>((size_t)m->n_local * K);
        m->nsplit = dh_nsplit(m->n_local, m->world);
        m->dh_partial = A.take<float>((size_t)m->world * m->nsplit * kMaxBpad * H);
        m->dh_sum = A.take<float>((size_t)K * H);
        m->n_loss_partial = 148 * 2;
        m->loss_partial = A.take<float>(m->n_loss_partial);
        m->sq_partial = A.take<float>(4 * kSqBlocks);
        m->cost_part = A.take<float>(1);
        m->cost = m->world > 1 ? A.take<float>(1) : m->cost_part;
    }
    if (m->needs_y) m->ybits = A.take<uint32_t>((size_t)m->n_local * (K / 32));
    m->err = A.take<int>(1);
    m->rowsum = A.take<float>(m->Bmax);
    m->h = A.take<float>((size_t)rows_h * H);
    m->h_d = A.take<__nv_bfloat16>((size_t)rows_h * H);
    m->h_dT = A.take<__nv_bfloat16>((size_t)H * rows_h);
    // published input of the global batch: training keeps every rank's segment on every rank (k_encode_fwd)
    const int n_seg = m->trainable ? m->world : 1;
    m->pub.seg_rows = m->Bmax; m->pub.seg_nnz = m->max_nnz;
    m->pub.row_ptr = A.take<int>((size_t)n_seg * m->Bmax);
    m->pub.row_len = A.take<int>((size_t)n_seg * m->Bmax);
    m->pub.col = A.take<int>((size_t)n_seg * m->max_nnz);
    m->pub.xn = A.take<float>((size_t)n_seg * m->max_nnz);
    if (m->trainable) m->bg.ctl = A.take<unsigned int>(kBgCtlWords);
    m->trace = A.take<unsigned long long>(32);
    for (int s = 0; s < 2; ++s) {
        Slot& sl = m->slots[s];
        layout_csr(A, &sl.xw, m->Bmax, m->max_nnz);
        sl.x_pos = A.take<long long>((size_t)m->max_nnz * 2);
        sl.x_val = A.take<float>(m->max_nnz);
        if (m->needs_y) {
            layout_csr(A, &sl.yw, m->Bmax, m->max_nnz);
            sl.y_pos = A.take<long long>((size_t)m->max_nnz * 2);
            sl.y_val = A.take<float>(m->max_nnz);
        }
    }
    A.off = (A.off + 1023) & ~size_t(1023);
}

static const char* kPhaseNames[PH_COUNT] = {"prepare_csr", "encode_fwd", "ybits", "decode_loss_dz", "dh", "da_all",
                                            "barriers", "scatter_dw_enc", "dw_dec", "adam_dec", "adam_enc", "adam_bias",
                                            "rec_dense_prefix", "rec_filter_mid", "rec_filter_full", "rec_select",
                                            "title_fwd_loss", "title_dw_adam", "title_dfeat", "title_cnn_bwd", "title_adam_small"};
void ph_begin(dae_model* m, int k, cudaStream_t s) {
    if (m->profiling) { cudaEventRecord(m->ph_ev[2 * k], s ? s : m->st); }
}
void ph_end(dae_model* m, int k, cudaStream_t s) {
    if (m->profiling) { cudaEventRecord(m->ph_ev[2 * k + 1], s ? s : m->st); m->ph_used[k] = true; }
}
void ph_collect(dae_model* m) {
    if (!m->profiling) return;
    cudaStreamSynchronize(m->st2);
    cudaStreamSynchronize(m->st);
    for (int k = 0; k < PH_COUNT; ++k) {
        if (!m->ph_used[k]) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, m->ph_ev[2 * k], m->ph_ev[2 * k + 1]) == cudaSuccess) { m->ph_ms[k] += ms; m->ph_n[k] += 1; }
        m->ph_used[k] = false;
    }
}

extern "C" int32_t dae_abi_version(void) { return DAE_B200_ABI_VERSION; }
extern "C" const char* dae_last_error(void) { return g_err.c_str(); }

extern "C" int32_t dae_model_create(const dae_config* cfg, dae_model** out) {
    if (!cfg || !out) return fail("null argument");
    *out = nullptr;
    if (cfg->n_hidden <= 0 || cfg->n_hidden % 64 != 0 || cfg->n_hidden > 256)
        return fail("n_hidden must be a multiple of 64 in [64,256], got %d", cfg->n_hidden);
    if (cfg->n_input <= 0 || cfg->n_tracks <= 0 || cfg->n_tracks > cfg->n_input)
        return fail("need 0 < n_tracks <= n_input (got %d, %d)", cfg->n_tracks, cfg->n_input);
    if (cfg->max_batch <= 0) return fail("max_batch must be positive");
    if (cfg->trainable != 0 && cfg->max_batch > kMaxBpad)
        return fail("training batch per GPU is limited to %d rows (one tensor-core batch tile); got %d", kMaxBpad,
                    cfg->max_batch);
    const int world = cfg->world > 0 ? cfg->world : 1;
    if (world > kMaxWorld || cfg->rank < 0 || cfg->rank >= world)
        return fail("need 1 <= world <= %d and 0 <= rank < world (got world %d, rank %d)", kMaxWorld, cfg->world, cfg->rank);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("no CUDA device: libdae_b200 has no CPU fallback");
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail("device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);

    ensure_loaded();

    dae_model* m = new dae_model();
    m->cfg = *cfg;
    m->N = cfg->n_input; m->T = cfg->n_tracks; m->H = cfg->n_hidden; m->Bmax = cfg->max_batch;
    m->tied = cfg->tied != 0; m->trainable = cfg->trainable == 1; m->needs_y = cfg->trainable != 0;
    m->world = world; m->rank = cfg->rank;
    if (cfg->stream) { m->st = reinterpret_cast<cudaStream_t>(cfg->stream); }
    else { CK(cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking)); m->own_stream = true; }
    CK(cudaStreamCreateWithFlags(&m->st2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&m->st3, cudaStreamNonBlocking));
    {   // the background streamer's blocks must be on the SMs BEFORE the encode's (one per SM, see k_adam_bg): its stream
        // has the highest priority, and the encode is released by the same event as the streamer
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&m->st4, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&m->ev_pre, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_touch, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_bg, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_dh, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_dec, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_y, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_da, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_bias, cudaEventDisableTiming));
    m->rows_alloc = m->Bmax <= kMaxBpad ? round_up(m->Bmax, 64) : round_up(m->Bmax, kMaxBpad);
    m->max_nnz = m->Bmax * 1024;
    const int tiles_total = (m->N + kTileItems - 1) / kTileItems;
    m->n_local = (tiles_total + world - 1) / world * kTileItems;

    layout(m);                                            // measuring pass
    const size_t bytes = m->arena.off;
    CK(cudaMalloc(reinterpret_cast<void**>(&m->arena.base), bytes));
    CK(cudaMemsetAsync(m->arena.base, 0, bytes, m->st));
    layout(m);
    m->pt.world = world; m->pt.rank = m->rank;
    m->pt.base[m->rank] = m->arena.base;
    m->attached = world == 1;

    TRY(halloc(m, &m->err_host, 1));
    TRY(halloc(m, &m->cost_host, 1));
    TRY(halloc(m, &m->cost_ring, 2));
    TRY(halloc(m, &m->err_ring, 2));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&m->ev_cost[i], cudaEventDisableTiming));
    for (int s = 0; s < 2; ++s) {
        Slot& sl = m->slots[s];
        CK(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&sl.prepared, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&sl.consumed, cudaEventDisableTiming));
        TRY(halloc(m, &sl.hx_pos, (size_t)m->max_nnz * 2)); TRY(halloc(m, &sl.hx_val, m->max_nnz));
        if (m->needs_y) { TRY(halloc(m, &sl.hy_pos, (size_t)m->max_nnz * 2)); TRY(halloc(m, &sl.hy_val, m->max_nnz)); }
    }
    CK(cudaStreamSynchronize(m->st));
    *out = m;
    return 0;
}

extern "C" void dae_model_destroy(dae_model* m) {
    if (!m) return;
    cudaStreamSynchronize(m->st);
    cudaStreamSynchronize(m->st2);
    if (m->st3) cudaStreamSynchronize(m->st3);
    if (m->st4) cudaStreamSynchronize(m->st4);
    for (int r = 0; r < kMaxWorld; ++r) if (m->ipc_opened[r]) cudaIpcCloseMemHandle(m->ipc_opened[r]);
    if (m->arena.base) cudaFree(m->arena.base);
    for (void* p : m->host_allocs) cudaFreeHost(p);
    if (m->scores) cudaFree(m->scores);
    if (m->topk_idx) cudaFree(m->topk_idx);
    if (m->topk_score) cudaFree(m->topk_score);
    if (m->seed_ptr) cudaFree(m->seed_ptr);
    if (m->seed_idx) cudaFree(m->seed_idx);
    cudaFree(m->ans_ptr); cudaFree(m->ans_idx); cudaFree(m->metrics);
    cudaFree(m->cand_val); cudaFree(m->cand_idx); cudaFree(m->cand_cnt); cudaFree(m->cand_thr);
    cudaFree(m->cand_tk_idx); cudaFree(m->cand_tk_score);
    for (int s = 0; s < 2; ++s) {
        if (m->slots[s].h2d_done) cudaEventDestroy(m->slots[s].h2d_done);
        if (m->slots[s].prepared) cudaEventDestroy(m->slots[s].prepared);
        if (m->slots[s].consumed) cudaEventDestroy(m->slots[s].consumed);
    }
    if (m->ph_ev[0]) for (int i = 0; i < 2 * PH_COUNT; ++i) cudaEventDestroy(m->ph_ev[i]);
    for (int i = 0; i < 2; ++i) if (m->ev_cost[i]) cudaEventDestroy(m->ev_cost[i]);
    cudaStreamDestroy(m->st2);
    if (m->st3) cudaStreamDestroy(m->st3);
    if (m->st4) cudaStreamDestroy(m->st4);
    if (m->ev_touch) cudaEventDestroy(m->ev_touch);
    if (m->ev_bg) cudaEventDestroy(m->ev_bg);
    if (m->ev_pre) cudaEventDestroy(m->ev_pre);
    if (m->ev_dh) cudaEventDestroy(m->ev_dh);
    if (m->ev_dec) cudaEventDestroy(m->ev_dec);
    if (m->ev_a) cudaEventDestroy(m->ev_a);
    if (m->ev_y) cudaEventDestroy(m->ev_y);
    if (m->ev_da) cudaEventDestroy(m->ev_da);
    if (m->ev_bias) cudaEventDestroy(m->ev_bias);
    if (m->own_stream) cudaStreamDestroy(m->st);
    delete m;
}

// ------------------------------------------------------------------------------------------
// data-parallel attachment: map every peer's arena (one process per GPU)
// ------------------------------------------------------------------------------------------
extern "C" int32_t dae_model_arena_bytes(dae_model* m, int64_t* bytes) {
    if (!m || !bytes) return fail("null argument");
    *bytes = (int64_t)m->arena.off;
    return 0;
}

extern "C" int32_t dae_model_ipc_handle(dae_model* m, void* handle_out64) {
    if (!m || !handle_out64) return fail("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, m->arena.base));
    memcpy(handle_out64, &h, 64);
    return 0;
}

extern "C" int32_t dae_model_attach_ipc(dae_model* m, const void* handles, int32_t n_handles) {
    if (!m || !handles) return fail("null argument");
    if (n_handles != m->world) return fail("expected %d IPC handles, got %d", m->world, n_handles);
    CK(cudaSetDevice(m->cfg.device));
    for (int r = 0; r < m->world; ++r) {
        if (r == m->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * r, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m->ipc_opened[r] = p;
        m->pt.base[r] = static_cast<char*>(p);
    }
    m->attached = true;
    return 0;
}

extern "C" int32_t dae_model_attach_local(dae_model* m, dae_model* const* peers, int32_t n_peers) {
    if (!m || !peers) return fail("null argument");
    if (n_peers != m->world) return fail("expected %d peers, got %d", m->world, n_peers);
    for (int r = 0; r < m->world; ++r) {
        if (!peers[r] || peers[r]->world != m->world || peers[r]->rank != r || peers[r]->arena.off != m->arena.off)
            return fail("peer %d does not match this model's layout", r);
        if (peers[r]->cfg.device != m->cfg.device) {
            CK(cudaSetDevice(m->cfg.device));
            cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->cfg.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            (void)cudaGetLastError();
        }
        m->pt.base[r] = peers[r]->arena.base;
    }
    m->attached = true;
    return 0;
}

static void barrier(dae_model* m) {
    if (m->world == 1) return;
    m->epoch += 1;
    launch_barrier(m->flags, m->epoch, m->pt, m->st);
    m->launches += 1;
}

// ------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------
// Host matrix [N,H] (global row order) <-> the local-tile-ordered rows of rank `owner`: global tile
// owner + k*world is local tile k, so one strided 2-D copy moves every full tile.
static int copy_rows(dae_model* m, float* dev_local, float* host, int owner, bool to_device) {
    const int N = m->N, H = m->H, W = m->world;
    const int tiles_total = (N + kTileItems - 1) / kTileItems;
    const size_t tile_bytes = (size_t)kTileItems * H * 4;
    const int full_tiles = N / kTileItems;                       // global tiles that are complete
    const int n_own = tiles_total > owner ? (tiles_total - owner + W - 1) / W : 0;
    int n_full = full_tiles > owner ? (full_tiles - owner + W - 1) / W : 0;
    if (n_full > n_own) n_full = n_own;
    float* h0 = host + (size_t)owner * kTileItems * H;
    if (n_full > 0) {
        if (to_device) CK(cudaMemcpy2DAsync(dev_local, tile_bytes, h0, tile_bytes * W, tile_bytes, n_full, cudaMemcpyHostToDevice, m->st));
        else CK(cudaMemcpy2DAsync(h0, tile_bytes * W, dev_local, tile_bytes, tile_bytes, n_full, cudaMemcpyDeviceToHost, m->st));
    }
    if (n_own > n_full) {                                        // the catalogue's last, partial tile
        const int gt = owner + n_full * W;
        const size_t rows = (size_t)N - (size_t)gt * kTileItems;
        float* hp = host + (size_t)gt * kTileItems * H;
        float* dp = dev_local + (size_t)n_full * kTileItems * H;
        if (to_device) CK(cudaMemcpyAsync(dp, hp, rows * H * 4, cudaMemcpyHostToDevice, m->st));
        else CK(cudaMemcpyAsync(hp, dp, rows * H * 4, cudaMemcpyDeviceToHost, m->st));
    }
    return 0;
}

static void refresh_shadow(dae_model* m) {
    launch_cast_bf16(m->W_dec, m->shadow, (long long)m->n_local * m->H, m->st);
    m->full_stale = true;
    m->launches += 1;
}

extern "C" int32_t dae_model_init_xavier(dae_model* m, uint64_t seed) {
    if (!m) return fail("null model");
    const float lim = sqrtf(6.0f / (float)(m->N + m->H));
    launch_xavier_init(m->W_enc, m->n_local, m->N, m->H, lim, seed, kStreamInit, m->world, m->rank, m->st);
    if (!m->tied) launch_xavier_init(m->W_dec, m->n_local, m->N, m->H, lim, seed, kStreamInit + 1, m->world, m->rank, m->st);
    CK(cudaMemsetAsync(m->b_enc, 0, sizeof(float) * m->H, m->st));
    CK(cudaMemsetAsync(m->b_dec, 0, sizeof(float) * m->N, m->st));
    m->launches += m->tied ? 1 : 2;
    refresh_shadow(m);
    CK(cudaStreamSynchronize(m->st));
    return 0;
}

extern "C" int32_t dae_model_set_params(dae_model* m, const float* W_enc, const float* W_dec, const float* b_enc,
                                        const float* b_dec) {
    if (!m || !W_enc || !b_enc || !b_dec) return fail("null argument");
    if (!m->tied && !W_dec) return fail("W_dec required for the untied model");
    TRY(copy_rows(m, m->W_enc, const_cast<float*>(W_enc), m->rank, true));
    if (!m->tied) TRY(copy_rows(m, m->W_dec, const_cast<float*>(W_dec), m->rank, true));
    CK(cudaMemcpyAsync(m->b_enc, b_enc, (size_t)m->H * 4, cudaMemcpyHostToDevice, m->st));
    CK(cudaMemcpyAsync(m->b_dec, b_dec, (size_t)m->N * 4, cudaMemcpyHostToDevice, m->st));
    refresh_shadow(m);
    CK(cudaStreamSynchronize(m->st));
    return 0;
}

static int gather_rows(dae_model* m, float* host, float* dev_local) {
    if (!host) return 0;
    for (int r = 0; r < m->world; ++r) TRY(copy_rows(m, peer_ptr(m->pt, r, dev_local), host, r, false));
    return 0;
}

// With world > 1 the caller must have synchronised every rank (no step in flight on any GPU).
extern "C" int32_t dae_model_get_params(dae_model* m, float* W_enc, float* W_dec, float* b_enc, float* b_dec) {
    if (!m) return fail("null model");
    if (!m->attached) return fail("peers are not attached");
    TRY(gather_rows(m, W_enc, m->W_enc));
    TRY(gather_rows(m, W_dec, m->W_dec));
    if (b_enc) CK(cudaMemcpyAsync(b_enc, m->b_enc, (size_t)m->H * 4, cudaMemcpyDeviceToHost, m->st));
    if (b_dec) CK(cudaMemcpyAsync(b_dec, m->b_dec, (size_t)m->N * 4, cudaMemcpyDeviceToHost, m->st));
    CK(cudaStreamSynchronize(m->st));
    return 0;
}

extern "C" int32_t dae_model_get_adam_state(dae_model* m, float* m_W_enc, float* v_W_enc, float* m_W_dec,
                                            float* v_W_dec, int64_t* step) {
    if (!m || !m->trainable) return fail("model is not trainable");
    if (!m->attached) return fail("peers are not attached");
    TRY(gather_rows(m, m_W_enc, m->mW_enc));
    TRY(gather_rows(m, v_W_enc, m->vW_enc));
    TRY(gather_rows(m, m_W_dec, m->mW_dec));
    TRY(gather_rows(m, v_W_dec, m->vW_dec));
    if (step) *step = m->step;
    CK(cudaStreamSynchronize(m->st));
    return 0;
}

// ------------------------------------------------------------------------------------------
// staging
// ------------------------------------------------------------------------------------------
// Device-side preparation of a staged batch on the SIDE stream: COO -> CSR (x and y), y bitmask.
// Independent of the parameters, so it overlaps with the previous step's GEMMs / Adam on the main stream.
static int prepare_slot(dae_model* m, int slot) {
    Slot& s = m->slots[slot];
    CK(cudaStreamWaitEvent(m->st2, s.consumed, 0));       // the previous step on this slot no longer reads it
    ph_begin(m, PH_PREPARE, m->st2);
    launch_coo_to_csr(s.x_pos, s.x_val, s.nnz_x, s.batch, m->N, s.xw, m->err, m->st2);
    m->launches += s.nnz_x > 0 ? 4 : 2;
    if (s.has_y) {
        launch_coo_to_csr(s.y_pos, s.y_val, s.nnz_y, s.batch, m->N, s.yw, m->err, m->st2);
        m->launches += s.nnz_y > 0 ? 4 : 2;
    }
    ph_end(m, PH_PREPARE, m->st2);
    if (m->trainable) stamp(m, m->st2, TR_PREPARED);          // (lands in the record of the step being enqueued next)
    CK(cudaEventRecord(s.prepared, m->st2));
    return 0;
}

int stage_impl(dae_model* m, int32_t slot, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                      const int64_t* y_pos, const float* y_val, int64_t nnz_y, int32_t batch, bool with_y) {
    if (!m) return fail("null model");
    if (slot < 0 || slot > 1) return fail("slot must be 0 or 1");
    if (batch <= 0 || batch > m->Bmax) return fail("batch %d outside (0, %d]", batch, m->Bmax);
    if (nnz_x < 0 || nnz_x > m->max_nnz || nnz_y < 0 || nnz_y > m->max_nnz)
        return fail("nnz (%lld, %lld) exceeds the staging capacity %d", (long long)nnz_x, (long long)nnz_y, m->max_nnz);
    if ((nnz_x > 0 && (!x_pos || !x_val)) || (nnz_y > 0 && (!y_pos || !y_val))) return fail("null COO pointer");
    if (nnz_y > 0 && !m->needs_y) return fail("y given to an inference-only model");
    Slot& s = m->slots[slot];
    CK(cudaEventSynchronize(s.h2d_done));   // previous H2D out of this slot's pinned mirror has finished
    CK(cudaStreamWaitEvent(m->st2, s.consumed, 0));
    s.nnz_x = (int)nnz_x; s.nnz_y = (int)nnz_y; s.batch = batch;
    s.has_y = with_y;
    if (nnz_x > 0) {
        memcpy(s.hx_pos, x_pos, (size_t)nnz_x * 16);
        memcpy(s.hx_val, x_val, (size_t)nnz_x * 4);
        CK(cudaMemcpyAsync(s.x_pos, s.hx_pos, (size_t)nnz_x * 16, cudaMemcpyHostToDevice, m->st2));
        CK(cudaMemcpyAsync(s.x_val, s.hx_val, (size_t)nnz_x * 4, cudaMemcpyHostToDevice, m->st2));
    }
    if (nnz_y > 0) {
        memcpy(s.hy_pos, y_pos, (size_t)nnz_y * 16);
        memcpy(s.hy_val, y_val, (size_t)nnz_y * 4);
        CK(cudaMemcpyAsync(s.y_pos, s.hy_pos, (size_t)nnz_y * 16, cudaMemcpyHostToDevice, m->st2));
        CK(cudaMemcpyAsync(s.y_val, s.hy_val, (size_t)nnz_y * 4, cudaMemcpyHostToDevice, m->st2));
    }
    CK(cudaEventRecord(s.h2d_done, m->st2));
    return prepare_slot(m, slot);
}

extern "C" int32_t dae_model_stage_batch(dae_model* m, int32_t slot, const int64_t* x_pos, const float* x_val,
                                         int64_t nnz_x, const int64_t* y_pos, const float* y_val, int64_t nnz_y,
                                         int32_t batch) {
    // a trainable model stages training batches (targets may legitimately be empty)
    return stage_impl(m, slot, x_pos, x_val, nnz_x, y_pos, y_val, nnz_y, batch, m && m->needs_y);
}

// Re-run the device-side preparation of the COO batch already resident in `slot` (bench `value`
// loop: inputs stay in HBM, every step still does its own COO->CSR / bitmask work, one step ahead).
extern "C" int32_t dae_model_restage(dae_model* m, int32_t slot) {
    if (!m) return fail("null model");
    if (slot < 0 || slot > 1) return fail("slot must be 0 or 1");
    if (m->slots[slot].batch <= 0) return fail("slot %d holds no batch", slot);
    return prepare_slot(m, slot);
}

static int invalid_batch(int e);
int check_device_flag(dae_model* m) {
    CK(cudaStreamSynchronize(m->st2));
    CK(cudaMemcpyAsync(m->err_host, m->err, sizeof(int), cudaMemcpyDeviceToHost, m->st));
    CK(cudaStreamSynchronize(m->st));
    CK(cudaGetLastError());
    const int e = *m->err_host;
    if (e != 0) {
        cudaMemsetAsync(m->err, 0, sizeof(int), m->st);
        return invalid_batch(e);
    }
    return 0;
}

// encode forward from a staged slot (shared by train / predict / recommend)
void run_encode(dae_model* m, int slot, int bpad, int rows_pad, float kp, float kp_in, int row_offset, bool train) {
    const Slot& s = m->slots[slot];
    m->cur = slot;
    ph_begin(m, PH_ENCODE);
    EncodeArgs e{};
    e.W_enc = m->W_enc; e.b_enc = m->b_enc; e.x = s.xw; e.pub = m->pub; e.rowsum = m->rowsum; e.h = m->h;
    e.h_d = m->h_d; e.h_dT = m->h_dT; e.B = s.batch; e.bpad = rows_pad; e.H = m->H;
    // training: this rank's rows of h / h_d (and columns of h_d^T) go into every rank's copy of the global batch
    e.bcast = train ? 1 : 0;
    e.K = train ? m->world * bpad : rows_pad;
    e.row0 = train ? m->rank * bpad : 0;
    e.kp = kp; e.kp_in = kp_in;
    e.seed = m->cfg.seed; e.step = (unsigned long long)m->step; e.row_offset = row_offset;
    e.pt = m->pt;
    launch_encode_fwd(e, m->st);
    m->launches += 1;
    ph_end(m, PH_ENCODE);
}

// target bitmask of the global batch over this rank's item rows, from every rank's slot CSR
void build_ybits(dae_model* m, int slot, int B, int bpad, cudaStream_t st) {
    const Slot& s = m->slots[slot];
    if (!st) st = m->st;
    YbitsArgs y{};
    y.y = s.yw; y.ybits = m->ybits; y.n_local = m->n_local; y.ywords = m->world * bpad / 32; y.B = B; y.bpad = bpad;
    y.err = m->err; y.pt = m->pt;
    ph_begin(m, PH_YBITS, st);
    launch_ybits_shard(y, st);
    ph_end(m, PH_YBITS, st);
    m->launches += 1;
}

static AdamArgs adam_args(dae_model* m) {
    AdamArgs a{};
    a.alpha = m->cfg.lr * sqrtf(1.0f - m->b2_pow) / (1.0f - m->b1_pow);   // [TF1] ApplyAdam, fp32
    a.one_minus_b1 = 1.0f - kBeta1; a.one_minus_b2 = 1.0f - kBeta2; a.eps = kAdamEps;
    a.lambda = m->cfg.reg_lambda;
    return a;
}

// sparse-row dW_enc of the rows this rank owns, from every rank's batch
static void run_scatter(dae_model* m, int B, int bpad) {
    ScatterArgs sc{};
    sc.pub = m->pub; sc.da = m->da; sc.g_enc = m->g_enc; sc.touch_cnt = m->touch_cnt; sc.hot_list = m->hot_list; sc.n_local = m->n_local;
    sc.B = B; sc.bpad = bpad; sc.H = m->H;
    // one GPU: the gather form (fixed summation order, bit-reproducible; it runs under the decoder update anyway).
    // Several GPUs: the sparse tail is on the critical path and the fp32 red.add form is the shorter one.
    // debug bit 11 forces the atomic form, bit 12 the deterministic one.
    sc.deterministic = ((m->world == 1 && !(m->debug & 2048)) || (m->debug & 4096)) ? 1 : 0;
    sc.pt = m->pt;
    ph_begin(m, PH_SCATTER);
    launch_scatter_shard(sc, m->st);
    ph_end(m, PH_SCATTER);
    m->launches += 1 + sc.deterministic;
}

static DwArgs dw_args(dae_model* m, int bpad) {
    DwArgs w{};
    w.dzT = m->dz_all; w.h_dT = m->h_dT; w.n_local = m->n_local; w.N = m->N; w.H = m->H; w.K = m->world * bpad;
    w.pt = m->pt;
    return w;
}

// Decoder update of the rows this rank owns.  Default: ONE kernel, dW_dec tile in tensor memory + dense TF1 Adam + bf16
// operand refresh (the gradient never exists in HBM).  debug bit 2: two kernels, dW_dec through HBM (buffer "g_dec").
static void run_decoder_update(dae_model* m, int bpad, cudaStream_t st) {
    const AdamArgs a0 = adam_args(m);
    DwArgs w = dw_args(m, bpad);
    if (!(m->debug & 4)) {
        w.w = m->W_dec; w.m = m->mW_dec; w.v = m->vW_dec;
        w.g_extra = m->tied ? m->g_enc : nullptr; w.touched = m->tied ? m->touched : nullptr;
        w.adam = AdamConst{a0.alpha, a0.one_minus_b1, a0.one_minus_b2, a0.eps, a0.lambda};
        w.shadow = m->shadow;
        ph_begin(m, PH_DW, st);
        launch_dw(w, st);
        ph_end(m, PH_DW, st);
        m->launches += 1;
    } else {
        AdamArgs a = a0;
        w.g = m->g_dec;
        ph_begin(m, PH_DW, st);
        launch_dw(w, st);
        ph_end(m, PH_DW, st);
        a.w = m->W_dec; a.m = m->mW_dec; a.v = m->vW_dec; a.g = m->g_dec; a.row_touched = m->tied ? m->touched : nullptr;
        a.n = (long long)m->n_local * m->H; a.row_len = m->H;
        ph_begin(m, PH_ADAM_DEC, st);
        launch_adam_rows(a, m->tied ? m->g_enc : nullptr, m->shadow, st);
        ph_end(m, PH_ADAM_DEC, st);
        m->launches += 2;
    }
    m->full_stale = true;
}

// Forward + backward of one step up to the gradients of the batch side (da, db_enc, db_dec, cost).  Three
// cross-GPU barriers order the exchanges (none when world == 1):
//   A  the previous step is over on every rank (its Adam wrote W_enc rows this step gathers; nobody still reads the
//      h / pub / dh_sum buffers this step overwrites) and every rank's slot CSR is prepared
//   B1 every rank's rows of h_d / h_d^T / h have landed  -> decode, dh over the global batch
//   B2 every rank's dh sums, db_dec rows and cost partial are complete -> da for the global batch
extern "C" int32_t dae_model_backward_staged(dae_model* m, int32_t slot, float keep_prob, float input_keep_prob,
                                             int32_t global_batch, int32_t row_offset) {
    if (!m || !m->trainable) return fail("model is not trainable");
    if (!m->attached) return fail("world = %d but the peers are not attached (dae_model_attach_ipc)", m->world);
    if (slot < 0 || slot > 1) return fail("slot must be 0 or 1");
    const Slot& s = m->slots[slot];
    if (s.batch <= 0) return fail("slot %d holds no batch", slot);
    if (!(keep_prob > 0.f) || !(input_keep_prob > 0.f)) return fail("keep probabilities must be > 0");
    const int B = s.batch, bpad = round_up(B, 64), H = m->H, N = m->N, R = m->world;
    if (bpad > kMaxBpad) return fail("training batch %d exceeds %d", B, kMaxBpad);
    // the loss is a mean over the GLOBAL batch (DAEs.py:100); dropout is keyed by the global row
    const int gb = global_batch > 0 ? global_batch : B * R;
    if (global_batch <= 0) row_offset = m->rank * B;
    // the hidden-dropout mask of the backward (k_da_own) is keyed like the forward's: rank s's rows start at row_offset0 + s * B
    if (R > 1 && row_offset != m->rank * B)
        return fail("world = %d: row_offset must be rank * batch = %d (got %d): the ranks' rows are consecutive blocks of the global batch",
                    R, m->rank * B, row_offset);
    m->row_offset0 = row_offset - m->rank * B;
    if (m->last_bpad != 0 && m->last_bpad != bpad) {
        if (R > 1) return fail("world = %d: the padded batch size may not change between steps (%d -> %d)", R, m->last_bpad, bpad);
        // the padding columns / rows of these buffers are only ever written as zeros for ONE padded batch size: a C caller
        // that changes the batch between steps gets them re-zeroed (the Python wrappers always pass conf.batch)
        CK(cudaMemsetAsync(m->dz_all, 0, sizeof(__nv_bfloat16) * (size_t)m->n_local * R * kMaxBpad, m->st));
        CK(cudaMemsetAsync(m->h_d, 0, sizeof(__nv_bfloat16) * (size_t)R * kMaxBpad * H, m->st));
        CK(cudaMemsetAsync(m->h_dT, 0, sizeof(__nv_bfloat16) * (size_t)R * kMaxBpad * H, m->st));
        CK(cudaMemsetAsync(m->h, 0, sizeof(float) * (size_t)R * kMaxBpad * H, m->st));
        CK(cudaMemsetAsync(m->da, 0, sizeof(float) * (size_t)R * kMaxBpad * H, m->st));
    }
    m->last_batch = B; m->last_bpad = bpad;
    if (!s.has_y) return fail("slot %d was staged without targets", slot);

    CK(cudaStreamWaitEvent(m->st, s.prepared, 0));
    ph_begin(m, PH_BARRIER);
    barrier(m);                                                               // A
    ph_end(m, PH_BARRIER);
    stamp(m, m->st, TR_START);
    // whole-step calls fork the work that is off the critical path onto st3 / st4: the touched-row flags and the target
    // bitmask (next to the encode), the encoder's Adam on the rows no playlist touches (st4, co-resident with the encode,
    // the decode and dh), the decoder update (next to the sparse tail) and the bias updates (next to the encoder's Adam).
    // Not while profiling: the per-phase times are taken with every kernel running alone.
    m->par_step = m->overlap_dec && !m->profiling && !(m->debug & (1 | 2 | 8));
    const bool fork_y = m->par_step && !(m->debug & 64);
    m->bg_inflight = false;
    m->enc_early = false;
    if (fork_y) {
        CK(cudaEventRecord(m->ev_a, m->st));
        CK(cudaStreamWaitEvent(m->st3, m->ev_a, 0));
        launch_touch_shard(s.xw, m->touched, m->touch_cnt, m->hot_list, m->touched_list, B, m->pt, m->st3);
        m->launches += 1;
        // (not with an l2 term: the cost's sum of squares reads ALL of W_enc as it is BEFORE this step's update).
        // OFF by default (debug bit 10 turns it on): measured on B200 it moves 7-18 % of the encoder rows under the front
        // but costs more than it saves -- see DESIGN.md section 4 "background encoder Adam".
        if (!m->tied && m->cfg.reg_lambda == 0.f && (m->debug & 1024)) {
            // encoder rows outside `touched` have g == 0 this step: their dense Adam update starts NOW and streams through
            // HBM while the compute-bound front of the step runs; it stops when the decoder update is about to start
            CK(cudaEventRecord(m->ev_touch, m->st3));
            CK(cudaStreamWaitEvent(m->st4, m->ev_touch, 0));
            CK(cudaMemsetAsync(m->bg.ctl, 0, sizeof(unsigned int) * kBgCtlWords, m->st4));
            CK(cudaEventRecord(m->ev_pre, m->st4));
            CK(cudaStreamWaitEvent(m->st, m->ev_pre, 0));      // encode and streamer become runnable together: priority decides
            AdamArgs a = adam_args(m);
            a.w = m->W_enc; a.m = m->mW_enc; a.v = m->vW_enc; a.row_touched = m->touched;
            a.n = (long long)m->n_local * H; a.row_len = H;
            stamp(m, m->st4, TR_BG_BEGIN);
            launch_adam_bg(a, m->bg, m->st4);
            stamp(m, m->st4, TR_BG_END);
            CK(cudaEventRecord(m->ev_bg, m->st4));
            m->bg_inflight = true;
            m->launches += 1;
        }
        // Four or more GPUs: a rank owns <= 1/4 of the rows, so the unlisted-row encoder pass (HBM-bound, 1/R of its
        // single-GPU duration) is no longer than the encode (latency-bound over NVLink, growing with R): it runs NOW, next
        // to the encode, instead of behind the decoder update on the critical path.  (One or two GPUs: it would only sit
        // in front of G1's CTAs, which need whole SMs.)  Debug bit 9 keeps it behind the decoder update.
        if (R >= 4 && !m->tied && m->cfg.reg_lambda == 0.f && !m->bg_inflight && !(m->debug & 512)) {
            CK(cudaEventRecord(m->ev_touch, m->st3));
            CK(cudaStreamWaitEvent(m->st4, m->ev_touch, 0));
            AdamArgs a = adam_args(m);
            a.w = m->W_enc; a.m = m->mW_enc; a.v = m->vW_enc; a.g = nullptr; a.row_touched = m->touched;
            a.n = (long long)m->n_local * H; a.row_len = H; a.touch_mode = 1;
            launch_adam_rows(a, nullptr, nullptr, m->st4, nullptr);
            CK(cudaEventRecord(m->ev_bg, m->st4));
            m->launches += 1;
            m->enc_early = true;
        }
        build_ybits(m, slot, B, bpad, m->st3);
        stamp(m, m->st3, TR_YBITS);
        CK(cudaEventRecord(m->ev_y, m->st3));
    } else {
        launch_touch_shard(s.xw, m->touched, m->touch_cnt, m->hot_list, m->touched_list, B, m->pt, m->st);
        m->launches += 1;
    }
    run_encode(m, slot, bpad, bpad, keep_prob, input_keep_prob, row_offset, true);
    stamp(m, m->st, TR_ENCODE);
    if (fork_y) CK(cudaStreamWaitEvent(m->st, m->ev_y, 0));
    else build_ybits(m, slot, B, bpad);
    barrier(m);                                                               // B1
    CK(cudaEventRecord(s.consumed, m->st));                 // every rank has read this slot: it may be re-prepared
    if (R > 1) { launch_transpose_hd(m->h_d, m->h_dT, R * bpad, H, m->st); m->launches += 1; }

    DecodeArgs d{};
    d.W = m->shadow; d.h_d = m->h_d; d.bias = m->b_dec; d.N = N; d.H = H; d.batch = B; d.bpad = bpad;
    d.n_batch_tiles = R; d.n_local = m->n_local; d.pt = m->pt;
    d.ybits = m->ybits; d.ywords = R * bpad / 32; d.dzT = m->dz_all;
    d.db_dec = m->g_b_dec_sh; d.db_parts = m->g_b_dec_parts;
    d.loss_partial = m->loss_partial; d.inv_batch = 1.0f / (float)gb;
    if (m->debug & 8192) {
        d.trace = m->trace + 16 * (m->step & 1) + TR_G1_FIRST;
        CK(cudaMemsetAsync(d.trace, 0xff, 8, m->st));
        CK(cudaMemsetAsync(d.trace + 1, 0, 8, m->st));
    }
    ph_begin(m, PH_DECODE_LOSS);
    launch_decode_train(d, m->st);
    ph_end(m, PH_DECODE_LOSS);
    stamp(m, m->st, TR_DECODE);

    int n_sq = 0;
    const float lam = m->cfg.reg_lambda;
    if (lam != 0.f) {      // l2 term (DAEs.py:79-82, :147-150): own rows of the matrices; the replicated biases once (rank 0)
        const long long LH = (long long)m->n_local * H;
        launch_sumsq(m->W_enc, LH, m->sq_partial, kSqBlocks, m->st);
        n_sq = kSqBlocks;
        m->launches += 1;
        if (!m->tied) { launch_sumsq(m->W_dec, LH, m->sq_partial + n_sq, kSqBlocks, m->st); n_sq += kSqBlocks; m->launches += 1; }
        if (m->rank == 0) {
            launch_sumsq(m->b_dec, N, m->sq_partial + n_sq, kSqBlocks, m->st); n_sq += kSqBlocks;
            launch_sumsq(m->b_enc, H, m->sq_partial + n_sq, kSqBlocks, m->st); n_sq += kSqBlocks;
            m->launches += 2;
        }
    }
    DhArgs q{}; q.dzT = m->dz_all; q.W = m->shadow; q.partial = m->dh_partial; q.N = m->n_local; q.H = H; q.bpad = bpad;
    q.nsplit = m->nsplit; q.n_batch_tiles = R; q.ld_dz = R * bpad;
    ph_begin(m, PH_DH);
    launch_dh(q, m->st);
    if (m->bg_inflight) { launch_bg_stop(m->bg, m->st); m->launches += 1; }   // the decoder update needs the SMs and the bandwidth now
    stamp(m, m->st, TR_DH);
    // The decoder update needs only dz and h_d^T, both final now (and the l2 term above has read W_dec).  In a whole step (dae_model_train_step_staged) of an
    // untied model it starts here on its own stream and streams w / m / v through HBM while the main stream runs the
    // latency-bound tail (split-K sums, da, sparse scatter) and the encoder's Adam; apply_adam joins the two.  Not while
    // profiling: the per-phase times (bench.py roofline) are taken with the kernels running alone.
    if (m->par_step && !m->tied && !(m->debug & 128)) {
        CK(cudaEventRecord(m->ev_dh, m->st));
        CK(cudaStreamWaitEvent(m->st3, m->ev_dh, 0));
        stamp(m, m->st3, TR_DEC_BEGIN);
        run_decoder_update(m, bpad, m->st3);
        stamp(m, m->st3, TR_DEC_END);
        // The encoder rows no playlist lists have g == 0 whatever the sparse tail (da, scatter) computes: their dense Adam
        // pass follows the decoder update directly on ITS stream (two HBM-bound kernels back to back), so the latency-
        // bound tail -- slow while it shares the SMs with them -- has both kernels' duration to finish; only the pass over
        // the listed rows (a few thousand) waits for it.
        if (m->enc_early) {                     // already running since the start of the step (world >= 4): just join it
            CK(cudaStreamWaitEvent(m->st3, m->ev_bg, 0));
            m->enc_split = true;
        } else {
            AdamArgs a = adam_args(m);
            a.w = m->W_enc; a.m = m->mW_enc; a.v = m->vW_enc; a.g = nullptr; a.row_touched = m->touched;
            a.n = (long long)m->n_local * H; a.row_len = H; a.touch_mode = 1;
            if (m->bg_inflight) CK(cudaStreamWaitEvent(m->st3, m->ev_bg, 0));
            stamp(m, m->st3, TR_REST_BEGIN);
            launch_adam_rows(a, nullptr, nullptr, m->st3, m->bg_inflight ? &m->bg : nullptr);
            stamp(m, m->st3, TR_REST_END);
            m->launches += 1;
            m->enc_split = true;
        }
        CK(cudaEventRecord(m->ev_dec, m->st3));
        m->dec_inflight = true;
    }
    launch_reduce_splits(m->dh_partial, m->nsplit, bpad, H, m->pt, m->dh_sum, m->st);   // peer stores into the row owners' dh_sum
    ph_end(m, PH_DH);
    m->launches += 3 + 2;

    launch_reduce_loss2(m->loss_partial, m->n_loss_partial, m->sq_partial, n_sq, lam, 1.0f / (float)gb, m->cost_part, m->st);
    m->launches += 1;

    ph_begin(m, PH_BARRIER);
    barrier(m);                                                               // B2
    ph_end(m, PH_BARRIER);
    DaArgs da{};
    da.dh_sum = m->dh_sum; da.h = m->h; da.da = m->da; da.db_enc = m->g_b_enc; da.B = B; da.bpad = bpad; da.H = H;
    da.kp = keep_prob; da.seed = m->cfg.seed; da.step = (unsigned long long)m->step; da.row_offset0 = m->row_offset0; da.pt = m->pt;
    ph_begin(m, PH_DA);
    launch_da_own(da, m->st);                                                 // own rows, stored into every rank's da
    barrier(m);                                                               // B3: every rank's da rows have arrived
    // db_enc (column sums of da) feeds the bias update only: a whole step runs it on the bias stream (apply_adam)
    m->colsum_deferred = m->par_step && !(m->debug & 256);
    if (!m->colsum_deferred) launch_da_colsum(da, m->st);
    m->launches += 2;
    // db_dec rows from their owners; the cost is the rank-ordered sum of every rank's partial.  Only the bias update and
    // the cost read-back need them: a whole step defers both gathers to the bias stream (apply_adam), off the critical path
    m->gather_deferred = R > 1 && m->par_step && !(m->debug & 256);
    if (R > 1 && !m->gather_deferred) {
        launch_gather_items_f32(m->g_b_dec_sh, m->g_b_dec, N, m->pt, m->st);
        launch_sum_partials(m->cost_part, m->cost, 1, m->pt, m->st);
        m->launches += 2;
    }
    ph_end(m, PH_DA);
    if (m->par_step) CK(cudaEventRecord(m->ev_da, m->st));    // da and the decoder's bias gradient are final

    if (m->debug & 1) {    // parity tests: form dW_dec / dW_enc now, where they can be inspected before Adam consumes them
        DwArgs w = dw_args(m, bpad);
        w.g = m->g_dec;
        launch_dw(w, m->st);
        m->launches += 1;
    }
    if (m->debug & 2) {
        run_scatter(m, B, bpad);
        m->scatter_done = true;
    }
    return 0;
}

// dW_dec = dz^T h_d and the sparse-row dW_enc of the rows this rank owns, then the dense TF1 Adam update of every
// variable.  Purely local: no cross-GPU traffic except the (tiny) reads of the peers' published sparse inputs.
extern "C" int32_t dae_model_apply_adam(dae_model* m) {
    if (!m || !m->trainable) return fail("model is not trainable");
    if (m->last_bpad <= 0) return fail("no gradients: call dae_model_backward_staged first");
    const int H = m->H, N = m->N, B = m->last_batch, bpad = m->last_bpad;
    AdamArgs a = adam_args(m);
    if (!m->scatter_done) run_scatter(m, B, bpad);
    m->scatter_done = false;
    stamp(m, m->st, TR_TAIL);

    if (!m->dec_inflight) run_decoder_update(m, bpad, m->st);
    else {
        // join before the encoder's Adam: two HBM-bound streams running together lose bandwidth (measured: 0.950 vs
        // 0.934 ms / step), so only the latency-bound tail above (split-K sums, da, scatter) overlaps the decoder update
        CK(cudaStreamWaitEvent(m->st, m->ev_dec, 0));
        m->dec_inflight = false;
    }

    cudaStream_t sb = m->st;
    const bool fork_b = m->par_step && !(m->debug & 256);
    if (fork_b) {                   // bias updates behind the decoder update on st3, next to the encoder's Adam
        sb = m->st3;
        CK(cudaStreamWaitEvent(sb, m->ev_da, 0));
        if (m->colsum_deferred) {
            DaArgs da{};
            da.da = m->da; da.db_enc = m->g_b_enc; da.bpad = bpad; da.H = H; da.pt = m->pt;
            launch_da_colsum(da, sb);
            m->colsum_deferred = false;
        }
        if (m->gather_deferred) {
            launch_gather_items_f32(m->g_b_dec_sh, m->g_b_dec, N, m->pt, sb);
            launch_sum_partials(m->cost_part, m->cost, 1, m->pt, sb);
            m->launches += 2;
            m->gather_deferred = false;
        }
    }
    ph_begin(m, PH_ADAM_BIAS, sb);
    a.row_touched = nullptr; a.w_bf16 = nullptr; a.row_len = 1;
    a.w = m->b_enc; a.m = m->mb_enc; a.v = m->vb_enc; a.g = m->g_b_enc; a.n = H;
    launch_adam(a, sb);
    a.w = m->b_dec; a.m = m->mb_dec; a.v = m->vb_dec; a.g = m->g_b_dec; a.n = N;
    launch_adam(a, sb);
    m->launches += 2 + (N % 4 ? 1 : 0);
    ph_end(m, PH_ADAM_BIAS, sb);
    if (fork_b) CK(cudaEventRecord(m->ev_bias, sb));

    ph_begin(m, PH_ADAM_ENC);
    if (!m->tied) {   // encoder: gradient rows exist only where a batch touched them; all rows still update (dense TF1 Adam)
        a.w = m->W_enc; a.m = m->mW_enc; a.v = m->vW_enc; a.g = nullptr; a.row_touched = m->touched;
        a.n = (long long)m->n_local * H; a.row_len = H;
        if (m->enc_early && !m->enc_split) {           // (decoder update not forked: join the early pass here)
            CK(cudaStreamWaitEvent(m->st, m->ev_bg, 0));
            m->enc_split = true;
        }
        if (m->enc_split) {
            launch_adam_listed(a, m->g_enc, m->touched_list, m->st);   // the rows a playlist lists, with their gradient: all that is left
        } else {
            // what the background streamer has not done: the rows it never claimed, and every touched row (with its gradient)
            if (m->bg_inflight) CK(cudaStreamWaitEvent(m->st, m->ev_bg, 0));
            launch_adam_rows(a, m->g_enc, nullptr, m->st, m->bg_inflight ? &m->bg : nullptr);
        }
        m->bg_inflight = false; m->enc_split = false; m->enc_early = false;
        m->launches += 1;
    }
    launch_clear_listed(H, m->g_enc, m->touched, m->touch_cnt, m->touched_list, m->st);
    stamp(m, m->st, TR_STEP_END);
    m->launches += 1;
    ph_end(m, PH_ADAM_ENC);
    if (fork_b) CK(cudaStreamWaitEvent(m->st, m->ev_bias, 0));
    m->par_step = false;
    ph_collect(m);
    m->b1_pow *= kBeta1;
    m->b2_pow *= kBeta2;
    m->step += 1;
    return 0;
}

extern "C" int32_t dae_model_train_step_staged(dae_model* m, int32_t slot, float keep_prob, float input_keep_prob) {
    if (m) m->overlap_dec = true;
    const int rc = dae_model_backward_staged(m, slot, keep_prob, input_keep_prob, 0, 0);
    if (m) m->overlap_dec = false;
    if (rc) return rc;
    return dae_model_apply_adam(m);
}

extern "C" int32_t dae_model_sync_cost(dae_model* m, float* cost_out) {
    if (!m) return fail("null model");
    if (m->trainable) CK(cudaMemcpyAsync(m->cost_host, m->cost, sizeof(float), cudaMemcpyDeviceToHost, m->st));
    TRY(check_device_flag(m));
    if (cost_out) *cost_out = m->trainable ? *m->cost_host : 0.f;
    return 0;
}

extern "C" int32_t dae_model_train_flush(dae_model* m, float* cost_out, int32_t* has_cost);
extern "C" int32_t dae_model_train_step(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                        const int64_t* y_pos, const float* y_val, int64_t nnz_y, int32_t batch,
                                        float keep_prob, float input_keep_prob, float* cost_out) {
    if (m && m->async_pending) TRY(dae_model_train_flush(m, nullptr, nullptr));   // drain a pipelined step first
    TRY(dae_model_stage_batch(m, 0, x_pos, x_val, nnz_x, y_pos, y_val, nnz_y, batch));
    TRY(dae_model_train_step_staged(m, 0, keep_prob, input_keep_prob));
    return dae_model_sync_cost(m, cost_out);
}

static int invalid_batch(int e) {
    return fail("invalid sparse batch:%s%s%s", (e & kErrIndexRange) ? " index out of range" : "",
                (e & kErrRowTooLong) ? " row longer than 2048 entries" : "", (e & kErrYNotBinary) ? " y values must be 0 or 1" : "");
}

// wait for the step issued from `slot` by dae_model_train_step_async and fetch its cost / validation flag
static int collect_async(dae_model* m, int slot, float* cost_out) {
    CK(cudaEventSynchronize(m->ev_cost[slot]));
    const int e = m->err_ring[slot];
    if (e != 0) {
        cudaMemsetAsync(m->err, 0, sizeof(int), m->st);
        return invalid_batch(e);
    }
    if (cost_out) *cost_out = m->cost_ring[slot];
    return 0;
}

// Pipelined form of dae_model_train_step for a training loop: the batch is staged into the slot the device is not
// using (H2D + COO->CSR on the side stream, overlapping the step in flight), the step is enqueued, and the call
// returns the cost of the PREVIOUS step (*has_prev = 0 on the first call) without waiting for this one.
extern "C" int32_t dae_model_train_step_async(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                              const int64_t* y_pos, const float* y_val, int64_t nnz_y, int32_t batch,
                                              float keep_prob, float input_keep_prob, float* prev_cost_out,
                                              int32_t* has_prev) {
    if (!m || !m->trainable) return fail("model is not trainable");
    const int slot = m->async_slot ^ 1;
    TRY(dae_model_stage_batch(m, slot, x_pos, x_val, nnz_x, y_pos, y_val, nnz_y, batch));
    TRY(dae_model_train_step_staged(m, slot, keep_prob, input_keep_prob));
    CK(cudaMemcpyAsync(m->cost_ring + slot, m->cost, sizeof(float), cudaMemcpyDeviceToHost, m->st));
    CK(cudaMemcpyAsync(m->err_ring + slot, m->err, sizeof(int), cudaMemcpyDeviceToHost, m->st));
    CK(cudaEventRecord(m->ev_cost[slot], m->st));
    const bool had = m->async_pending;
    m->async_slot = slot;
    m->async_pending = true;
    if (has_prev) *has_prev = had ? 1 : 0;
    if (had) return collect_async(m, slot ^ 1, prev_cost_out);
    return 0;
}

// cost of the last step issued by dae_model_train_step_async (*has_cost = 0 if none is pending)
extern "C" int32_t dae_model_train_flush(dae_model* m, float* cost_out, int32_t* has_cost) {
    if (!m) return fail("null model");
    if (has_cost) *has_cost = m->async_pending ? 1 : 0;
    if (!m->async_pending) return 0;
    m->async_pending = false;
    return collect_async(m, m->async_slot, cost_out);
}

extern "C" int32_t dae_model_set_debug(dae_model* m, int32_t flags) {
    if (!m) return fail("null model");
    m->debug = flags;
    set_itemtile_pair((flags & 32768) ? 1 : 0);      // process-wide: multicast batch-tile pairs + 128-row tiles (A/B)
    set_itemtile_tune((flags >> 16) & 0xff);         // process-wide: A/B switches of the FILTER epilogue
    return 0;
}

// ------------------------------------------------------------------------------------------
// inference
// ------------------------------------------------------------------------------------------
static int ensure_scores(dae_model* m, size_t elems) {
    if (m->scores_elems >= elems) return 0;
    if (m->scores) { CK(cudaStreamSynchronize(m->st)); CK(cudaFree(m->scores)); m->scores = nullptr; }
    CK(cudaMalloc(reinterpret_cast<void**>(&m->scores), elems * sizeof(float)));
    m->scores_elems = elems;
    return 0;
}

// Batch tiles of an inference decode: up to 256 rows one tile, larger batches 256-row tiles.  Debug bit 15 (experiment,
// measured slower: 9.1 vs 6.3 ms for the cfg5 full pass): 128-row tiles in multicast pairs with a 10-stage ring.
static void infer_tiling(const dae_model* m, int B, int* bpad, int* nbt) {
    if (B <= kMaxBpad) { *bpad = round_up(B, 64); *nbt = 1; }
    else if (!(m->debug & 32768)) { *bpad = kMaxBpad; *nbt = (B + kMaxBpad - 1) / kMaxBpad; }
    else { *bpad = 128; *nbt = round_up((B + 127) / 128, 2); }
}

static int run_predict(dae_model* m, int slot, int n_cols, float* out_dev, long long ld) {
    const Slot& s = m->slots[slot];
    const int B = s.batch;
    int bpad, nbt;
    infer_tiling(m, B, &bpad, &nbt);
    if (!m->attached) return fail("world = %d but the peers are not attached (dae_model_attach_ipc)", m->world);
    if (m->world > 1 && m->full_stale) {   // inference scores every item: gather the operand rows from their owners
        launch_gather_rows_bf16(m->shadow, m->shadow_full, m->N, m->H, m->pt, m->st);   // (all ranks idle: caller's contract)
        m->launches += 1;
    }
    m->full_stale = false;
    CK(cudaStreamWaitEvent(m->st, s.prepared, 0));
    run_encode(m, slot, bpad, bpad * nbt, 1.0f, 1.0f, 0, false); // keep_prob = input_keep_prob = 1 (main_train.py:68)
    CK(cudaEventRecord(s.consumed, m->st));
    DecodeArgs d{};
    d.W = m->shadow_full; d.h_d = m->h_d; d.bias = m->b_dec; d.N = m->N; d.H = m->H; d.batch = B; d.bpad = bpad;
    d.n_batch_tiles = nbt; d.out = out_dev; d.ld_out = ld; d.n_out = n_cols;
    launch_decode_predict(d, m->st);
    m->launches += 1;
    return 0;
}

extern "C" int32_t dae_model_predict(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                     int32_t batch, int32_t n_cols, float* y_pred_out) {
    if (!m || !y_pred_out) return fail("null argument");
    if (n_cols <= 0 || n_cols > m->N) return fail("n_cols must be in (0, n_input]");
    TRY(stage_impl(m, 0, x_pos, x_val, nnz_x, nullptr, nullptr, 0, batch, false));
    TRY(ensure_scores(m, (size_t)batch * n_cols));
    TRY(run_predict(m, 0, n_cols, m->scores, n_cols));
    CK(cudaMemcpyAsync(y_pred_out, m->scores, (size_t)batch * n_cols * sizeof(float), cudaMemcpyDeviceToHost, m->st));
    return check_device_flag(m);
}

template <typename T>
static int ensure_buf(T** p, size_t* have, size_t need, cudaStream_t st) {
    if (*have >= need) return 0;
    if (*p) { CK(cudaStreamSynchronize(st)); CK(cudaFree(*p)); *p = nullptr; }
    CK(cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)));
    *have = need;
    return 0;
}

// ---- fused decode + top-K (SURVEY 8d "challenge decode + top-K": T.H.s_w bytes, not B.T scores) ----------------------
// The [B, T] score matrix is never materialised.  Three filter passes over growing prefixes of the item range
// ([lo, lo+M1) c [lo, lo+M2) c [lo, hi)): a pass appends every logit >= the playlist's threshold to its candidate list,
// an exact top-(k + max #seeds) of the list gives the threshold of the next pass.  The kp-th largest logit of a SUBSET is
// a lower bound of the kp-th largest of the whole range, so no member of the final top-k can be filtered out.  (The final
// order is on p = sigmoid(z) in fp32, where distinct logits can TIE -- near saturation, and all of z >= ~16.6 at exactly
// 1.0f -- and the lower id wins a tie: k_thr_from_topk therefore lowers every threshold by the width of the sigmoid's
// flat spot around it, and to 15 beyond that; tests/test_gpu_model.py has the saturated adversarial case.)  With
// M2 / M1 = 16 and T / M2 = 8 the lists hold ~kp x 16 and ~kp x 8 entries for exchangeable scores (popularity-ranked ids
// make the prefix bound tighter still).  A list that overflows its capacity is detected and the call falls back to the
// tiled dense path, so the result is exact in every case.  Total decode work: (M1 + M2 + T) / T = 1.13x.
constexpr int kCandCap = 16384;
constexpr int kFusedMinItems = 131072;     // smaller ranges: dense scores + k_topk

static int ensure_cand(dae_model* m, size_t rows, int kp) {
    if (m->cand_rows >= rows) return 0;
    CK(cudaStreamSynchronize(m->st));
    cudaFree(m->cand_val); cudaFree(m->cand_idx); cudaFree(m->cand_cnt); cudaFree(m->cand_thr);
    cudaFree(m->cand_tk_idx); cudaFree(m->cand_tk_score);
    m->cand_rows = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&m->cand_val), rows * kCandCap * sizeof(float)));
    CK(cudaMalloc(reinterpret_cast<void**>(&m->cand_idx), rows * kCandCap * sizeof(int)));
    CK(cudaMalloc(reinterpret_cast<void**>(&m->cand_cnt), rows * 3 * sizeof(int)));
    CK(cudaMalloc(reinterpret_cast<void**>(&m->cand_thr), rows * sizeof(float)));
    CK(cudaMalloc(reinterpret_cast<void**>(&m->cand_tk_idx), rows * 1024 * sizeof(int)));
    CK(cudaMalloc(reinterpret_cast<void**>(&m->cand_tk_score), rows * 1024 * sizeof(float)));
    if (!m->cand_cnt_host) TRY(halloc(m, &m->cand_cnt_host, 4));
    m->cand_rows = rows;
    (void)kp;
    return 0;
}

// max over rows and passes of the candidate counts -> one int for the host
__global__ void k_max_int(const int* __restrict__ v, int n, int* __restrict__ out) {
    int mx = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mx = max(mx, v[i]);
    atomicMax(out, mx);
}

// top-k of the item range [lo, hi) for the batch staged in slot 0 -> m->topk_idx / m->topk_score (global ids, sigmoid
// scores).  *overflow_out = 1 when a candidate list overflowed (the caller redoes the range densely).
static int run_recommend_fused(dae_model* m, int k, int lo, int hi, const int* sp, const int* si, int max_seeds) {
    const Slot& s = m->slots[0];
    const int B = s.batch;
    int bpad, nbt;
    infer_tiling(m, B, &bpad, &nbt);
    if (!m->attached) return fail("world = %d but the peers are not attached (dae_model_attach_ipc)", m->world);
    if (m->world > 1 && m->full_stale) {
        launch_gather_rows_bf16(m->shadow, m->shadow_full, m->N, m->H, m->pt, m->st);
        m->launches += 1;
    }
    m->full_stale = false;
    const int rows = bpad * nbt, Tn = hi - lo, kp = k + max_seeds;
    // Item shards (m->thr_exchange): every shard's threshold select asks for its share kq = ceil(kp / world) and the minimum
    // over the shards filters the next pass (exchange_min_thresholds: at least kp items of the catalogue exceed it)
    const int xw = exchange_world(m->thr_exchange);
    const int kq = (kp + xw - 1) / xw;
    int n_thr_x = 0;
    TRY(ensure_cand(m, (size_t)rows, kp));
    CK(cudaStreamWaitEvent(m->st, s.prepared, 0));
    run_encode(m, 0, bpad, rows, 1.0f, 1.0f, 0, false);
    CK(cudaEventRecord(s.consumed, m->st));

    // Pass A is dense and cheap: make its prefix as long as a candidate list (16 256 items), so that its thresholds are as
    // tight as one pass can make them.  The expected list length of a pass over n items behind a prefix of M1 is
    // kp x n / M1 (exchangeable scores; popularity-ranked ids only tighten it): when the WHOLE range fits 60 % of the list
    // capacity -- item shards of <= ~260 k tracks, i.e. 8-way sharded challenge inference -- the middle pass and its
    // select are skipped: two passes instead of three.
    const int M1 = std::min(Tn, kCandCap - 128);
    const bool two_pass = (double)kq * (double)Tn / (double)M1 <= 0.6 * kCandCap;
    // middle prefix: as short as the candidate lists of the full-range pass allow (expected length kq x Tn / M2 <= 60 % of
    // the capacity), but not below Tn / 16 -- a pass costs ~0.2 ms + 2.6 ms per million items + ~0.05 us per candidate
    // and playlist (4096 playlists), which puts the optimum near Tn / 16 for 2 M items
    const int M2min = (int)std::min((double)Tn, (double)kq * (double)Tn / (0.6 * kCandCap));
    const int M2 = two_pass ? M1 : std::min(Tn, std::max(std::max(round_up(Tn / 16, kTileItems), round_up(M2min, kTileItems)), M1));
    CK(cudaMemsetAsync(m->cand_cnt, 0, sizeof(int) * 3 * rows, m->st));
    launch_thr_from_topk(nullptr, nullptr, kp, B, rows, m->cand_thr, m->st);          // pass A keeps everything
    DecodeArgs d{};
    d.W = m->shadow_full; d.h_d = m->h_d; d.bias = m->b_dec; d.N = m->N; d.H = m->H; d.batch = B; d.bpad = bpad;
    d.n_batch_tiles = nbt; d.item0 = lo;
    d.thr = m->cand_thr; d.cand_val = m->cand_val; d.cand_idx = m->cand_idx; d.cand_cap = kCandCap;
    TopkArgs a{};
    a.scores = m->cand_val; a.ld = kCandCap; a.B = B; a.T = kCandCap; a.remap = m->cand_idx; a.idx_base = 0;
    const int stops[3] = {M1, M2, Tn};
    int prev = 0;
    if (M1 < Tn) {
        // pass A keeps everything: dense logits of the first M1 items + the dense radix select, instead of M1 list appends
        // per playlist
        TRY(ensure_scores(m, (size_t)B * M1));
        ph_begin(m, PH_REC_A);
        DecodeArgs da = d;
        da.n_out = M1; da.out = m->scores; da.ld_out = M1; da.raw_logits = 1;
        launch_decode_predict(da, m->st);
        TopkArgs ta{};
        ta.scores = m->scores; ta.ld = M1; ta.B = B; ta.T = M1; ta.k = k; ta.idx_base = 0;
        ta.seed_ptr = sp; ta.thr_div = xw;                            // per row: its share of k + its own number of seeds
        ta.thr_out = m->cand_thr;                                     // threshold only: no collection, no sort
        launch_topk(ta, m->st);
        if (xw > 1) { TRY(exchange_min_thresholds(m->thr_exchange, m->cand_thr, B, m->st)); ++n_thr_x; }
        ph_end(m, PH_REC_A);
        m->launches += 2;
        prev = M1;
    }
    for (int pass = 0; pass < 3; ++pass) {
        if (stops[pass] == prev) continue;                         // small ranges need fewer passes
        prev = stops[pass];
        d.n_out = stops[pass];
        d.cand_cnt = m->cand_cnt + pass * rows;
        const int ph = stops[pass] == Tn ? PH_REC_C : PH_REC_B;     // the full-range pass is the tensor-bound one
        ph_begin(m, ph);
        launch_decode_filter(d, m->st);
        ph_end(m, ph);
        m->launches += 1;
        a.row_n = d.cand_cnt;
        if (stops[pass] < Tn) {                                    // threshold of the next pass: kp-th largest so far
            a.k = k; a.seed_ptr = sp; a.seed_idx = nullptr; a.sigmoid_out = 0; a.thr_div = xw;
            a.thr_out = m->cand_thr;
            launch_topk(a, m->st);
            a.thr_out = nullptr;
            m->launches += 1;
            if (xw > 1) { TRY(exchange_min_thresholds(m->thr_exchange, m->cand_thr, B, m->st)); ++n_thr_x; }
        }
    }
    if (n_thr_x > kThrExchangesPerCall) return fail("internal: %d threshold exchanges in one call", n_thr_x);
    for (; xw > 1 && n_thr_x < kThrExchangesPerCall; ++n_thr_x) TRY(exchange_min_thresholds(m->thr_exchange, nullptr, 0, m->st));
    a.k = k; a.seed_ptr = sp; a.seed_idx = si; a.sigmoid_out = 1;
    a.out_idx = m->topk_idx; a.out_score = m->topk_score;
    ph_begin(m, PH_REC_SELECT);
    launch_topk(a, m->st);
    ph_end(m, PH_REC_SELECT);
    CK(cudaMemsetAsync(m->cand_tk_idx, 0, sizeof(int), m->st));
    k_max_int<<<64, 256, 0, m->st>>>(m->cand_cnt, 3 * rows, m->cand_tk_idx);
    CK(cudaMemcpyAsync(m->cand_cnt_host, m->cand_tk_idx, sizeof(int), cudaMemcpyDeviceToHost, m->st));
    m->launches += 2;
    return 0;
}

// Top-k of the catalogue range [item_lo, item_hi) (clipped to the tracks) for every playlist of the batch: global item
// ids, sigmoid scores, seeds removed (metrics.py:58-68, main_challenge.py:26-36).  The whole-catalogue call is
// dae_model_recommend; item-sharded challenge inference gives every GPU one range and merges the lists
// (dae_topk_merge_device / dp.ShardedRecommender).
extern "C" int32_t dae_model_recommend_range(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                             int32_t batch, const int32_t* seed_ptr, const int32_t* seed_idx, int32_t k,
                                             int32_t item_lo, int32_t item_hi, int32_t* out_idx, float* out_score) {
    if (!m) return fail("null argument");          // out_idx == NULL: results stay on the device ("topk_idx" / "topk_score")
    if (k <= 0 || k > 1024) return fail("k must be in [1,1024]");
    const int T = m->T;
    if (item_lo < 0) item_lo = 0;
    if (item_hi > T) item_hi = T;
    if (item_lo >= item_hi) return fail("empty item range [%d, %d)", item_lo, item_hi);
    TRY(stage_impl(m, 0, x_pos, x_val, nnz_x, nullptr, nullptr, 0, batch, false));
    size_t tk = m->topk_elems;
    TRY(ensure_buf(&m->topk_idx, &tk, (size_t)batch * k, m->st));
    tk = m->topk_elems;
    TRY(ensure_buf(&m->topk_score, &tk, (size_t)batch * k, m->st));
    m->topk_elems = tk;
    const int* sp = nullptr; const int* si = nullptr;
    int max_seeds = 0;
    if (seed_ptr) {
        const int nseed = seed_ptr[batch];
        for (int r = 0; r < batch; ++r) max_seeds = std::max(max_seeds, seed_ptr[r + 1] - seed_ptr[r]);
        TRY(ensure_buf(&m->seed_ptr, &m->seed_ptr_elems, (size_t)batch + 1, m->st));
        TRY(ensure_buf(&m->seed_idx, &m->seed_idx_elems, (size_t)(nseed > 0 ? nseed : 1), m->st));
        CK(cudaMemcpyAsync(m->seed_ptr, seed_ptr, ((size_t)batch + 1) * 4, cudaMemcpyHostToDevice, m->st));
        if (nseed > 0) CK(cudaMemcpyAsync(m->seed_idx, seed_idx, (size_t)nseed * 4, cudaMemcpyHostToDevice, m->st));
        sp = m->seed_ptr; si = m->seed_idx;
    }
    const int Tn = item_hi - item_lo;
    // debug bit 4: always fused; bit 5: never
    bool fused = ((Tn >= kFusedMinItems) || (m->debug & 16)) && !(m->debug & 32) && k + max_seeds <= 1024 && Tn >= 2 * (k + max_seeds);
    if (!fused)             // a shard on the dense path still takes part in its peers' threshold exchanges
        for (int i = 0; m->thr_exchange && i < kThrExchangesPerCall; ++i) TRY(exchange_min_thresholds(m->thr_exchange, nullptr, 0, m->st));
    if (fused) {
        TRY(run_recommend_fused(m, k, item_lo, item_hi, sp, si, max_seeds));
        CK(cudaStreamSynchronize(m->st));
        ph_collect(m);
        if (*m->cand_cnt_host > kCandCap) fused = false;           // a candidate list overflowed: exact dense fallback
    }
    if (!fused) {
        // dense path: scores of the range in row tiles that fit the scratch buffer, exact radix-select top-k per row
        if (batch > kMaxBpad && (size_t)batch * Tn > ((size_t)1 << 31))
            return fail("dense top-k of %d rows x %d items needs %zu GB of scores: reduce the batch", batch, Tn,
                        (size_t)batch * Tn * 4 >> 30);
        TRY(ensure_scores(m, (size_t)batch * T));
        TRY(run_predict(m, 0, T, m->scores, T));
        TopkArgs a{};
        a.scores = m->scores + item_lo; a.ld = T; a.B = batch; a.T = Tn; a.k = k; a.seed_ptr = sp; a.seed_idx = si;
        a.idx_base = item_lo;
        a.out_idx = m->topk_idx; a.out_score = m->topk_score;
        launch_topk(a, m->st);
        m->launches += 1;
    }
    if (out_idx) CK(cudaMemcpyAsync(out_idx, m->topk_idx, (size_t)batch * k * 4, cudaMemcpyDeviceToHost, m->st));
    if (out_score) CK(cudaMemcpyAsync(out_score, m->topk_score, (size_t)batch * k * 4, cudaMemcpyDeviceToHost, m->st));
    return check_device_flag(m);
}

// Item-sharded inference: share the filter thresholds of the fused decode + top-K with the other shards through `x` (NULL
// detaches).  From then on dae_model_recommend_range is COLLECTIVE over the exchange's ranks: same batch and k everywhere.
extern "C" int32_t dae_model_set_threshold_exchange(dae_model* m, dae_exchange* x) {
    if (!m) return fail("null argument");
    m->thr_exchange = x;
    return 0;
}

// answers CSR (host) -> device; metrics of the lists in `idx_dev` [batch, k] -> out_host [batch, 3] doubles
int run_metrics(dae_model* m, const int* idx_dev, int32_t batch, int32_t k, const int32_t* ans_ptr, const int32_t* ans_idx,
                double* out_host) {
    if (!ans_ptr || !out_host) return fail("null argument");
    const int n_ans = ans_ptr[batch];
    for (int r = 0; r < batch; ++r) {
        const int a = ans_ptr[r + 1] - ans_ptr[r];
        if (a <= 0) return fail("playlist %d has no answers: r-precision divides by len(answer) (metrics.py:26)", r);
        if (a > 2048) return fail("playlist %d has %d answers (limit 2048)", r, a);
    }
    TRY(ensure_buf(&m->ans_ptr, &m->ans_ptr_elems, (size_t)batch + 1, m->st));
    TRY(ensure_buf(&m->ans_idx, &m->ans_idx_elems, (size_t)(n_ans > 0 ? n_ans : 1), m->st));
    TRY(ensure_buf(&m->metrics, &m->metrics_elems, (size_t)batch * 3, m->st));
    CK(cudaMemcpyAsync(m->ans_ptr, ans_ptr, ((size_t)batch + 1) * 4, cudaMemcpyHostToDevice, m->st));
    if (n_ans > 0) CK(cudaMemcpyAsync(m->ans_idx, ans_idx, (size_t)n_ans * 4, cudaMemcpyHostToDevice, m->st));
    launch_metrics(idx_dev, k, batch, k, m->ans_ptr, m->ans_idx, m->metrics, m->st);
    m->launches += 1;
    CK(cudaMemcpyAsync(out_host, m->metrics, (size_t)batch * 3 * sizeof(double), cudaMemcpyDeviceToHost, m->st));
    return 0;
}

// recommend + met.single_eval on the device: only 3 doubles per playlist come back (main_train.py:62-100)
extern "C" int32_t dae_model_evaluate(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x, int32_t batch,
                                      const int32_t* seed_ptr, const int32_t* seed_idx, const int32_t* ans_ptr,
                                      const int32_t* ans_idx, int32_t k, double* metrics_out) {
    if (!m || !metrics_out) return fail("null argument");
    TRY(dae_model_recommend_range(m, x_pos, x_val, nnz_x, batch, seed_ptr, seed_idx, k, 0, m->T, nullptr, nullptr));
    TRY(run_metrics(m, m->topk_idx, batch, k, ans_ptr, ans_idx, metrics_out));
    return check_device_flag(m);
}

extern "C" int32_t dae_model_recommend(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                       int32_t batch, const int32_t* seed_ptr, const int32_t* seed_idx, int32_t k,
                                       int32_t* out_idx, float* out_score) {
    if (!m) return fail("null argument");
    return dae_model_recommend_range(m, x_pos, x_val, nnz_x, batch, seed_ptr, seed_idx, k, 0, m->T, out_idx, out_score);
}

// ------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------
extern "C" int32_t dae_model_buffer(dae_model* m, const char* name, void** dev_ptr, int64_t* n_elem,
                                    int32_t* elem_size) {
    if (!m || !name || !dev_ptr) return fail("null argument");
    const int64_t NH = (int64_t)m->N * m->H, LH = (int64_t)m->n_local * m->H;
    const Slot& sl = m->slots[m->cur];
    struct E { const char* n; void* p; int64_t c; int32_t s; };
    const int64_t KH = (int64_t)m->world * kMaxBpad * m->H;
    const E table[] = {
        {"g_dec", m->g_dec, LH, 4}, {"g_enc", m->g_enc, LH, 4}, {"g_b_enc", m->g_b_enc, m->H, 4},
        {"g_b_dec", m->g_b_dec, m->N, 4}, {"g_b_dec_sh", m->g_b_dec_sh, m->n_local, 4},
        {"touched", m->touched, m->n_local, 1}, {"cost", m->cost, 1, 4},
        {"W_enc", m->W_enc, LH, 4}, {"W_dec", m->W_dec, LH, 4}, {"W_dec_bf16", m->shadow, LH, 2},
        {"W_dec_bf16_full", m->shadow_full, NH, 2},
        {"b_enc", m->b_enc, m->H, 4}, {"b_dec", m->b_dec, m->N, 4},
        {"h", m->h, KH, 4}, {"h_d", m->h_d, KH, 2}, {"h_dT", m->h_dT, KH, 2},
        {"dzT", m->dz_all, (int64_t)m->n_local * m->world * kMaxBpad, 2},
        {"dh_partial", m->dh_partial, (int64_t)m->world * m->nsplit * kMaxBpad * m->H, 4}, {"dh_sum", m->dh_sum, KH, 4},
        {"da", m->da, KH, 4},
        {"x_row_ptr", sl.xw.row_ptr, m->Bmax + 1, 4}, {"x_row_len", sl.xw.row_len, m->Bmax, 4},
        {"x_col", sl.xw.col, m->max_nnz, 4},
        {"x_val", m->pub.xn + (m->trainable ? (size_t)m->rank * m->max_nnz : 0), m->max_nnz, 4}, {"x_rowsum", m->rowsum, m->Bmax, 4},
        {"y_row_ptr", sl.yw.row_ptr, m->Bmax + 1, 4}, {"y_row_len", sl.yw.row_len, m->Bmax, 4},
        {"y_col", sl.yw.col, m->max_nnz, 4}, {"ybits", m->ybits, (int64_t)m->n_local * m->world * (kMaxBpad / 32), 4},
        {"scores", m->scores, (int64_t)m->scores_elems, 4},
        {"topk_idx", m->topk_idx, (int64_t)m->topk_elems, 4}, {"topk_score", m->topk_score, (int64_t)m->topk_elems, 4},
        {"bg_ctl", m->bg.ctl, kBgCtlWords, 4}, {"trace", m->trace, 32, 8},
        {"cand_cnt", m->cand_cnt, (int64_t)m->cand_rows * 3, 4},
        {"mW_dec", m->mW_dec, LH, 4}, {"vW_dec", m->vW_dec, LH, 4}, {"mW_enc", m->mW_enc, LH, 4}, {"vW_enc", m->vW_enc, LH, 4},
    };
    for (const E& e : table) {
        if (strcmp(e.n, name) == 0) {
            if (!e.p) return fail("buffer '%s' is not allocated for this model", name);
            *dev_ptr = e.p;
            if (n_elem) *n_elem = e.c;
            if (elem_size) *elem_size = e.s;
            return 0;
        }
    }
    return fail("unknown buffer '%s'", name);
}

extern "C" int64_t dae_model_launch_count(dae_model* m) { return m ? m->launches : 0; }

extern "C" int32_t dae_model_set_profiling(dae_model* m, int32_t on) {
    if (!m) return fail("null model");
    if (on && !m->ph_ev[0]) for (int i = 0; i < 2 * PH_COUNT; ++i) CK(cudaEventCreate(&m->ph_ev[i]));
    m->profiling = on != 0;
    for (int k = 0; k < PH_COUNT; ++k) { m->ph_ms[k] = 0.0; m->ph_n[k] = 0; m->ph_used[k] = false; }
    return 0;
}
extern "C" int32_t dae_model_phase_count(void) { return PH_COUNT; }
extern "C" const char* dae_model_phase_name(int32_t k) { return (k >= 0 && k < PH_COUNT) ? kPhaseNames[k] : ""; }
extern "C" int32_t dae_model_phase_time(dae_model* m, int32_t k, double* total_ms, int64_t* count) {
    if (!m || k < 0 || k >= PH_COUNT) return fail("bad phase");
    if (total_ms) *total_ms = m->ph_ms[k];
    if (count) *count = m->ph_n[k];
    return 0;
}

// ------------------------------------------------------------------------------------------
// kernel-level entry points
// ------------------------------------------------------------------------------------------
extern "C" int32_t dae_topk_device(const float* scores_dev, int64_t ld, int32_t batch, int32_t n_tracks, int32_t k,
                                   const int32_t* seed_ptr_dev, const int32_t* seed_idx_dev, int32_t idx_base,
                                   int32_t* out_idx_dev, float* out_score_dev, void* stream) {
    ensure_loaded();
    if (!scores_dev || !out_idx_dev || !out_score_dev) return fail("null argument");
    if (k <= 0 || k > 1024) return fail("k must be in [1,1024]");
    TopkArgs a{};
    a.scores = scores_dev; a.ld = ld; a.B = batch; a.T = n_tracks; a.k = k; a.seed_ptr = seed_ptr_dev;
    a.seed_idx = seed_idx_dev; a.idx_base = idx_base; a.out_idx = out_idx_dev; a.out_score = out_score_dev;
    launch_topk(a, reinterpret_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
    return 0;
}

// Merge of per-shard top-k lists (item-sharded challenge inference): row r holds n (score, global id) pairs
// (shards concatenated, -1 / -inf padded); the first k by (score desc, id asc) are exactly the unsharded list.
extern "C" int32_t dae_topk_merge_device(const float* scores_dev, const int32_t* idx_dev, int32_t n, int32_t batch, int32_t k,
                                         int32_t* out_idx_dev, float* out_score_dev, void* stream) {
    ensure_loaded();
    if (!scores_dev || !idx_dev || !out_idx_dev || !out_score_dev) return fail("null argument");
    if (k <= 0 || k > 1024 || n <= 0) return fail("need 1 <= k <= 1024 and n > 0");
    TopkArgs a{};
    a.scores = scores_dev; a.ld = n; a.B = batch; a.T = n; a.k = k; a.remap = idx_dev;
    a.out_idx = out_idx_dev; a.out_score = out_score_dev;
    launch_topk(a, reinterpret_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int32_t dae_metrics_device(const int32_t* cand_dev, int64_t ld, int32_t batch, int32_t k, const int32_t* ans_ptr_dev,
                                      const int32_t* ans_idx_dev, double* out_dev, void* stream) {
    ensure_loaded();
    if (!cand_dev || !ans_ptr_dev || !ans_idx_dev || !out_dev) return fail("null argument");
    if (k <= 0 || k > 1024) return fail("k must be in [1,1024]");
    launch_metrics(cand_dev, ld, batch, k, ans_ptr_dev, ans_idx_dev, out_dev, reinterpret_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int32_t dae_adam_device(float* w_dev, float* m_dev, float* v_dev, const float* g_dev, uint16_t* w_bf16_dev,
                                   int64_t n, float lr, float beta1_power, float beta2_power, float reg_lambda,
                                   void* stream) {
    ensure_loaded();
    AdamArgs a{};
    a.w = w_dev; a.m = m_dev; a.v = v_dev; a.g = g_dev; a.w_bf16 = reinterpret_cast<__nv_bfloat16*>(w_bf16_dev);
    a.n = n; a.row_len = 1;
    a.alpha = lr * sqrtf(1.0f - beta2_power) / (1.0f - beta1_power);
    a.one_minus_b1 = 1.0f - kBeta1; a.one_minus_b2 = 1.0f - kBeta2; a.eps = kAdamEps; a.lambda = reg_lambda;
    launch_adam(a, reinterpret_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int32_t dae_coo_to_csr_device(const int64_t* pos_dev, const float* val_dev, int64_t nnz, int32_t batch,
                                         int32_t n_input, int32_t* row_ptr_dev, int32_t* row_len_dev, int32_t* col_dev,
                                         float* val_out_dev, void* stream) {
    ensure_loaded();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CsrWork w{};
    int* scratch = nullptr;
    unsigned long long* keys = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&scratch), sizeof(int) * (2 * (size_t)batch + 1)));
    CK(cudaMalloc(reinterpret_cast<void**>(&keys), sizeof(unsigned long long) * (size_t)(nnz > 0 ? nnz : 1)));
    CK(cudaMemsetAsync(scratch, 0, sizeof(int) * (2 * (size_t)batch + 1), st));
    w.cnt = scratch; w.cursor = scratch + batch; int* err = scratch + 2 * batch;
    w.row_ptr = row_ptr_dev; w.keys = keys; w.row_len = row_len_dev; w.col = col_dev; w.val = val_out_dev;
    launch_coo_to_csr(reinterpret_cast<const long long*>(pos_dev), val_dev, (int)nnz, batch, n_input, w, err, st);
    int herr = 0;
    CK(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(scratch); cudaFree(keys);
    CK(cudaGetLastError());
    if (herr) return fail("invalid sparse batch (flag %d)", herr);
    return 0;
}

extern "C" int32_t dae_dh_nsplit(int32_t n_items) { return dh_nsplit(n_items, 1); }

extern "C" int32_t dae_gemm_test_device(int32_t op, const uint16_t* a_dev, const uint16_t* b_dev,
                                        const float* bias_dev, float* out_dev, int32_t n_items, int32_t n_hidden,
                                        int32_t batch, int32_t bpad, int32_t lbo, int32_t sbo, int32_t* nsplit_out,
                                        void* stream) {
    ensure_loaded();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n_hidden % 64 || n_hidden > 256 || bpad % 64 || bpad > 256) return fail("bad shape");
    const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(a_dev);
    const __nv_bfloat16* Bm = reinterpret_cast<const __nv_bfloat16*>(b_dev);
    if (op == 0) {
        DecodeArgs d{};
        d.W = A; d.h_d = Bm; d.bias = bias_dev; d.N = n_items; d.H = n_hidden; d.batch = batch; d.bpad = bpad;
        d.n_batch_tiles = (batch + bpad - 1) / bpad; d.out = out_dev; d.ld_out = n_items; d.n_out = n_items;
        d.pt.world = 1;
        launch_decode_predict(d, st);
    } else if (op == 1) {
        DwArgs w{}; w.dzT = A; w.h_dT = Bm; w.g = out_dev; w.n_local = n_items; w.N = n_items; w.H = n_hidden; w.K = bpad;
        w.pt.world = 1;
        launch_dw(w, st);
    } else if (op == 2) {
        DhArgs q{}; q.dzT = A; q.W = Bm; q.partial = out_dev; q.N = n_items; q.H = n_hidden; q.bpad = bpad;
        q.nsplit = dh_nsplit(n_items, 1); q.lbo = lbo; q.sbo = sbo;
        if (nsplit_out) *nsplit_out = q.nsplit;
        launch_dh(q, st);
    } else {
        return fail("unknown op %d", op);
    }
    CK(cudaGetLastError());
    return 0;
}
