// C-ABI of the item-sharded challenge inference exchange (include/dae_b200.h, "item-sharded inference"): every rank ranks
// its slice of the track catalogue (dae_model_recommend_range) and the per-shard top-k lists are merged into the
// unsharded list (main_challenge.py:80-90 on one GPU upstream).  No collective library: a rank's lists are
// STORED straight into every peer's merge buffer over NVLink (CUDA IPC mapping, plain 16-byte stores), one flag barrier
// orders them, and the (score desc, id asc) merge of world x k candidates per playlist runs locally (k_topk, candidate mode).
#include "model.h"

struct dae_exchange {
    int device = 0, world = 1, rank = 0, max_batch = 0, max_k = 0;
    char* base = nullptr;
    size_t bytes = 0;
    PeerTable pt{};
    void* ipc_opened[kMaxWorld] = {};
    unsigned int* flags = nullptr;
    int* cat_idx[2] = {nullptr, nullptr};       // [max_batch, world * max_k] per call parity
    float* cat_score[2] = {nullptr, nullptr};
    int *out_idx = nullptr;
    float* out_score = nullptr;
    float* thr_buf[2] = {nullptr, nullptr};     // [world, max_batch] per threshold-exchange parity
    long long thr_calls = 0;
    unsigned int epoch = 0;
    cudaStream_t st = nullptr;                  // private stream (a NULL `stream` argument): ranks sharing ONE process must not
                                                // share a stream -- a rank's barrier kernel would sit in front of its peers' stores
    long long calls = 0, launches = 0;
    bool attached = false;
    bool profiling = false;
    cudaEvent_t ev[5] = {};                     // profiling: start, stores done, barrier passed, merged, copied back
    float phase_ms[4] = {0.f, 0.f, 0.f, 0.f};
    int phase_calls = 0;
};

// this rank's [batch, k] lists -> columns [rank * k, rank * k + k) of the [batch, world * k] merge buffer of EVERY rank
// (rows_per_rank == 0) or of the rank that merges the row (row / rows_per_rank)
__global__ void k_exchange_store(const int* __restrict__ idx, const float* __restrict__ score, int batch, int k, int* cat_idx,
                                 float* cat_score, int rows_per_rank, const __grid_constant__ PeerTable pt) {
    const int row = blockIdx.x, dst = rows_per_rank ? row / rows_per_rank : blockIdx.y;
    int* di = peer_ptr(pt, dst, cat_idx) + ((size_t)row * pt.world + pt.rank) * k;
    float* ds = peer_ptr(pt, dst, cat_score) + ((size_t)row * pt.world + pt.rank) * k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        di[i] = idx[(size_t)row * k + i];
        ds[i] = score[(size_t)row * k + i];
    }
}

// Filter thresholds shared across the item shards (fused decode + top-K, api.cu run_recommend_fused).  Every shard takes the
// ceil(kp / world)-th largest logit of what it has seen as ITS threshold; the minimum over the shards is then exceeded by
// at least kp items of the whole catalogue, so it is a valid filter for every shard -- and about `world` times tighter than
// each shard's own kp-th largest: the candidate lists of the next pass shrink by `world`.
__global__ void k_thr_store(const float* __restrict__ thr, int n, float* thr_buf, int ld, const __grid_constant__ PeerTable pt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t = thr[i];
    for (int d = 0; d < pt.world; ++d) peer_ptr(pt, d, thr_buf)[(size_t)pt.rank * ld + i] = t;
}
__global__ void k_thr_min(float* __restrict__ thr, int n, const float* __restrict__ thr_buf, int ld, int world) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t = thr_buf[i];
    for (int q = 1; q < world; ++q) t = fminf(t, thr_buf[(size_t)q * ld + i]);
    thr[i] = t;
}
int exchange_world(const dae_exchange* x) { return x ? x->world : 1; }
// Collective on `st` (every rank, the same number of times per call): thr[0..n) <- min over the ranks; thr == nullptr only
// keeps the barrier count in step (a rank whose call took a path without thresholds).
int exchange_min_thresholds(dae_exchange* x, float* thr, int n, cudaStream_t st) {
    if (!x || x->world == 1) return 0;
    if (!x->attached) return fail("world = %d but the peers are not attached (dae_exchange_attach_ipc)", x->world);
    if (n > x->max_batch) return fail("threshold exchange of %d rows exceeds the exchange's capacity %d", n, x->max_batch);
    const int p = (int)(x->thr_calls & 1);
    x->thr_calls += 1;
    if (thr) k_thr_store<<<(n + 255) / 256, 256, 0, st>>>(thr, n, x->thr_buf[p], x->max_batch, x->pt);
    x->epoch += 1;
    launch_barrier(x->flags, x->epoch, x->pt, st);
    if (thr) k_thr_min<<<(n + 255) / 256, 256, 0, st>>>(thr, n, x->thr_buf[p], x->max_batch, x->world);
    x->launches += thr ? 3 : 1;
    return 0;
}

extern "C" int32_t dae_exchange_create(int32_t device, int32_t world, int32_t rank, int32_t max_batch, int32_t max_k,
                                       dae_exchange** out) {
    if (!out) return fail("null argument");
    *out = nullptr;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail("need 1 <= world <= %d and 0 <= rank < world", kMaxWorld);
    if (max_batch <= 0 || max_k <= 0 || max_k > 1024) return fail("need max_batch > 0 and 0 < max_k <= 1024");
    CK(cudaSetDevice(device));
    ensure_loaded();
    dae_exchange* x = new dae_exchange();
    x->device = device; x->world = world; x->rank = rank; x->max_batch = max_batch; x->max_k = max_k;
    Arena A;
    for (int pass = 0; pass < 2; ++pass) {
        A.off = 0;
        x->flags = A.take<unsigned int>(kMaxWorld);
        for (int p = 0; p < 2; ++p) {
            x->cat_idx[p] = A.take<int>((size_t)max_batch * world * max_k);
            x->cat_score[p] = A.take<float>((size_t)max_batch * world * max_k);
        }
        x->out_idx = A.take<int>((size_t)max_batch * max_k);
        x->out_score = A.take<float>((size_t)max_batch * max_k);
        for (int p = 0; p < 2; ++p) x->thr_buf[p] = A.take<float>((size_t)world * max_batch);
        if (pass == 0) {
            x->bytes = (A.off + 1023) & ~size_t(1023);
            CK(cudaMalloc(reinterpret_cast<void**>(&x->base), x->bytes));
            CK(cudaMemset(x->base, 0, x->bytes));
            A.base = x->base;
        }
    }
    CK(cudaStreamCreateWithFlags(&x->st, cudaStreamNonBlocking));
    x->pt.world = world; x->pt.rank = rank; x->pt.base[rank] = x->base;
    x->attached = world == 1;
    cudaFuncAttributes a;
    PRELOAD_KERNEL(k_exchange_store);
    PRELOAD_KERNEL(k_thr_store);
    PRELOAD_KERNEL(k_thr_min);
    *out = x;
    return 0;
}

extern "C" void dae_exchange_destroy(dae_exchange* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < kMaxWorld; ++r) if (x->ipc_opened[r]) cudaIpcCloseMemHandle(x->ipc_opened[r]);
    cudaFree(x->base);
    for (int i = 0; i < 5; ++i) if (x->ev[i]) cudaEventDestroy(x->ev[i]);
    if (x->st) cudaStreamDestroy(x->st);
    delete x;
}

extern "C" int32_t dae_exchange_ipc_handle(dae_exchange* x, void* handle_out64) {
    if (!x || !handle_out64) return fail("null argument");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, x->base));
    memcpy(handle_out64, &h, 64);
    return 0;
}

extern "C" int32_t dae_exchange_attach_ipc(dae_exchange* x, const void* handles, int32_t n_handles) {
    if (!x || !handles) return fail("null argument");
    if (n_handles != x->world) return fail("expected %d IPC handles, got %d", x->world, n_handles);
    CK(cudaSetDevice(x->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * r, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->ipc_opened[r] = p;
        x->pt.base[r] = static_cast<char*>(p);
    }
    x->attached = true;
    return 0;
}

extern "C" int32_t dae_exchange_attach_local(dae_exchange* x, dae_exchange* const* peers, int32_t n_peers) {
    if (!x || !peers) return fail("null argument");
    if (n_peers != x->world) return fail("expected %d peers, got %d", x->world, n_peers);
    for (int r = 0; r < x->world; ++r) {
        if (!peers[r] || peers[r]->world != x->world || peers[r]->rank != r || peers[r]->bytes != x->bytes)
            return fail("peer %d does not match this exchange's layout", r);
        if (peers[r]->device != x->device) {
            CK(cudaSetDevice(x->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            (void)cudaGetLastError();
        }
        x->pt.base[r] = peers[r]->base;
    }
    x->attached = true;
    return 0;
}

// Every rank calls this with its own per-shard lists (device pointers, [batch, k], global ids, -1 / -inf padded): the lists
// are stored into every rank's merge buffer, one flag barrier, then the local merge.  The merged lists (identical on every
// rank, exactly the unsharded ranking) stay in the exchange's device buffers and are copied to out_idx / out_score (HOST,
// either may be NULL).  Call parity double-buffers the merge buffer, so one barrier per call is enough: a rank can only
// overwrite the buffer of call n-2, which every rank has finished merging before it could signal the barrier of call n-1.
static int32_t merge_impl(dae_exchange* x, const int32_t* idx_dev, const float* score_dev, int32_t batch, int32_t k,
                          bool own_rows, int32_t* row_begin, int32_t* row_end, int32_t* out_idx, float* out_score, void* stream) {
    if (!x || !idx_dev || !score_dev) return fail("null argument");
    if (!x->attached) return fail("world = %d but the peers are not attached (dae_exchange_attach_ipc)", x->world);
    if (batch <= 0 || batch > x->max_batch || k <= 0 || k > x->max_k) return fail("batch / k outside the exchange's capacity");
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : x->st;
    const int p = (int)(x->calls & 1);
    x->calls += 1;
    // own_rows: rank r merges (and returns) rows [r * rpr, (r + 1) * rpr) only -- 1 / world of the stores, the merge and the
    // read-back; otherwise every rank merges every row
    const int rpr = own_rows ? (batch + x->world - 1) / x->world : 0;
    const int r0 = own_rows ? (x->rank * rpr < batch ? x->rank * rpr : batch) : 0;
    const int r1 = own_rows ? (r0 + rpr < batch ? r0 + rpr : batch) : batch;
    if (row_begin) *row_begin = r0;
    if (row_end) *row_end = r1;
    const bool prof = x->profiling;
    if (prof) CK(cudaEventRecord(x->ev[0], st));
    k_exchange_store<<<dim3(batch, own_rows ? 1 : x->world), 128, 0, st>>>(idx_dev, score_dev, batch, k, x->cat_idx[p], x->cat_score[p],
                                                                         rpr, x->pt);
    if (prof) CK(cudaEventRecord(x->ev[1], st));
    if (x->world > 1) {
        x->epoch += 1;
        launch_barrier(x->flags, x->epoch, x->pt, st);
    }
    if (prof) CK(cudaEventRecord(x->ev[2], st));
    const size_t ld = (size_t)x->world * k;
    if (r1 > r0) {
        TopkArgs a{};
        a.scores = x->cat_score[p] + r0 * ld; a.ld = (long long)ld; a.B = r1 - r0; a.T = x->world * k; a.k = k;
        a.remap = x->cat_idx[p] + r0 * ld;
        a.out_idx = x->out_idx + (size_t)r0 * k; a.out_score = x->out_score + (size_t)r0 * k;
        launch_topk(a, st);
    }
    if (prof) CK(cudaEventRecord(x->ev[3], st));
    x->launches += x->world > 1 ? 3 : 2;
    const size_t off = (size_t)r0 * k, n = (size_t)(r1 - r0) * k;
    if (out_idx && n) CK(cudaMemcpyAsync(out_idx + off, x->out_idx + off, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (out_score && n) CK(cudaMemcpyAsync(out_score + off, x->out_score + off, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (prof) CK(cudaEventRecord(x->ev[4], st));
    if (out_idx || out_score || prof) CK(cudaStreamSynchronize(st));
    if (prof) {
        for (int i = 0; i < 4; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, x->ev[i], x->ev[i + 1]));
            x->phase_ms[i] += ms;
        }
        x->phase_calls += 1;
    }
    CK(cudaGetLastError());
    return 0;
}

// Every rank calls this with its own per-shard lists (device pointers, [batch, k], global ids, -1 / -inf padded): the lists
// are stored into every rank's merge buffer, one flag barrier, then the local merge.  The merged lists (identical on every
// rank, exactly the unsharded ranking) stay in the exchange's device buffers and are copied to out_idx / out_score (HOST,
// either may be NULL).  Call parity double-buffers the merge buffer, so one barrier per call is enough: a rank can only
// overwrite the buffer of call n-2, which every rank has finished merging before it could signal the barrier of call n-1.
extern "C" int32_t dae_exchange_merge_topk(dae_exchange* x, const int32_t* idx_dev, const float* score_dev, int32_t batch,
                                           int32_t k, int32_t* out_idx, float* out_score, void* stream) {
    return merge_impl(x, idx_dev, score_dev, batch, k, false, nullptr, nullptr, out_idx, out_score, stream);
}

// The same exchange with the merge itself sharded by playlist: rank r receives, merges and returns rows
// [*row_begin, *row_end) = its 1 / world of the batch (out_idx / out_score are still [batch, k] host arrays; only those rows
// are written).  Per rank (world - 1) / world of ONE list set crosses NVLink instead of world - 1 copies of it.
extern "C" int32_t dae_exchange_merge_topk_rows(dae_exchange* x, const int32_t* idx_dev, const float* score_dev, int32_t batch,
                                                int32_t k, int32_t* row_begin, int32_t* row_end, int32_t* out_idx,
                                                float* out_score, void* stream) {
    return merge_impl(x, idx_dev, score_dev, batch, k, true, row_begin, row_end, out_idx, out_score, stream);
}

// per-call device times of the exchange (ms, averaged since profiling was switched on): stores, barrier, merge, read-back
extern "C" int32_t dae_exchange_set_profiling(dae_exchange* x, int32_t on) {
    if (!x) return fail("null argument");
    CK(cudaSetDevice(x->device));
    if (on && !x->ev[0]) for (int i = 0; i < 5; ++i) CK(cudaEventCreate(&x->ev[i]));
    x->profiling = on != 0;
    for (int i = 0; i < 4; ++i) x->phase_ms[i] = 0.f;
    x->phase_calls = 0;
    return 0;
}
extern "C" int32_t dae_exchange_phase_ms(dae_exchange* x, float* out4) {
    if (!x || !out4) return fail("null argument");
    for (int i = 0; i < 4; ++i) out4[i] = x->phase_calls ? x->phase_ms[i] / (float)x->phase_calls : 0.f;
    return 0;
}

extern "C" int64_t dae_exchange_launch_count(dae_exchange* x) { return x ? x->launches : 0; }
