// Internal launcher interface between api.cu (the C-ABI) and the kernel translation units.
// All pointers are DEVICE pointers unless noted; every launcher enqueues on `st` and never syncs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "adam.cuh"

namespace dae {

constexpr int kTileItems = 128;    // UMMA M: catalogue items per accumulator tile
constexpr int kMaxBpad = 256;      // UMMA N limit: batch columns per tile
constexpr int kMaxRowNnz = 2048;   // per-playlist COO entries the row sorter holds in smem
constexpr int kMaxWorld = 8;       // GPUs of one NVSwitch box
constexpr float kEpsLog = 1e-10f;  // DAEs.py:42,98-99
constexpr float kNegWeight = 0.55f;  // DAEs.py:99

// error flag bits written by device-side validation (prepare kernels)
enum : int { kErrIndexRange = 1, kErrRowTooLong = 2, kErrYNotBinary = 4 };

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- multi-GPU layout ------------------------------------------------------------------------
// Every rank allocates the same arena layout, so a buffer of rank r is `base[r] + (local - base[rank])`.
// Peers are mapped with CUDA IPC (one process per GPU) and addressed by plain loads / stores over
// NVLink / NVSwitch.  Catalogue rows (W_enc, W_dec, their Adam moments, gradients, the bf16 decoder
// operand, the dz tiles) are owned tile-cyclically: 128-item tile t lives on rank t % world as local
// tile t / world, which spreads the popularity-ranked head of the catalogue (Zipf ids)
// evenly over the GPUs.  The batch is sharded by playlist for the sparse side (encode, dW_enc) and
// all-gathered (K = world x batch-tile rows of h_d, 128 KB each) for the dense side: every rank
// decodes ITS item rows against the GLOBAL batch, so dz, dW_dec and the Adam update never leave
// the GPU that owns the rows; only h_d / h rows, split-K sums of dh and the sparse inputs cross NVLink.
struct PeerTable {
    char* base[kMaxWorld];
    int world, rank;
};
template <typename T>
__host__ __device__ __forceinline__ T* peer_ptr(const PeerTable& pt, int r, T* local) {
    return reinterpret_cast<T*>(pt.base[r] + (reinterpret_cast<const char*>(local) - pt.base[pt.rank]));
}
__host__ __device__ __forceinline__ int item_owner(int item, int world) { return (item >> 7) % world; }
__host__ __device__ __forceinline__ int item_local(int item, int world) {
    return (((item >> 7) / world) << 7) | (item & 127);
}
__host__ __device__ __forceinline__ int item_global(int local, int world, int rank) {
    return ((((local >> 7) * world) + rank) << 7) | (local & 127);
}

// ---- sparse.cu ---------------------------------------------------------------------------
struct CsrWork {        // COO -> per-row sorted, de-duplicated (last occurrence wins) CSR-with-gaps
    int* cnt;           // [B]   zeroed by the launcher
    int* row_ptr;       // [B+1] raw (pre-dedup) offsets
    int* cursor;        // [B]
    unsigned long long* keys;  // [max_nnz]  (col << 32 | entry)
    int* row_len;       // [B]   entries kept per row
    int* col;           // [max_nnz] at row_ptr[r] .. row_ptr[r]+row_len[r]
    float* val;         // [max_nnz] value of the last occurrence
};
void launch_coo_to_csr(const long long* pos, const float* val, int nnz, int B, int N, CsrWork w, int* err,
                       cudaStream_t st);

// The normalised input of the step in flight, published for the sparse-row scatter of every rank
// (same offsets as the slot's CSR): x_n = x_d / (s + 1e-10) per kept entry.  Training keeps the WHOLE global
// batch on every rank: segment s (seg_rows rows, seg_nnz entries) is written by rank s's encode into every copy.
struct PubInput {
    int* row_ptr;       // [world][seg_rows]
    int* row_len;       // [world][seg_rows]
    int* col;           // [world][seg_nnz]
    float* xn;          // [world][seg_nnz]
    int seg_rows, seg_nnz;
};

struct EncodeArgs {
    const float* W_enc;   // this rank's rows of the fp32 master, local-tile order
    const float* b_enc;   // [H]
    CsrWork x;            // read only
    PubInput pub;         // written: col, x_n
    float* rowsum;        // [B] s = sum_j x_d (DAEs.py:41)
    float* h;             // [rows, H] fp32 sigma(a)
    __nv_bfloat16* h_d;   // [rows, H] dropout(h), bf16, padding rows zero
    __nv_bfloat16* h_dT;  // [H, K]
    int B, bpad, H, K;    // K: columns of h_dT (train: world * bpad)
    int row0;             // first row / h_dT column of this rank's block (train: rank * bpad)
    int bcast;            // 1: store h / h_d rows and h_dT columns into EVERY rank's copy (training), 0: local only
    float kp, kp_in;
    unsigned long long seed, step;
    int row_offset;       // global row index of local row 0 for the dropout keys (rank * B)
    PeerTable pt;
};
void launch_encode_fwd(const EncodeArgs& a, cudaStream_t st);
// several ranks: the encode does not store h_d^T remotely; every rank transposes its copy of the global h_d [K, H] after B1
void launch_transpose_hd(const __nv_bfloat16* h_d, __nv_bfloat16* h_dT, int K, int H, cudaStream_t st);

// y of the GLOBAL batch restricted to the item rows this rank owns, as a bitmask [local rows, K/32 words]:
// bit (s * bpad + i) of row item_local(c) is set when row i of rank s's target CSR contains column c.
struct YbitsArgs {
    CsrWork y;                // this rank's slot CSR (local pointers; peers' through pt)
    uint32_t* ybits;          // [local rows, ywords], zeroed by the launcher
    int n_local, ywords, B, bpad;
    int* err;
    PeerTable pt;
};
void launch_ybits_shard(const YbitsArgs& a, cudaStream_t st);

// dh of the global batch over this rank's item rows: fixed-order sum of the split-K partials; the rows of rank s go into
// slot [this rank] of rank s's dh_sum [world][bpad][H] (peer stores)
void launch_reduce_splits(const float* partial, int nsplit, int bpad, int H, const PeerTable& pt, float* dh_sum, cudaStream_t st);

struct DaArgs {               // da = dh * (keep/kp) * h(1-h): each rank forms its own rows and stores them into every rank's da
    const float* dh_sum;      // [world][bpad][H] every rank's item-shard contribution to THIS rank's rows, summed in rank order
    const float* h;           // [K, H] fp32, only this rank's rows are filled
    float* da;                // [K, H] of the global batch, complete after the barrier that follows launch_da_own
    float* db_enc;            // [H] column sums over the global batch
    int B, bpad, H;           // rows per rank, padded rows per rank
    float kp;
    unsigned long long seed, step;
    int row_offset0;          // global row of rank 0's first playlist in the dropout keys (the forward uses row_offset0 + rank * B)
    PeerTable pt;
};
void launch_da_own(const DaArgs& a, cudaStream_t st);
void launch_da_colsum(const DaArgs& a, cudaStream_t st);   // db_enc, once da is complete

struct ScatterArgs {          // dW_enc rows owned by this rank, from the published input of the global batch and da (both local)
    PubInput pub;
    const float* da;          // [K, H]
    float* g_enc;             // [local rows, H], zero outside the touched rows
    const int* touch_cnt;     // [local rows] playlists of the global batch listing the row (launch_touch_shard)
    const int* hot_list;      // [1 + local rows] count, then the rows listed by three or more playlists
    int n_local, B, bpad, H;
    int deterministic;        // 1: rows with > 2 contributors are gathered in a fixed order (bit-reproducible), 0: fp32 red.add only
    PeerTable pt;
};
void launch_scatter_shard(const ScatterArgs& a, cudaStream_t st);
// touched[local row] = 1 for every catalogue row this rank owns that occurs in the input CSR of any rank's slot
// hot_list / touched_list: [1 + local rows]: count, then the rows listed by >= 3 playlists / by any playlist (arbitrary order)
void launch_touch_shard(const CsrWork& x, unsigned char* touched, int* touch_cnt, int* hot_list, int* touched_list, int B,
                        const PeerTable& pt, cudaStream_t st);

// out[global item] = src_of_owner[local item] for every catalogue item (fp32 vector / bf16 rows of H)
void launch_gather_items_f32(const float* src_local, float* out, int N, const PeerTable& pt, cudaStream_t st);
void launch_gather_rows_bf16(const __nv_bfloat16* src_local, __nv_bfloat16* out, int N, int H, const PeerTable& pt,
                             cudaStream_t st);

// out[i] = sum over ranks (fixed order) of part[i] read from every rank's arena
void launch_sum_partials(const float* part_local, float* out, int n, const PeerTable& pt, cudaStream_t st);
// all ranks have reached `epoch` on their streams, and everything they wrote before is visible
void launch_barrier(unsigned int* flags_local, unsigned int epoch, const PeerTable& pt, cudaStream_t st);

// ---- gemm_sm100.cu -------------------------------------------------------------------------
struct DecodeArgs {
    const __nv_bfloat16* W;    // bf16 decoder operand: train [local rows, H] (this rank's items), predict [N, H]
    const __nv_bfloat16* h_d;  // [n_batch_tiles * bpad, H] bf16
    const float* bias;         // [N] (global item order)
    int N, H;                  // catalogue size (global)
    int batch;                 // valid rows per batch tile
    int bpad;                  // rows per batch tile (multiple of 64, <= 256)
    int n_batch_tiles;         // train: world (the global batch), predict: ceil(batch / 256)
    // train
    int n_local;               // item rows held by this rank (multiple of 128)
    PeerTable pt;              // world / rank: local item -> catalogue id
    const uint32_t* ybits;     // [local rows, ywords]
    int ywords;
    __nv_bfloat16* dzT;        // [local rows, n_batch_tiles * bpad]
    float* db_dec;             // [local rows]
    float* db_parts;           // [4 * n_batch_tiles, local rows] workspace: the epilogue warps' shares of db_dec
    unsigned long long* trace; // debug: [2] %globaltimer of the first / last CTA start (nullptr: off)
    float* loss_partial;       // [grid.x * grid.y]
    float inv_batch;
    // predict
    float* out;                // [batch, ld_out] fp32 scores
    long long ld_out;
    int n_out;                 // columns written (n_tracks or N)
    const float* mix_wp;       // [batch] or nullptr: y_pred = title*w_t + p*w_p (DAEs.py:180)
    const float* mix_wt;
    const float* title_score;  // [batch, ld_out] or nullptr
    int raw_logits;            // predict: write the logits z instead of sigmoid(z) (no title mix)
    // predict / filter (fused decode + top-K): rows [item0, item0 + n_out) of W / bias are scored
    int item0;
    const float* thr;          // [n_batch_tiles * bpad] per-playlist logit threshold (+inf: emit nothing)
    float* cand_val;           // [n_batch_tiles * bpad, cand_cap]
    int* cand_idx;
    int* cand_cnt;             // [n_batch_tiles * bpad]
    int cand_cap;
};
int decode_grid(int N, int n_batch_tiles);
void launch_decode_train(const DecodeArgs& a, cudaStream_t st);    // G1: z, loss, dz, db_dec
void set_itemtile_tune(int bits); // A/B switches of the FILTER epilogue (dae_model_set_debug bits 16..)
void set_itemtile_pair(int on);   // PREDICT / FILTER: clusters of two batch tiles share each W chunk (TMA multicast)
void launch_decode_predict(const DecodeArgs& a, cudaStream_t st);  // G1: z, sigmoid, scores
void launch_decode_filter(const DecodeArgs& a, cudaStream_t st);   // G1f: z > threshold -> candidate lists

struct DwArgs {
    const __nv_bfloat16* dzT;   // [local rows, K] d cost/dz of the item tiles this rank owns, all ranks' batch columns
    const __nv_bfloat16* h_dT;  // [H, K]
    float* g;                   // [local rows, H] fp32 raw dW_dec, overwritten (nullptr: not materialised)
    int n_local;                // local rows (multiple of 128)
    int N, H, K;                // N: catalogue size (global), bounds the last tile
    // fused dense TF1 Adam on the tile while it is still in TMEM (w == nullptr: no update)
    float* w; float* m; float* v;       // [local rows, H] fp32 master + moments
    const float* g_extra;               // tied model: sparse-row dW_enc [local rows, H], added where touched[row] != 0
    const unsigned char* touched;       // [local rows]
    AdamConst adam;
    __nv_bfloat16* shadow;              // bf16 operand copy of the same rows (same indexing as w), or nullptr
    PeerTable pt;                       // world / rank: which catalogue rows the local tiles are
    int ld, col0;                       // row stride (0: H) and first column of g / w / m / v (title head: column blocks of [N, 512])
    int shadow_ld, shadow_col0;         // fused path: row stride (0: H) and first column of shadow (a column block of the title operand)
};
void launch_dw(const DwArgs& a, cudaStream_t st);                  // G2: dW_dec = dz^T . h_d (+ Adam)

struct DhArgs {
    const __nv_bfloat16* dzT;   // [rows, ld_dz]: batch tile bt is columns [bt * bpad, +bpad)
    const __nv_bfloat16* W;     // [rows, H] (row stride ldW)
    float* partial;             // [n_batch_tiles][nsplit][bpad][H]
    int N;                      // rows contracted over
    int H, bpad, nsplit;        // nsplit: split-K CTAs per batch tile
    int n_batch_tiles, ld_dz;   // 0 -> 1 tile, ld_dz = bpad
    int lbo, sbo;               // MN-major descriptor strides (bytes); 0 -> defaults
    int ldW;                    // row stride of W in elements (0: H)
};
int dh_nsplit(int N, int n_batch_tiles = 1);   // split-K CTAs per batch tile
void launch_dh(const DhArgs& a, cudaStream_t st);                  // G3: dh = dz . W_dec (split-K)

// ---- optim.cu -------------------------------------------------------------------------------
struct AdamArgs {
    float* w; float* m; float* v;
    const float* g;
    __nv_bfloat16* w_bf16;            // shadow or nullptr
    const unsigned char* row_touched; // nullptr: g dense; else g row read only where touched[row] != 0
    long long n;                      // elements
    int row_len;                      // H for matrices (row = idx / row_len), 1 for vectors
    float alpha, one_minus_b1, one_minus_b2, eps, lambda;
    int touch_mode;                   // launch_adam_rows: 0 every row, 1 only rows with row_touched == 0, 2 only rows with row_touched != 0
};
void launch_adam(const AdamArgs& a, cudaStream_t st);
// matrices: a.w/m/v are [rows, row_len] (row_len % 4 == 0); gradient = a.g (dense, or nullptr) + g_sparse rows where
// a.row_touched != 0 (or nullptr); shadow: bf16 copy of the same rows (same indexing), or nullptr
// background streamer of the encoder's untouched rows (optim.cu: k_adam_bg): control words in device memory
struct BgAdam {
    unsigned int* ctl;                // [kBgCtlWords], zeroed before every launch of launch_adam_bg
};
constexpr int kBgCtlWords = 2 + 256;
void launch_adam_rows(const AdamArgs& a, const float* g_sparse, __nv_bfloat16* shadow, cudaStream_t st,
                      const BgAdam* bg = nullptr);
// a.w / m / v [rows, row_len] with a.row_touched: dense TF1 Adam (g = 0) of the untouched rows, claimed chunk by chunk from
// the top of the row range until launch_bg_stop; launch_adam_rows(..., &bg) afterwards does everything that is left
void launch_adam_bg(const AdamArgs& a, const BgAdam& bg, cudaStream_t st);
void launch_bg_stop(const BgAdam& bg, cudaStream_t st);
void set_trap_log_optim(unsigned int* host_mapped);
// U(-limit, limit) keyed by the GLOBAL element index; w: this rank's rows in local-tile order
void launch_xavier_init(float* w, int n_local_rows, int N, int H, float limit, unsigned long long seed,
                        unsigned stream_id, int world, int rank, cudaStream_t st);
void launch_sumsq(const float* x, long long n, float* partial, int nblocks, cudaStream_t st);
void launch_reduce_loss2(const float* partial, int n, const float* sumsq_partial, int n_sq, float lambda,
                         float inv_batch, float* loss_out, cudaStream_t st);
void launch_clear_flagged(int N, int H, float* g_enc, unsigned char* touched, int* touch_cnt, cudaStream_t st);
// the same two jobs over a LIST of rows (list[0] = count): dense TF1 Adam of the listed rows of w / m / v [rows, row_len] with
// their gradient rows g_sparse; zeroing of the listed gradient rows, flags and counts
void launch_adam_listed(const AdamArgs& a, const float* g_sparse, const int* list, cudaStream_t st);
void launch_clear_listed(int H, float* g_enc, unsigned char* touched, int* touch_cnt, const int* list, cudaStream_t st);

// ---- title.cu / title_sm100.cu: the title branch (Char_CNN.py:16-75, DAEs.py:153-201) ----------
constexpr int kTitleFpad = 512;    // feature columns of the output layer's operands (D = filters x widths <= 512, zero padded)
constexpr int kTitleMaxLen = 32;   // strmaxlen limit (25 in every shipped config)
constexpr int kTitleMaxWidths = 8;

struct CnnShape {
    int C, L, E, F, n_widths;            // charsize, strmaxlen, char_emb, filter_num, len(filter_size)
    int width[kTitleMaxWidths];          // filter_size
    int w_off[kTitleMaxWidths];          // offset of width i's [w][E][F] block inside conv_W
};
struct CnnFwdArgs {
    const long long* titles;             // [B, L] char ids, -1 = pad
    const float* emb; const float* conv_W; const float* conv_b;
    CnnShape shape;
    float* feat;                         // [B, D] max-over-time features (before dropout)
    unsigned char* argpos;               // [B, D] arg-max position
    __nv_bfloat16* feat_d;               // [bpad, 512] dropout(feat), bf16
    __nv_bfloat16* feat_dT;              // [512, bpad]
    int B, bpad;
    float kp_t;
    unsigned long long seed, step;
    int row_offset;
};
void launch_charcnn_fwd(const CnnFwdArgs& a, cudaStream_t st);
void launch_mix_weights(const float* rowsum, const float* titles_use, float kp_in, int B, int bpad, float* w_t,
                        float* w_p, cudaStream_t st);
struct CnnBwdArgs {
    const long long* titles;
    const float* emb; const float* conv_W;
    const float* conv_WT;                // conv_W with every width block transposed to [f][k][e] (launch_conv_transpose)
    CnnShape shape;
    const float* dh_partial;             // [2 halves][nsplit][bpad][256] split-K partials of d cost / d feat_d
    int nsplit, bpad, B;
    const float* feat; const unsigned char* argpos;
    float kp_t;
    unsigned long long seed, step;
    int row_offset;
    float* d;                            // [B, D] workspace: gradient at the arg-max positions
    float* dx;                           // [B, L, E] workspace: gradient at the embedded characters, per title
    float* g_emb; float* g_conv_W; float* g_conv_b;
};
void launch_charcnn_bwd(const CnnBwdArgs& a, cudaStream_t st);
void launch_conv_transpose(const float* W, float* WT, const CnnShape& s, cudaStream_t st);
// columns [col0, col0 + ncols) (ncols <= 0: all) of a [rows, row_len] variable -> w[r * ld + (c - col0)]
void launch_trunc_normal(float* w, long long rows, int row_len, int ld, float stddev, unsigned long long seed,
                         unsigned stream_id, cudaStream_t st, int col0 = 0, int ncols = 0);
void launch_cast_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st);
void launch_cast_block_bf16(const float* src, __nv_bfloat16* dst, long long rows, int cols, int ld, int col0, cudaStream_t st);
void launch_transpose_pad(const float* src, float* dst, int D, int N, int ld, int to_item_major, cudaStream_t st);

struct TitleTileArgs {                   // y_pred = title_score * w_t + sigmoid(z_dae) * w_p over one batch tile  (DAEs.py:175-181)
    const __nv_bfloat16* W_dec;          // [N, H]   frozen DAE decoder operand
    const __nv_bfloat16* h_d;            // [bpad, H]
    const float* b_dec;                  // [N]
    const __nv_bfloat16* W_out;          // [N, 512] title output layer (item-major)
    const __nv_bfloat16* feat_d;         // [bpad, 512]
    const float* b_out;                  // [N]
    const float* w_t; const float* w_p;  // [bpad]
    int N, H, batch, bpad;
    int kf;                              // 64-column chunks of the feature operand actually used (ceil(D / 64))
    // train (DAEs.py:194-196)
    const uint32_t* ybits; int ywords;
    __nv_bfloat16* dzT;                  // [N, bpad] d cost / d z_title
    float* db_out;                       // [N]
    float* loss_partial;
    float inv_batch;
    // predict
    float* out; long long ld_out; int n_out;
};
void launch_title_train(const TitleTileArgs& a, cudaStream_t st);
void launch_title_predict(const TitleTileArgs& a, cudaStream_t st);

// ---- topk.cu --------------------------------------------------------------------------------
struct TopkArgs {
    const float* scores;      // [B, ld]
    long long ld;
    int B, T, k;
    const int* seed_ptr;      // [B+1] CSR of seed track ids to exclude (may contain ids >= T or < 0)
    const int* seed_idx;
    int idx_base;             // added to output indices (item-sharded inference)
    int* out_idx;             // [B,k]  (-1 padded)
    float* out_score;         // [B,k]
    // candidate-list mode (fused decode + top-K): row r holds min(row_n[r], ld) entries, entry i is item remap[r * ld + i]
    const int* remap;         // [B, ld] or nullptr (entry i is item i)
    const int* row_n;         // [B] or nullptr (every row has T entries)
    int sigmoid_out;          // 1: scores are logits; out_score = sigmoid(logit)
    float* thr_out;           // [B] or nullptr.  Not null: threshold-only mode -- thr_out[r] = the filter threshold derived
                              // from the kreq-th largest score of row r (-inf when the row has fewer), where
                              // kreq = ceil((k + #seeds of row r) / thr_div); nothing else is written (seed_idx is not read)
    int thr_div;              // threshold-only mode: number of item shards that share the thresholds (0 / 1: none)
};
void launch_topk(const TopkArgs& a, cudaStream_t st);
// thr[r] = score[r, kp-1] (r < batch, -inf when that slot is padding or score == nullptr), +inf for r in [batch, rows)
void launch_thr_from_topk(const float* score, const int* idx, int kp, int batch, int rows, float* thr, cudaStream_t st);

// (r-precision, ndcg, clicks) per playlist from its ranked list [B, ld] (k ranks, -1 padded) and the answers CSR -> out [B, 3]
void launch_metrics(const int* cand, long long ld, int B, int k, const int* ans_ptr, const int* ans_idx, double* out,
                    cudaStream_t st);

// bounded-spin diagnostics: 8 words of mapped pinned host memory written before a trap (umma.cuh trap_report, k_barrier)
void set_trap_log_gemm(unsigned int* host_mapped);
void set_trap_log_title_gemm(unsigned int* host_mapped);
void set_trap_log_sparse(unsigned int* host_mapped);
// Load a kernel now: CUDA's lazy loading would otherwise synchronise the context at its first launch, against a
// cross-GPU flag barrier that may already be spinning.  (Asking every kernel for the largest shared-memory carveout was
// tried and reverted: the streaming optimizer kernels lose 30 % of their bandwidth with a minimal L1.)
#define PRELOAD_KERNEL(k) cudaFuncGetAttributes(&a, k)
// load every kernel of a translation unit (see sparse.cu: preload_sparse)
void preload_sparse();
void preload_optim();
void preload_gemm();
void preload_topk();
void preload_title_cnn();
void preload_title_gemm();
inline void preload_title() { preload_title_cnn(); preload_title_gemm(); }

}  // namespace dae
