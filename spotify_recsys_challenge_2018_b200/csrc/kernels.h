// Internal launcher interface between api.cu (the C-ABI) and the kernel translation units.
// All pointers are DEVICE pointers unless noted; every launcher enqueues on `st` and never syncs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dae {

constexpr int kTileItems = 128;    // UMMA M: catalogue items per accumulator tile
constexpr int kMaxBpad = 256;      // UMMA N limit: batch columns per tile
constexpr int kMaxRowNnz = 2048;   // per-playlist COO entries the row sorter holds in smem
constexpr float kEpsLog = 1e-10f;  // DAEs.py:42,98-99
constexpr float kNegWeight = 0.55f;  // DAEs.py:99

// error flag bits written by device-side validation (prepare kernels)
enum : int { kErrIndexRange = 1, kErrRowTooLong = 2, kErrYNotBinary = 4 };

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

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
void launch_ybits_set(const CsrWork& y, int B, uint32_t* ybits, int ywords, int set, int* err, cudaStream_t st);

struct EncodeArgs {
    const float* W_enc;   // [N,H] fp32 master
    const float* b_enc;   // [H]
    CsrWork x;            // val is overwritten with x_n (a3), needed again by the backward scatter
    float* rowsum;        // [B] s = sum_j x_d (DAEs.py:41)
    float* h;             // [B,H] fp32 sigma(a)
    __nv_bfloat16* h_d;   // [bpad,H]  dropout(h), bf16, rows >= B zero
    __nv_bfloat16* h_dT;  // [H,bpad]
    int B, bpad, H;
    float kp, kp_in;
    unsigned long long seed, step;
    int row_offset;       // global row index of local row 0 (data-parallel shards)
};
void launch_encode_fwd(const EncodeArgs& a, cudaStream_t st);

struct EncodeBwdArgs {
    const float* dh_partial;  // [nsplit, bpad, H]
    int nsplit;
    const float* h;           // [B,H]
    CsrWork x;                // col, val = x_n
    float* da;                // [B,H]
    float* g_enc;             // [N,H] scatter-add target (dW_enc rows, or the tied dW)
    unsigned char* touched;   // [N] or nullptr
    float* db_enc;            // [H]
    int B, bpad, H;
    float kp;
    unsigned long long seed, step;
    int row_offset;
};
void launch_encode_bwd(const EncodeBwdArgs& a, cudaStream_t st);
void launch_clear_touched(const CsrWork& x, int B, int H, float* g_enc, unsigned char* touched, cudaStream_t st);

// ---- gemm_sm100.cu -------------------------------------------------------------------------
struct DecodeArgs {
    const __nv_bfloat16* W;    // [N,H] bf16 decoder operand
    const __nv_bfloat16* h_d;  // [rows_alloc,H] bf16
    const float* bias;         // [N]
    int N, H;
    int batch;                 // valid rows
    int bpad;                  // rows per batch tile (multiple of 64, <= 256)
    int n_batch_tiles;         // predict only
    // train
    const uint32_t* ybits;     // [N, ywords]
    int ywords;
    __nv_bfloat16* dzT;        // [N,bpad]
    float* db_dec;             // [N]
    float* loss_partial;       // [grid]
    float inv_batch;
    // predict
    float* out;                // [batch, ld_out] fp32 scores
    long long ld_out;
    int n_out;                 // columns written (n_tracks or N)
    const float* mix_wp;       // [batch] or nullptr: y_pred = title*w_t + p*w_p (DAEs.py:180)
    const float* mix_wt;
    const float* title_score;  // [batch, ld_out] or nullptr
};
int decode_grid(int N, int n_batch_tiles);
void launch_decode_train(const DecodeArgs& a, cudaStream_t st);    // G1: z, loss, dz, db_dec
void launch_decode_predict(const DecodeArgs& a, cudaStream_t st);  // G1: z, sigmoid, scores

struct DwArgs {
    const __nv_bfloat16* dzT;   // [N,bpad]
    const __nv_bfloat16* h_dT;  // [H,bpad]
    float* g;                   // [N,H] fp32, overwritten
    int N, H, bpad;
};
void launch_dw(const DwArgs& a, cudaStream_t st);                  // G2: dW_dec = dz^T . h_d

struct DhArgs {
    const __nv_bfloat16* dzT;   // [N,bpad]
    const __nv_bfloat16* W;     // [N,H]
    float* partial;             // [nsplit,bpad,H]
    int N, H, bpad, nsplit;
    int lbo, sbo;               // MN-major descriptor strides (bytes); 0 -> defaults
};
int dh_nsplit(int N);
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
};
void launch_adam(const AdamArgs& a, cudaStream_t st);
void launch_xavier_init(float* w, long long n, float limit, unsigned long long seed, unsigned stream_id,
                        cudaStream_t st);
void launch_cast_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st);
void launch_sumsq(const float* x, long long n, float* partial, int nblocks, cudaStream_t st);
void launch_reduce_loss2(const float* partial, int n, const float* sumsq_partial, int n_sq, float lambda,
                         float inv_batch, float* loss_out, cudaStream_t st);
void launch_clear_flagged(int N, int H, float* g_enc, unsigned char* touched, cudaStream_t st);

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
};
void launch_topk(const TopkArgs& a, cudaStream_t st);

}  // namespace dae
