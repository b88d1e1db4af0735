// Sparse side of the DAE step: the reader's COO batch -> de-duplicated CSR, the y bitmask, the
// encode forward (gather) and the encode backward (sparse-row scatter-add).  HBM-bound integer /
// gather work: coalesced row reads of the item-major weight matrix, no densification.
//
// Reference semantics restated here:
//   tf.sparse_tensor_to_dense(validate_indices=False)  models/DAEs.py:33-38   (duplicates: last wins)
//   dropout + reduce_sum + divide                      models/DAEs.py:40-42
//   encoder: matmul + bias + sigmoid + dropout         models/DAEs.py:64-70
#include <cuda_runtime.h>

#include "kernels.h"
#include "philox.cuh"

namespace dae {

// ------------------------------------------------------------------------------------------
// COO [nnz,2] int64 (row-in-batch, item) -> per-row sorted unique columns, value of the LAST
// occurrence (utils/data_reader.py:48-54 emits "track block then artist block", so rows are not
// contiguous in the COO list: bucket by row first).
// ------------------------------------------------------------------------------------------
__global__ void k_coo_count(const long long* __restrict__ pos, int nnz, int B, int N, int* __restrict__ cnt,
                            int* __restrict__ err) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const long long r = pos[2 * (size_t)e], c = pos[2 * (size_t)e + 1];
    if (r < 0 || r >= B || c < 0 || c >= N) {
        atomicOr(err, kErrIndexRange);
        return;
    }
    atomicAdd(&cnt[r], 1);
}

__global__ void k_row_scan(const int* __restrict__ cnt, int B, int* __restrict__ row_ptr, int* __restrict__ cursor,
                           int* __restrict__ err) {
    // single block; B is a few hundred to a few thousand
    __shared__ int s_part[1024];
    const int t = threadIdx.x;
    const int per = (B + blockDim.x - 1) / blockDim.x;
    const int lo = min(t * per, B), hi = min(lo + per, B);
    int sum = 0;
    for (int i = lo; i < hi; ++i) {
        const int c = cnt[i];
        if (c > kMaxRowNnz) atomicOr(err, kErrRowTooLong);
        sum += c;
    }
    s_part[t] = sum;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
        const int v = (t >= o) ? s_part[t - o] : 0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    int run = s_part[t] - sum;
    for (int i = lo; i < hi; ++i) {
        row_ptr[i] = run;
        cursor[i] = run;
        run += cnt[i];
    }
    if (t == blockDim.x - 1) row_ptr[B] = s_part[t];
}

__global__ void k_coo_fill(const long long* __restrict__ pos, int nnz, int B, int N, int* __restrict__ cursor,
                           unsigned long long* __restrict__ keys) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const long long r = pos[2 * (size_t)e], c = pos[2 * (size_t)e + 1];
    if (r < 0 || r >= B || c < 0 || c >= N) return;
    const int slot = atomicAdd(&cursor[r], 1);
    keys[slot] = (static_cast<unsigned long long>(c) << 32) | static_cast<unsigned int>(e);
}

// One CTA per row: bitonic sort of (col, entry) keys in shared memory, keep the last entry of each
// column run (largest entry index == last occurrence), compact.
__global__ void __launch_bounds__(256)
k_row_sort_dedup(const int* __restrict__ row_ptr, const unsigned long long* __restrict__ keys,
                 const float* __restrict__ val_in, int* __restrict__ row_len, int* __restrict__ col_out,
                 float* __restrict__ val_out) {
    __shared__ unsigned long long s_key[kMaxRowNnz];
    __shared__ int s_scan[256];
    const int r = blockIdx.x;
    const int beg = row_ptr[r];
    int n = row_ptr[r + 1] - beg;
    if (n > kMaxRowNnz) n = kMaxRowNnz;   // flagged by k_row_scan
    if (n == 0) {
        if (threadIdx.x == 0) row_len[r] = 0;
        return;
    }
    int npow = 1;
    while (npow < n) npow <<= 1;
    for (int i = threadIdx.x; i < npow; i += blockDim.x) s_key[i] = (i < n) ? keys[beg + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = s_key[i], b = s_key[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_key[i] = b; s_key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    // keep flags + block scan (each thread owns a contiguous span so output stays column-sorted)
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int lo = min((int)threadIdx.x * per, n), hi = min(lo + per, n);
    int mine = 0;
    for (int i = lo; i < hi; ++i) {
        const bool keep = (i == n - 1) || ((s_key[i] >> 32) != (s_key[i + 1] >> 32));
        mine += keep ? 1 : 0;
    }
    s_scan[threadIdx.x] = mine;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
        const int v = (threadIdx.x >= o) ? s_scan[threadIdx.x - o] : 0;
        __syncthreads();
        s_scan[threadIdx.x] += v;
        __syncthreads();
    }
    int w = s_scan[threadIdx.x] - mine;
    for (int i = lo; i < hi; ++i) {
        const bool keep = (i == n - 1) || ((s_key[i] >> 32) != (s_key[i + 1] >> 32));
        if (keep) {
            col_out[beg + w] = static_cast<int>(s_key[i] >> 32);
            val_out[beg + w] = val_in[static_cast<unsigned int>(s_key[i] & 0xFFFFFFFFull)];
            ++w;
        }
    }
    if (threadIdx.x == blockDim.x - 1) row_len[r] = s_scan[threadIdx.x];
}

void launch_coo_to_csr(const long long* pos, const float* val, int nnz, int B, int N, CsrWork w, int* err,
                       cudaStream_t st) {
    cudaMemsetAsync(w.cnt, 0, sizeof(int) * B, st);
    if (nnz > 0) k_coo_count<<<(nnz + 255) / 256, 256, 0, st>>>(pos, nnz, B, N, w.cnt, err);
    k_row_scan<<<1, 1024, 0, st>>>(w.cnt, B, w.row_ptr, w.cursor, err);
    if (nnz > 0) k_coo_fill<<<(nnz + 255) / 256, 256, 0, st>>>(pos, nnz, B, N, w.cursor, w.keys);
    k_row_sort_dedup<<<B, 256, 0, st>>>(w.row_ptr, w.keys, val, w.row_len, w.col, w.val);
}

// ------------------------------------------------------------------------------------------
// encode forward: one CTA per playlist row, blockDim = H threads (H <= 256), thread k owns
// hidden unit k; every gathered W_enc row is one fully coalesced H*4-byte read.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += s_red[i];   // fixed order: deterministic
    return t;
}

__global__ void __launch_bounds__(512)
k_encode_fwd(const float* __restrict__ W, const float* __restrict__ b_enc, const int* __restrict__ row_ptr,
             const int* __restrict__ row_len, const int* __restrict__ col, const float* __restrict__ val,
             const __grid_constant__ PubInput pub, float* __restrict__ rowsum, float* __restrict__ h, __nv_bfloat16* __restrict__ h_d,
             __nv_bfloat16* __restrict__ h_dT, int B, int bpad, int H, int K, int row0, int bcast, float kp,
             float kp_in, unsigned long long seed, unsigned long long step, int row_offset, const __grid_constant__ PeerTable pt) {
    __shared__ __align__(16) float s_x[kMaxRowNnz];
    __shared__ int s_c[kMaxRowNnz];
    __shared__ float s_red[16];
    const int r = blockIdx.x;
    const int k = threadIdx.x;
    const int world = pt.world;
    // this rank's rows of h / h_d ([rows, H]) and columns of h_d^T ([H, K]); training stores them into EVERY rank's
    // copy (the dense side contracts over the global batch), inference keeps them local
    const size_t hT_off = (size_t)k * K + row0 + r;
    const size_t h_off = (size_t)(row0 + r) * H + k;
    const int n_dst = bcast ? world : 1;
    // h_d^T ([H, K], one 2-byte element per thread at a stride of K) is only written here on ONE GPU: with peers, 2-byte
    // stores over NVLink are one packet each (the 8-GPU encode spent most of its 0.11 ms on them) -- every rank transposes
    // its own copy of the global h_d after barrier B1 instead (k_transpose_hd)
    const bool write_hT = !(bcast && world > 1);
    if (r >= B) {   // padding rows of the tensor-core operand
        if (k < H) {
            const __nv_bfloat16 z = __float2bfloat16(0.f);
            for (int s = 0; s < n_dst; ++s) {
                (bcast ? peer_ptr(pt, s, h_d) : h_d)[h_off] = z;
            }
            if (write_hT) h_dT[hT_off] = z;
        }
        return;
    }
    const int beg = row_ptr[r], n = row_len[r];
    const uint32_t grow = static_cast<uint32_t>(r + row_offset);
    // a3: x_d = x / kp_in * keep ; s = sum x_d
    float part = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = col[beg + i];
        const float v = val[beg + i];
        const bool keep = philox_keep(seed, kStreamInput, step, grow, static_cast<uint32_t>(c), kp_in);
        const float xd = keep ? __fdiv_rn(v, kp_in) : 0.f;
        s_x[i] = xd;
        s_c[i] = c;
        part += xd;
    }
    const float s = block_sum(part, s_red);
    __syncthreads();
    // x_n and its column are re-read by the sparse-row scatter of EVERY rank: each rank keeps a copy of the whole global
    // batch's published input (segment = producing rank), so the scatter reads local memory only
    const size_t seg_e = (size_t)(bcast ? pt.rank : 0) * pub.seg_nnz + beg;
    const int seg_r = (bcast ? pt.rank : 0) * pub.seg_rows + r;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float xn = __fdiv_rn(s_x[i], s + kEpsLog);
        s_x[i] = xn;
        for (int d = 0; d < n_dst; ++d) {
            (bcast ? peer_ptr(pt, d, pub.xn) : pub.xn)[seg_e + i] = xn;
            (bcast ? peer_ptr(pt, d, pub.col) : pub.col)[seg_e + i] = s_c[i];
        }
    }
    if (threadIdx.x == 0) {
        rowsum[r] = s;
        for (int d = 0; d < n_dst; ++d) {
            (bcast ? peer_ptr(pt, d, pub.row_ptr) : pub.row_ptr)[seg_r] = beg;
            (bcast ? peer_ptr(pt, d, pub.row_len) : pub.row_len)[seg_r] = n;
        }
    }
    __syncthreads();
    // a4: a = sum_j x_n[j] * W_enc[col_j, :].  Thread (g, t): row group g takes entries j = g (mod G),
    // lane t owns columns [4t, 4t+4) as one float4 -> every gathered row is a run of coalesced 16 B loads
    // and up to 8*G rows (64 at H = 256) are in flight per CTA: the kernel lasts as long as its longest playlist
    // (250 entries), i.e. ceil(250 / 64) round trips to HBM.  Groups are combined through smem in a fixed order.
    // Rows live on their owner GPU (tile-cyclic); remote rows are plain loads over NVLink.
    const int tpr = H >> 2;                      // threads per row
    const int G = blockDim.x / tpr;              // concurrent row groups
    const int g = threadIdx.x / tpr, t = threadIdx.x - g * tpr;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto row_of = [&](int c) -> const float4* {
        const float* base = world == 1 ? W : peer_ptr(pt, item_owner(c, world), W);
        return reinterpret_cast<const float4*>(base + (size_t)(world == 1 ? c : item_local(c, world)) * H) + t;
    };
    if (g < G) {
        int i = g;
        for (; i + 7 * G < n; i += 8 * G) {
            float x[8];
            float4 w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                x[u] = s_x[i + u * G];
                w[u] = x[u] != 0.f ? __ldg(row_of(s_c[i + u * G])) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                acc.x = fmaf(x[u], w[u].x, acc.x); acc.y = fmaf(x[u], w[u].y, acc.y);
                acc.z = fmaf(x[u], w[u].z, acc.z); acc.w = fmaf(x[u], w[u].w, acc.w);
            }
        }
        for (; i < n; i += G) {
            const float x0 = s_x[i];
            if (x0 != 0.f) {
                const float4 w0 = __ldg(row_of(s_c[i]));
                acc.x = fmaf(x0, w0.x, acc.x); acc.y = fmaf(x0, w0.y, acc.y);
                acc.z = fmaf(x0, w0.z, acc.z); acc.w = fmaf(x0, w0.w, acc.w);
            }
        }
    }
    __syncthreads();                              // s_x / s_c are dead from here: reuse s_x as the [G][H] staging area
    float* s_acc = s_x;
    if (g < G) *reinterpret_cast<float4*>(s_acc + g * H + 4 * t) = acc;
    __syncthreads();
    if (k >= H) return;
    float a = 0.f;
    for (int gg = 0; gg < G; ++gg) a += s_acc[gg * H + k];
    a += b_enc[k];
    const float hv = __fdividef(1.f, 1.f + __expf(-a));
    const bool keep = philox_keep(seed, kStreamHidden, step, grow, static_cast<uint32_t>(k), kp);
    const float hd = keep ? __fdiv_rn(hv, kp) : 0.f;
    const __nv_bfloat16 hb = __float2bfloat16(hd);
    h[h_off] = hv;                                 // fp32 h: only this rank's da needs it
    for (int s2 = 0; s2 < n_dst; ++s2) (bcast ? peer_ptr(pt, s2, h_d) : h_d)[h_off] = hb;
    if (write_hT) h_dT[hT_off] = hb;
}

// h_d [K, H] -> h_d^T [H, K] (bf16), 32 x 32 tiles through shared memory: the local transpose of the gathered global batch
__global__ void k_transpose_hd(const __nv_bfloat16* __restrict__ h_d, __nv_bfloat16* __restrict__ h_dT, int K, int H) {
    __shared__ __nv_bfloat16 tile[32][34];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += 8) tile[y][threadIdx.x] = h_d[(size_t)(r0 + y) * H + c0 + threadIdx.x];
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += 8) h_dT[(size_t)(c0 + y) * K + r0 + threadIdx.x] = tile[threadIdx.x][y];
}
void launch_transpose_hd(const __nv_bfloat16* h_d, __nv_bfloat16* h_dT, int K, int H, cudaStream_t st) {
    k_transpose_hd<<<dim3(K / 32, H / 32), dim3(32, 8), 0, st>>>(h_d, h_dT, K, H);
}

void launch_encode_fwd(const EncodeArgs& a, cudaStream_t st) {
    const int threads = 512;                      // G = 2048 / H row groups of H/4 threads
    k_encode_fwd<<<a.bpad, threads, 0, st>>>(a.W_enc, a.b_enc, a.x.row_ptr, a.x.row_len, a.x.col, a.x.val, a.pub,
                                             a.rowsum, a.h, a.h_d, a.h_dT, a.B, a.bpad, a.H, a.K, a.row0, a.bcast,
                                             a.kp, a.kp_in, a.seed, a.step, a.row_offset, a.pt);
}

// ------------------------------------------------------------------------------------------
// y of the global batch over the item rows this rank owns (bitmask for the decode epilogue)
// ------------------------------------------------------------------------------------------
__global__ void k_ybits_shard(const int* __restrict__ row_ptr, const int* __restrict__ row_len,
                              const int* __restrict__ col, const float* __restrict__ val, uint32_t* __restrict__ ybits,
                              int ywords, int bpad, int* __restrict__ err, const __grid_constant__ PeerTable pt) {
    const int r = blockIdx.x, s = blockIdx.y;                  // row r of rank s's batch
    const int world = pt.world;
    const int* prow_ptr = world == 1 ? row_ptr : peer_ptr(pt, s, row_ptr);
    const int* prow_len = world == 1 ? row_len : peer_ptr(pt, s, row_len);
    const int* pcol = world == 1 ? col : peer_ptr(pt, s, col);
    const float* pval = world == 1 ? val : peer_ptr(pt, s, val);
    const int beg = prow_ptr[r], n = prow_len[r];
    const int bit = s * bpad + r;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = pcol[beg + i];
        if (world != 1 && item_owner(c, world) != pt.rank) continue;
        const float v = pval[beg + i];
        if (v == 1.0f) atomicOr(ybits + (size_t)(world == 1 ? c : item_local(c, world)) * ywords + (bit >> 5), 1u << (bit & 31));
        else if (v != 0.0f) atomicOr(err, kErrYNotBinary);     // the runner feeds ones (main_train.py:206)
    }
}
void launch_ybits_shard(const YbitsArgs& a, cudaStream_t st) {
    cudaMemsetAsync(a.ybits, 0, sizeof(uint32_t) * (size_t)a.n_local * a.ywords, st);
    k_ybits_shard<<<dim3(a.B, a.pt.world), 128, 0, st>>>(a.y.row_ptr, a.y.row_len, a.y.col, a.y.val, a.ybits, a.ywords, a.bpad,
                                                         a.err, a.pt);
}

// ------------------------------------------------------------------------------------------
// Rows of W_enc this rank owns that ANY playlist of the global batch lists in its input (before dropout): the only
// rows whose dW_enc can be non-zero and the only rows the encode gathers.  Known as soon as the slot CSRs are (barrier
// A), long before the backward: every other row's dense Adam update (g == 0) may start right away (optim.cu: k_adam_bg).
// ------------------------------------------------------------------------------------------
__global__ void k_touch_shard(const int* __restrict__ row_ptr, const int* __restrict__ row_len, const int* __restrict__ col,
                              unsigned char* __restrict__ touched, int* __restrict__ touch_cnt, int* __restrict__ hot_list,
                              int* __restrict__ touched_list, const __grid_constant__ PeerTable pt) {
    const int r = blockIdx.x, s = blockIdx.y;                  // row r of rank s's batch
    const int world = pt.world;
    const int* prow_ptr = world == 1 ? row_ptr : peer_ptr(pt, s, row_ptr);
    const int* prow_len = world == 1 ? row_len : peer_ptr(pt, s, row_len);
    const int* pcol = world == 1 ? col : peer_ptr(pt, s, col);
    const int beg = prow_ptr[r], n = prow_len[r];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = pcol[beg + i];
        if (world != 1 && item_owner(c, world) != pt.rank) continue;
        const int lc = world == 1 ? c : item_local(c, world);
        touched[lc] = 1;
        // playlists of the global batch that list the row (integer atomics: order-independent).  The THIRD one puts the
        // row on the hot list: hot_list[0] = number of entries, rows from hot_list[1] on (each exactly once)
        const int before = atomicAdd(touch_cnt + lc, 1);
        if (before == 0) touched_list[1 + atomicAdd(touched_list, 1)] = lc;      // every listed row, once (same layout)
        if (before == 2) hot_list[1 + atomicAdd(hot_list, 1)] = lc;
    }
}
void launch_touch_shard(const CsrWork& x, unsigned char* touched, int* touch_cnt, int* hot_list, int* touched_list, int B,
                        const PeerTable& pt, cudaStream_t st) {
    cudaMemsetAsync(hot_list, 0, sizeof(int), st);
    cudaMemsetAsync(touched_list, 0, sizeof(int), st);
    k_touch_shard<<<dim3(B, pt.world), 128, 0, st>>>(x.row_ptr, x.row_len, x.col, touched, touch_cnt, hot_list, touched_list, pt);
}

// ------------------------------------------------------------------------------------------
// encode backward, part 1: dh -> da for the whole global batch
// ------------------------------------------------------------------------------------------
// fixed-order sum of the split-K partials [bt][split][bpad][H] of batch tile bt (= the rows of rank bt) -> straight into
// rank bt's dh_sum, slot [this rank][row][H]: the reduce-scatter half of the dh all-reduce, as 1 KB peer stores
__global__ void k_reduce_splits(const float* __restrict__ partial, int nsplit, int bpad, int H, float* __restrict__ out,
                                const __grid_constant__ PeerTable pt) {
    const int bt = blockIdx.y, r = blockIdx.x, k = threadIdx.x;
    if (k >= H) return;
    const size_t stride = (size_t)bpad * H;
    const float* p = partial + (size_t)bt * nsplit * stride + (size_t)r * H + k;
    int s = 0;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    for (; s + 4 <= nsplit; s += 4) {                     // fixed order: deterministic
        d0 += p[(size_t)s * stride];
        d1 += p[(size_t)(s + 1) * stride];
        d2 += p[(size_t)(s + 2) * stride];
        d3 += p[(size_t)(s + 3) * stride];
    }
    for (; s < nsplit; ++s) d0 += p[(size_t)s * stride];
    float* dst = pt.world == 1 ? out : peer_ptr(pt, bt, out);
    dst[((size_t)pt.rank * bpad + r) * H + k] = (d0 + d1) + (d2 + d3);
}
void launch_reduce_splits(const float* partial, int nsplit, int bpad, int H, const PeerTable& pt, float* dh_sum, cudaStream_t st) {
    k_reduce_splits<<<dim3(bpad, pt.world), 256, 0, st>>>(partial, nsplit, bpad, H, dh_sum, pt);
}

// Row i of THIS rank's playlists: dh = sum over ranks q (fixed order) of the item-shard sums they stored into slot q of
// the local dh_sum, da = dh * (keep / kp) * h (1 - h), stored into every rank's da [K, H] at row rank * bpad + i: the
// all-gather half.  Per rank and step 2 * (world - 1) * bpad * H * 4 bytes cross NVLink as stores, none as loads.
__global__ void __launch_bounds__(256)
k_da_own(const float* __restrict__ dh_sum, const float* __restrict__ h, float* __restrict__ da_out, int B, int bpad, int H,
         float kp, unsigned long long seed, unsigned long long step, int row_offset0, const __grid_constant__ PeerTable pt) {
    // 4 rows per block, thread = 4 consecutive hidden units of one row
    const int H4 = H >> 2;
    const int sub = threadIdx.x / H4, k4 = threadIdx.x - sub * H4;
    const int i = blockIdx.x * (blockDim.x / H4) + sub;
    if (i >= bpad) return;
    const size_t o = (((size_t)pt.rank * bpad + i) * H >> 2) + k4;    // float4 index of the row in h and da
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < B) {
        float4 part[kMaxWorld];
#pragma unroll
        for (int q = 0; q < kMaxWorld; ++q)
            if (q < pt.world) part[q] = reinterpret_cast<const float4*>(dh_sum)[(((size_t)q * bpad + i) * H >> 2) + k4];
        float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < kMaxWorld; ++q)                            // rank order
            if (q < pt.world) { dh.x += part[q].x; dh.y += part[q].y; dh.z += part[q].z; dh.w += part[q].w; }
        const float4 hv = reinterpret_cast<const float4*>(h)[o];
        // the same key as the forward's mask (k_encode_fwd: local row + row_offset = row_offset0 + rank * B)
        const uint32_t grow = static_cast<uint32_t>(row_offset0 + pt.rank * B + i);
        const float ikp = __fdiv_rn(1.f, kp);
        r.x = philox_keep(seed, kStreamHidden, step, grow, static_cast<uint32_t>(4 * k4 + 0), kp) ? dh.x * ikp * (hv.x * (1.f - hv.x)) : 0.f;
        r.y = philox_keep(seed, kStreamHidden, step, grow, static_cast<uint32_t>(4 * k4 + 1), kp) ? dh.y * ikp * (hv.y * (1.f - hv.y)) : 0.f;
        r.z = philox_keep(seed, kStreamHidden, step, grow, static_cast<uint32_t>(4 * k4 + 2), kp) ? dh.z * ikp * (hv.z * (1.f - hv.z)) : 0.f;
        r.w = philox_keep(seed, kStreamHidden, step, grow, static_cast<uint32_t>(4 * k4 + 3), kp) ? dh.w * ikp * (hv.w * (1.f - hv.w)) : 0.f;
    }
    if (pt.world == 1) { reinterpret_cast<float4*>(da_out)[o] = r; return; }
#pragma unroll
    for (int d = 0; d < kMaxWorld; ++d)
        if (d < pt.world) reinterpret_cast<float4*>(peer_ptr(pt, d, da_out))[o] = r;
}

// db_enc[k] = sum_r da[r,k]: 8 row groups per block, combined in a fixed order
__global__ void k_colsum(const float* __restrict__ x, int rows, int H, float* __restrict__ out) {
    __shared__ float s_p[8][32];
    const int k = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (k < H)
        for (int r = threadIdx.y; r < rows; r += 8) s += x[(size_t)r * H + k];
    s_p[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && k < H) {
        float t = 0.f;
        for (int g = 0; g < 8; ++g) t += s_p[g][threadIdx.x];
        out[k] = t;
    }
}

void launch_da_own(const DaArgs& a, cudaStream_t st) {
    const int rows_per_block = 256 / (a.H / 4);
    k_da_own<<<(a.bpad + rows_per_block - 1) / rows_per_block, 256, 0, st>>>(a.dh_sum, a.h, a.da, a.B, a.bpad, a.H, a.kp, a.seed,
                                                                            a.step, a.row_offset0, a.pt);
}
void launch_da_colsum(const DaArgs& a, cudaStream_t st) {
    k_colsum<<<(a.H + 31) / 32, dim3(32, 8), 0, st>>>(a.da, a.bpad * a.pt.world, a.H, a.db_enc);
}

// ------------------------------------------------------------------------------------------
// encode backward, part 2: sparse-row scatter-add dW_enc[col_j,:] += x_n[j] * da[row,:] into the rows
// THIS rank owns.  Every rank holds the published input (x_n, columns) of the whole global batch (k_encode_fwd
// stores its rows into every copy), so block (r, s) walks row r of rank s's segment in LOCAL memory and keeps the
// entries whose item tile lives here; da is local too.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_scatter_shard(const __grid_constant__ PubInput pub, const float* __restrict__ da, float* __restrict__ g_enc,
                const int* __restrict__ touch_cnt, int max_cnt, int bpad, int H, const __grid_constant__ PeerTable pt) {
    __shared__ __align__(16) float s_da[256];
    const int r = blockIdx.x, s = blockIdx.y;
    const int world = pt.world;
    const int* pcol = pub.col + (size_t)s * pub.seg_nnz;
    const float* pxn = pub.xn + (size_t)s * pub.seg_nnz;
    if (threadIdx.x < H) s_da[threadIdx.x] = da[((size_t)s * bpad + r) * H + threadIdx.x];
    __syncthreads();
    // thread (g, t) takes entries j = g (mod G), lane t adds 4 columns with one 16-byte reduction
    const int tpr = H >> 2, G = blockDim.x / tpr;
    const int g = threadIdx.x / tpr, t = threadIdx.x - g * tpr;
    if (g >= G) return;
    const float4 dav = *reinterpret_cast<const float4*>(s_da + 4 * t);
    const int beg = pub.row_ptr[s * pub.seg_rows + r], n = pub.row_len[s * pub.seg_rows + r];
    for (int i = g; i < n; i += G) {
        const int c = pcol[beg + i];
        if (world != 1 && item_owner(c, world) != pt.rank) continue;
        const float x = pxn[beg + i];
        if (x != 0.f) {
            const int lc = world == 1 ? c : item_local(c, world);
            if (touch_cnt != nullptr && touch_cnt[lc] > max_cnt) continue;     // deterministic mode: gathered by k_scatter_det
            float* dst = g_enc + (size_t)lc * H + 4 * t;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x * dav.x), "f"(x * dav.y),
                         "f"(x * dav.z), "f"(x * dav.w)
                         : "memory");
        }
    }
}

// Deterministic form ("determinism test for scatter-add").  fp32 addition commutes, so a row with at
// most TWO contributing playlists is order-independent under the red.add form above (0 + a + b == 0 + b + a exactly);
// only the rows listed by three or more playlists of the global batch (touch_cnt > 2: the popular head of the
// catalogue, a few hundred rows) need a fixed order.  One warp per such row gathers its contributions: every lane
// binary-searches the row's catalogue id in the (column-sorted, de-duplicated) input rows of 8 playlists at a time --
// the 8 searches advance in lockstep, so 8 loads are in flight per lane instead of one dependent chain -- and the hits
// are added in ascending global batch row.  dW_enc (and everything downstream of it) is then bit-identical run to run.
// 128 threads x <= 64 registers: the blocks must fit NEXT to the decoder update's CTAs (148 x 640 threads x 64 registers,
// 217 KB of shared memory), under which this kernel runs; the first version (256 x 97) only got on an SM once those had left.
constexpr int kDetRows = 4;     // playlists per lane per sweep: 128 playlists of the global batch per warp sweep
__global__ void __maxnreg__(64)
k_scatter_det(const __grid_constant__ PubInput pub, const float* __restrict__ da, float* __restrict__ g_enc,
              const int* __restrict__ hot_list, int B, int bpad, int H, const __grid_constant__ PeerTable pt) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int world = pt.world;
    const int K = world * B;                                   // valid rows of the global batch (rank-major)
    const int cpl = H >> 5;                                    // columns per lane: k = lane + 32 j
    const int n_hot = hot_list[0];
    {
        // one warp per row of the hot list (the list's order is arbitrary; every row is independent of the others)
        for (int hi_ = warp_global; hi_ < n_hot; hi_ += nwarps) {
            const int lrow = hot_list[1 + hi_];
            const int c = world == 1 ? lrow : item_global(lrow, world, pt.rank);
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            bool any = false;
            for (int g0 = 0; g0 < K; g0 += 32 * kDetRows) {
                // lane l searches playlists g0 + 32 u + l, u = 0..7, in lockstep
                int lo[kDetRows], hi[kDetRows], len[kDetRows];
                const int* pc[kDetRows];
                const float* px[kDetRows];
#pragma unroll
                for (int u = 0; u < kDetRows; ++u) {
                    const int gi = g0 + 32 * u + lane;
                    lo[u] = 0; hi[u] = 0; len[u] = 0; pc[u] = pub.col; px[u] = pub.xn;
                    if (gi < K) {
                        const int s = gi / B, r = gi - s * B;
                        const int beg = pub.row_ptr[s * pub.seg_rows + r];
                        len[u] = hi[u] = pub.row_len[s * pub.seg_rows + r];
                        pc[u] = pub.col + (size_t)s * pub.seg_nnz + beg;
                        px[u] = pub.xn + (size_t)s * pub.seg_nnz + beg;
                    }
                }
                bool more = true;
                while (more) {                                 // lower bound of c in each row's sorted columns
                    more = false;
                    int v[kDetRows];
#pragma unroll
                    for (int u = 0; u < kDetRows; ++u) v[u] = lo[u] < hi[u] ? pc[u][(lo[u] + hi[u]) >> 1] : 0;
#pragma unroll
                    for (int u = 0; u < kDetRows; ++u) {
                        if (lo[u] < hi[u]) {
                            const int mid = (lo[u] + hi[u]) >> 1;
                            if (v[u] < c) lo[u] = mid + 1; else hi[u] = mid;
                            more |= lo[u] < hi[u];
                        }
                    }
                }
                float x[kDetRows];
#pragma unroll
                for (int u = 0; u < kDetRows; ++u) x[u] = (lo[u] < len[u] && pc[u][lo[u]] == c) ? px[u][lo[u]] : 0.f;
#pragma unroll
                for (int u = 0; u < kDetRows; ++u) {           // ascending global batch row
                    unsigned hits = __ballot_sync(0xffffffffu, x[u] != 0.f);
                    while (hits) {
                        const int b = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const float xb = __shfl_sync(0xffffffffu, x[u], b);
                        const int gb = g0 + 32 * u + b;
                        const int s = gb / B, r = gb - s * B;
                        const float* drow = da + ((size_t)s * bpad + r) * H;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < cpl) acc[j] = __fadd_rn(acc[j], __fmul_rn(xb, drow[lane + 32 * j]));
                        any = true;
                    }
                }
            }
            if (any) {
                float* dst = g_enc + (size_t)lrow * H;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < cpl) dst[lane + 32 * j] = acc[j];   // these rows are zero between steps and untouched by the red.add form
            }
        }
    }
}

void launch_scatter_shard(const ScatterArgs& a, cudaStream_t st) {
    // deterministic: rows with <= 2 contributors through red.add (order-independent), the others gathered in a fixed order
    k_scatter_shard<<<dim3(a.B, a.pt.world), 256, 0, st>>>(a.pub, a.da, a.g_enc, a.deterministic ? a.touch_cnt : nullptr, 2,
                                                           a.bpad, a.H, a.pt);
    if (a.deterministic)
        k_scatter_det<<<148 * 4, 128, 0, st>>>(a.pub, a.da, a.g_enc, a.hot_list, a.B, a.bpad, a.H, a.pt);
}

// out[global item] = the owner's src[local item]
__global__ void k_gather_items_f32(const float* __restrict__ src, float* __restrict__ out, int N, const __grid_constant__ PeerTable pt) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    out[g] = pt.world == 1 ? src[g] : peer_ptr(pt, item_owner(g, pt.world), src)[item_local(g, pt.world)];
}
void launch_gather_items_f32(const float* src_local, float* out, int N, const PeerTable& pt, cudaStream_t st) {
    k_gather_items_f32<<<(N + 255) / 256, 256, 0, st>>>(src_local, out, N, pt);
}
__global__ void k_gather_rows_bf16(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ out, int N, int H8,
                                   const __grid_constant__ PeerTable pt) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * H8; i += stride) {
        const int g = (int)(i / H8), c = (int)(i - (long long)g * H8);
        const uint4* row = reinterpret_cast<const uint4*>(peer_ptr(pt, item_owner(g, pt.world), src)) + (size_t)item_local(g, pt.world) * H8;
        reinterpret_cast<uint4*>(out)[i] = row[c];
    }
}
void launch_gather_rows_bf16(const __nv_bfloat16* src_local, __nv_bfloat16* out, int N, int H, const PeerTable& pt,
                             cudaStream_t st) {
    k_gather_rows_bf16<<<1184, 256, 0, st>>>(src_local, out, N, H / 8, pt);
}

// out[i] = sum_s part_s[i], ranks in a fixed order (identical result on every rank)
__global__ void k_sum_partials(const float* __restrict__ part, float* __restrict__ out, int n, const __grid_constant__ PeerTable pt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t = 0.f;
    for (int s = 0; s < pt.world; ++s) t += peer_ptr(pt, s, part)[i];
    out[i] = t;
}
void launch_sum_partials(const float* part_local, float* out, int n, const PeerTable& pt, cudaStream_t st) {
    k_sum_partials<<<(n + 255) / 256, 256, 0, st>>>(part_local, out, n, pt);
}

// ------------------------------------------------------------------------------------------
// cross-GPU barrier on the stream: rank r stores `epoch` into slot r of every peer's flag array
// (release, system scope), then waits until all of its own slots have reached `epoch`.  Kernels
// enqueued before it have completed (stream order), so their NVLink stores are visible to the peers
// that observe the flag.  Bounded spin: a lost peer traps instead of hanging the box.
// ------------------------------------------------------------------------------------------
static __device__ unsigned int* g_trap_log_sparse = nullptr;
void set_trap_log_sparse(unsigned int* host_mapped) {
    cudaMemcpyToSymbol(g_trap_log_sparse, &host_mapped, sizeof(host_mapped));
}
__global__ void k_barrier(unsigned int* flags_local, unsigned int epoch, const __grid_constant__ PeerTable pt) {
    const int t = threadIdx.x;
    if (t >= pt.world) return;
    __threadfence_system();
    unsigned int* remote = peer_ptr(pt, t, flags_local) + pt.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const unsigned int* mine = flags_local + t;
    unsigned int seen = 0;
    for (unsigned long long spins = 0; spins < (1ull << 25); ++spins) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int)(seen - epoch) >= 0) return;
        __nanosleep(64);
    }
    unsigned int* l = g_trap_log_sparse;      // {magic, 0xBA221E2 = cross-GPU barrier, rank, peer, epoch, seen}
    if (l != nullptr && atomicCAS(l, 0u, 0xDAE0DEADu) == 0u) {
        l[1] = 0xBA221E2u; l[2] = pt.rank; l[3] = t; l[4] = epoch; l[5] = seen;
        __threadfence_system();
    }
    __trap();
}
void launch_barrier(unsigned int* flags_local, unsigned int epoch, const PeerTable& pt, cudaStream_t st) {
    k_barrier<<<1, 32, 0, st>>>(flags_local, epoch, pt);
}

// Force the module / functions to load now: with CUDA's lazy loading the FIRST launch of a kernel may
// synchronise the context, which would deadlock against a cross-GPU flag barrier already spinning.
void preload_sparse() {
    cudaFuncAttributes a;
    PRELOAD_KERNEL(k_coo_count);
    PRELOAD_KERNEL(k_row_scan);
    PRELOAD_KERNEL(k_coo_fill);
    PRELOAD_KERNEL(k_row_sort_dedup);
    PRELOAD_KERNEL(k_ybits_shard);
    PRELOAD_KERNEL(k_reduce_splits);
    PRELOAD_KERNEL(k_da_own);
    PRELOAD_KERNEL(k_gather_items_f32);
    PRELOAD_KERNEL(k_gather_rows_bf16);
    PRELOAD_KERNEL(k_encode_fwd);
    PRELOAD_KERNEL(k_transpose_hd);
    PRELOAD_KERNEL(k_colsum);
    PRELOAD_KERNEL(k_scatter_shard);
    PRELOAD_KERNEL(k_scatter_det);
    PRELOAD_KERNEL(k_touch_shard);
    PRELOAD_KERNEL(k_sum_partials);
    PRELOAD_KERNEL(k_barrier);
    (void)cudaGetLastError();
}

}  // namespace dae
