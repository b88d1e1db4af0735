/* dae_b200.h -- C ABI of the B200-native DAE hot path (libdae_b200.so).
 *
 * Drop-in boundary for the denoising-autoencoder path of hojinYang/spotify_recSys_challenge_2018.
 * The reference has no FFI: its boundary is the Python protocol between the runners and the TF1
 * graph objects of models/DAEs.py (`sess.run(fetches, feed_dict)`).  Each entry point below names
 * the reference call it replaces (file:line relative to the reference root).  All `const T*`
 * inputs of the model-level calls are HOST pointers in exactly the layout the reference's readers
 * produce (utils/data_reader.py): `pos` = int64 [nnz,2] (row-in-batch, item id), `val` = float32
 * [nnz].  No torch types appear anywhere in this ABI.
 *
 * Every function returns 0 on success, non-zero on error; dae_last_error() describes the last
 * error of the calling thread.  There is no CPU fallback: creation fails if no sm_100 device is
 * present.
 */
#ifndef DAE_B200_H
#define DAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAE_B200_ABI_VERSION 3

typedef struct dae_model dae_model; /* opaque: parameters, Adam state, workspaces, stream */

typedef struct dae_config {
    int32_t n_input;     /* conf.n_input  = tracks + artists            (DAEs.py:18)            */
    int32_t n_tracks;    /* conf.n_tracks: ranking uses [:, :n_tracks]  (main_train.py:86)       */
    int32_t n_hidden;    /* conf.hidden   (multiple of 64, <= 256)      (DAEs.py:19)            */
    int32_t max_batch;   /* conf.batch    (train: <= 256 per GPU)       (DAEs.py:17)            */
    int32_t tied;        /* 1 = DAE_tied (DAEs.py:13), 0 = DAE (DAEs.py:114)                     */
    float lr;            /* conf.lr                                      (DAEs.py:20, :102)      */
    float reg_lambda;    /* conf.reg_lambda                              (DAEs.py:21, :100)      */
    uint64_t seed;       /* Philox key of the dropout masks and of dae_model_init_xavier          */
    int32_t device;      /* CUDA device ordinal                                                    */
    int32_t trainable;   /* 1: trainable; 0: inference only (no Adam state / gradient buffers); 2: frozen but staged
                            with targets -- the constant DAE inside DAE_title while the title branch trains (DAEs.py:164-171) */
    void* stream;        /* cudaStream_t to run on, or NULL to create a private stream            */
    int32_t world;       /* data-parallel ranks (GPUs of one NVSwitch box, <= 8); 0 or 1 = single GPU       */
    int32_t rank;        /* this model's rank in [0, world)                                                  */
} dae_config;

int32_t dae_abi_version(void);
const char* dae_last_error(void);

/* DAE_tied(conf) / DAE(conf) + model.fit() + sess.run(init_op)      DAEs.py:14,84-105,115; main_train.py:171-173 */
int32_t dae_model_create(const dae_config* cfg, dae_model** out);
void dae_model_destroy(dae_model* m);

/* tf.contrib.layers.xavier_initializer / zeros_initializer           DAEs.py:54-59, :121-128 */
int32_t dae_model_init_xavier(dae_model* m, uint64_t seed);

/* pickle.load -> tf.constant initialisers / save_model: the four arrays of d_params, host fp32:
 * W_enc [n_input,n_hidden], W_dec [n_input,n_hidden] (ignored/aliased when tied), b_enc [n_hidden],
 * b_dec [n_input].                                                    DAEs.py:107-111, :129-138 */
int32_t dae_model_set_params(dae_model* m, const float* W_enc, const float* W_dec, const float* b_enc,
                             const float* b_dec);
int32_t dae_model_get_params(dae_model* m, float* W_enc, float* W_dec, float* b_enc, float* b_dec);
/* Adam moments + step counter (not saved by the reference; exposed for exact resume and tests) */
int32_t dae_model_get_adam_state(dae_model* m, float* m_W_enc, float* v_W_enc, float* m_W_dec, float* v_W_dec,
                                 int64_t* step);

/* sess.run([model.optimizer, model.cost], feed_dict={x_positions, x_ones, y_positions, y_ones,
 * keep_prob, input_keep_prob}) -> cost.  Synchronous, host buffers in, host scalar out.
 *                                                                     main_train.py:204-213 */
int32_t dae_model_train_step(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                             const int64_t* y_pos, const float* y_val, int64_t nnz_y, int32_t batch,
                             float keep_prob, float input_keep_prob, float* cost_out);

/* sess.run(model.y_pred, {x_positions, x_ones, keep_prob: 1, input_keep_prob: 1}) -> [batch, n_cols]
 * fp32 host matrix, n_cols = n_input or n_tracks (the runner slices [:, :n_tracks] anyway).
 *                                                                     main_train.py:66-68, :86 */
int32_t dae_model_predict(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x, int32_t batch,
                          int32_t n_cols, float* y_pred_out);

/* The same step, pipelined for a training loop (the reference's loop only accumulates the cost,
 * main_train.py:223): stages this batch while the previous step is still running, enqueues the step
 * and returns the PREVIOUS step's cost (*has_prev = 0 on the first call).  dae_model_train_flush
 * returns the last pending cost (epoch end / before reading parameters). */
int32_t dae_model_train_step_async(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                   const int64_t* y_pos, const float* y_val, int64_t nnz_y, int32_t batch,
                                   float keep_prob, float input_keep_prob, float* prev_cost_out, int32_t* has_prev);
int32_t dae_model_train_flush(dae_model* m, float* cost_out, int32_t* has_cost);

/* y_pred + np.argsort + seed removal + [:k] fused on the device: met.single_eval /
 * cand_generate.  seed_ptr [batch+1] / seed_idx: CSR of the seed track ids to exclude (host).
 * out_idx [batch,k] int32 (-1 padded), out_score [batch,k] fp32 (may be NULL).
 *                                                   metrics.py:58-68; main_challenge.py:26-36,80-90 */
int32_t dae_model_recommend(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x, int32_t batch,
                            const int32_t* seed_ptr, const int32_t* seed_idx, int32_t k, int32_t* out_idx,
                            float* out_score);
/* The same ranking restricted to the catalogue range [item_lo, item_hi) (clipped to the tracks): global ids, seeds
 * removed.  Item-sharded challenge inference gives every GPU one range and merges the per-shard lists with
 * dae_topk_merge_device; the merged list is exactly the unsharded one.  Ranges of >= 131072 items (and debug bit 4)
 * run the fused decode + top-K (threshold-filtered candidate lists, the [batch, T] score matrix never exists; batch
 * up to max_batch rows in 256-row tiles); smaller ranges, or a candidate list that overflowed, take the dense path.
 * out_idx == NULL leaves the lists on the device (buffers "topk_idx" / "topk_score", [batch, k]).
 *                                                   main_challenge.py:80-90 (y_pred[:, :n_tracks] + cand_generate) */
int32_t dae_model_recommend_range(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x, int32_t batch,
                                  const int32_t* seed_ptr, const int32_t* seed_idx, int32_t k, int32_t item_lo,
                                  int32_t item_hi, int32_t* out_idx, float* out_score);

/* dae_model_recommend followed by met.single_eval of every playlist ON THE DEVICE: answers as a CSR (ans_ptr [batch+1],
 * ans_idx; an answer list may contain -1 = a track outside the vocabulary: never matched, still counted), metrics_out
 * [batch, 3] doubles = (r-precision, ndcg with the reference's own IDCG, recommended-songs clicks).  Replaces the
 * runner's per-playlist Python loop over a [batch, 500] id matrix: 24 bytes per playlist come back instead of 2 KB.
 *                                             main_train.py:62-100; metrics.py:20-27 (r-precision), :29-42, :44-49 */
int32_t dae_model_evaluate(dae_model* m, const int64_t* x_pos, const float* x_val, int64_t nnz_x, int32_t batch,
                           const int32_t* seed_ptr, const int32_t* seed_idx, const int32_t* ans_ptr, const int32_t* ans_idx,
                           int32_t k, double* metrics_out);

/* ---- device-resident / asynchronous variants (bench `value`, data-parallel training) ---------- */

/* Copy one batch (host COO) into device staging slot 0 or 1 and build its CSR / target bitmask, all
 * asynchronously on the model's side stream (overlaps with a step running on the main stream). */
int32_t dae_model_stage_batch(dae_model* m, int32_t slot, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                              const int64_t* y_pos, const float* y_val, int64_t nnz_y, int32_t batch);
/* Re-run the device-side preparation (COO -> CSR, target bitmask) of the batch already resident in
 * `slot` on the side stream: with inputs kept in HBM every step still does all of its own work,
 * one step ahead of the main stream. */
int32_t dae_model_restage(dae_model* m, int32_t slot);
/* Forward + backward from a staged slot; gradients stay in device buffers; no host sync.
 * The loss is a mean over the GLOBAL batch (DAEs.py:100) and dropout is keyed by the global row, so
 * N ranks x B_local == one rank x (N*B_local).  global_batch = 0: world * batch, row_offset = rank * batch.
 * With world > 1 this enqueues a cross-GPU barrier first (all ranks must call it). */
int32_t dae_model_backward_staged(dae_model* m, int32_t slot, float keep_prob, float input_keep_prob,
                                  int32_t global_batch, int32_t row_offset);
/* dW_dec = dz^T h_d and the sparse-row dW_enc of the rows this rank owns, then the dense TF1 Adam
 * update of every variable and step += 1.                                                      DAEs.py:102 */
int32_t dae_model_apply_adam(dae_model* m);
/* backward_staged + apply_adam (single GPU). */
int32_t dae_model_train_step_staged(dae_model* m, int32_t slot, float keep_prob, float input_keep_prob);
/* Synchronise the stream, check the device-side error flag, return the last step's cost. */
int32_t dae_model_sync_cost(dae_model* m, float* cost_out);

/* ---- data parallelism over the GPUs of one box (the reference has none) --------------
 * One process (or model) per GPU, `world` of them.  The split is hybrid.  Sparse side, by playlist: rank r stages,
 * de-duplicates and encodes ITS rows of the global batch, gathering W_enc rows from the GPUs that own them, and stores
 * its h / h_d rows and its normalised input into every rank's copy of the global batch.  Dense side, by item: the
 * catalogue-sized state (W_enc, W_dec, Adam moments, the bf16 operand, dz, dW_dec) is row-sharded tile-cyclically, ZeRO
 * style, and every rank decodes, differentiates and updates its OWN item rows against the whole global batch -- that
 * state never crosses NVLink.  All exchange is plain loads / stores into the peers' arenas, issued by the kernels
 * that produce or consume the data; four flag barriers per step order it (A: previous step over everywhere, B1: the
 * global h_d has landed, B2: the split-K sums of dh have been stored into their row owners (a reduce-scatter by peer
 * stores) and every rank's db_dec rows / cost partial are complete, B3: every rank's da rows -- formed by the row owner
 * only and stored into every rank's copy, an all-gather by peer stores -- have arrived).  Every rank must stage
 * batches of the same size and call the step functions in the same order.  Attach once after dae_model_create on
 * every rank:
 *   dae_model_ipc_handle  -> 64-byte cudaIpcMemHandle_t of this rank's arena (exchange them out of band)
 *   dae_model_attach_ipc  <- all `world` handles in rank order (own slot ignored)
 *   dae_model_attach_local: peers living in the SAME process (tests; several models on one or more GPUs) */
int32_t dae_model_ipc_handle(dae_model* m, void* handle_out64);
int32_t dae_model_attach_ipc(dae_model* m, const void* handles, int32_t n_handles);
int32_t dae_model_attach_local(dae_model* m, dae_model* const* peers, int32_t n_peers);
int32_t dae_model_arena_bytes(dae_model* m, int64_t* bytes);

/* ---- item-sharded challenge inference (main_challenge.py:80-90 on one GPU upstream) -------------------
 * Every rank holds a replica of the inference model and ranks ITS slice of the track catalogue with
 * dae_model_recommend_range(..., out_idx = NULL) (lists stay on the device: buffers "topk_idx" / "topk_score").  A
 * dae_exchange merges the per-shard lists without a collective library: dae_exchange_merge_topk stores this rank's lists
 * into every peer's merge buffer (plain stores over NVLink through the CUDA IPC mapping), one flag barrier, then the
 * (score desc, id asc) merge of world x k candidates per playlist runs locally -- exactly the unsharded ranking, on every
 * rank.  Create one per rank, exchange the 64-byte handles out of band (as for the models), attach, then call
 * dae_exchange_merge_topk collectively (same batch and k on every rank). */
typedef struct dae_exchange dae_exchange;
int32_t dae_exchange_create(int32_t device, int32_t world, int32_t rank, int32_t max_batch, int32_t max_k, dae_exchange** out);
void dae_exchange_destroy(dae_exchange* x);
int32_t dae_exchange_ipc_handle(dae_exchange* x, void* handle_out64);
int32_t dae_exchange_attach_ipc(dae_exchange* x, const void* handles, int32_t n_handles);
int32_t dae_exchange_attach_local(dae_exchange* x, dae_exchange* const* peers, int32_t n_peers);
int32_t dae_exchange_merge_topk(dae_exchange* x, const int32_t* idx_dev, const float* score_dev, int32_t batch, int32_t k,
                                int32_t* out_idx, float* out_score, void* stream);
/* the merge itself sharded by playlist: rank r receives, merges and returns rows [*row_begin, *row_end) = its 1 / world of
 * the batch (ceil(batch / world) rows per rank); out_idx / out_score stay [batch, k] host arrays of which only those rows
 * are written.  (world - 1) / world of ONE list set leaves each rank instead of world - 1 copies of it, and the merge and
 * the read-back shrink by world: the call for a caller that writes its own rows of the submission. */
int32_t dae_exchange_merge_topk_rows(dae_exchange* x, const int32_t* idx_dev, const float* score_dev, int32_t batch, int32_t k,
                                     int32_t* row_begin, int32_t* row_end, int32_t* out_idx, float* out_score, void* stream);
/* Filter thresholds of the fused decode + top-K shared across the item shards: while an exchange is set,
 * dae_model_recommend_range on `m` is COLLECTIVE over the exchange's ranks (same batch, k and seeds everywhere; call it on
 * the stream later handed to dae_exchange_merge_topk*).  Every shard selects its ceil((k + max seeds) / world)-th largest
 * logit after a pass; the minimum over the shards (stored into every peer, one flag barrier) is exceeded by at least
 * k + max seeds items of the whole catalogue, so no member of the global top-k is filtered out, and the candidate lists of
 * the next pass are ~world times shorter than with per-shard thresholds.  x = NULL detaches. */
int32_t dae_model_set_threshold_exchange(dae_model* m, dae_exchange* x);
/* device times per call (ms, mean since switched on): out4 = stores, barrier, merge, read-back */
int32_t dae_exchange_set_profiling(dae_exchange* x, int32_t on);
int32_t dae_exchange_phase_ms(dae_exchange* x, float* out4);
int64_t dae_exchange_launch_count(dae_exchange* x);

/* debug / tuning flags.  bit 0: dae_model_backward_staged also forms dW_dec of the rows this rank owns
 * (buffer "g_dec") and bit 1: the sparse-row dW_enc scatter ("g_enc", "touched") -- both otherwise
 * happen inside dae_model_apply_adam, after the step's second barrier.  bit 2: apply_adam forms dW_dec
 * in HBM and runs the decoder's Adam update as a second kernel, instead of the default fused kernel
 * that applies Adam to the dW tile while it is still in tensor memory.  bit 3: dae_model_train_step_staged keeps
 * the decoder update on the main stream instead of overlapping it with the sparse / encoder tail of the step.
 * bit 4: dae_model_recommend[_range] always takes the fused decode + top-K path, bit 5: never.
 * bits 6 / 7 / 8: keep the target bitmask / the decoder update / the bias updates on the main stream (bisecting the
 * three forks of the whole-step call).  bit 9: with >= 4 GPUs keep the encoder Adam of the unlisted rows behind the decoder
 * update instead of next to the encode.  bit 10: stream the encoder Adam of the rows no playlist lists in the background,
 * under the front of the step (k_adam_bg; bit-identical results, measured slower: off by default).  bit 11: dW_enc
 * through fp32 red.add only (the multi-GPU default), bit 12: ordered gather for the rows listed by >= 3 playlists (the
 * single-GPU default: bit-reproducible steps).  bit 13: %globaltimer stamps of the step's fork / join points into the
 * buffer "trace" (tools/gpu_trace.py).  bit 14: the title branch forms dW_out in HBM ("g_W_out") and runs its Adam as a
 * second kernel instead of the fused one.  bit 15 (process-wide, experiment): inference batches of more than 256 rows
 * decode in 128-row tiles, clusters of two tiles sharing every W chunk through TMA multicast (measured slower than the
 * default 256-row tiles; kept for A/B).  bits 17 / 18 (process-wide, TIMING EXPERIMENTS ONLY -- the rankings are wrong):
 * the filter pass of the fused decode + top-K releases every accumulator unread / scans but never queues a row (the
 * floors quoted in DESIGN.md section 4, "FILTER epilogue"). */
int32_t dae_model_set_debug(dae_model* m, int32_t flags);

/* Named device buffers (pointer, element count, element size) for parity tests.  Catalogue-row
 * buffers hold the rows this rank owns in local-tile order (== global order when world == 1).  Names:
 * "g_dec" "g_enc" "g_b_enc" "g_b_dec" "g_b_enc_part" "g_b_dec_part" "touched" "cost" "W_enc" "W_dec"
 * "W_dec_bf16" "b_enc" "b_dec" "h" "h_d" "h_dT" "dzT" "dz_all" "dh_partial" "da" "x_row_ptr" "x_row_len"
 * "x_col" "x_val" "x_rowsum" "y_row_ptr" "y_row_len" "y_col" "ybits" "scores" "topk_idx" "topk_score" "mW_dec" "vW_dec" "mW_enc" "vW_enc"
 * "bg_ctl" "trace".  "touched" = the rows of W_enc the batch lists in x (known before the backward), cleared at the end of the step. */
int32_t dae_model_buffer(dae_model* m, const char* name, void** dev_ptr, int64_t* n_elem, int32_t* elem_size);
/* number of kernels launched by this model since creation (bench.py `gpu_launches`) */
int64_t dae_model_launch_count(dae_model* m);

/* Per-phase device timing (CUDA events on the model's stream around each phase of a train step;
 * a profiled step synchronises once at its end).  Used by bench.py for the roofline of the
 * dominant kernel; leave off for throughput runs. */
int32_t dae_model_set_profiling(dae_model* m, int32_t on);
int32_t dae_model_phase_count(void);
const char* dae_model_phase_name(int32_t k);
int32_t dae_model_phase_time(dae_model* m, int32_t k, double* total_ms, int64_t* count);

/* ---- title branch: character CNN + output layer on top of a constant DAE --------------------------
 * models/title_get.py:10-22 (get_model), models/title_models/Char_CNN.py:5-75 (Char_CNN) and
 * models/DAEs.py:153-201 (DAE_title: y_pred = title_score * w_title + sigmoid(decoder) * w_playlist,
 * weighted BCE on y_pred, only the title variables train).  The DAE model is created by the caller
 * with trainable = 2 (training the title branch) or 0 (inference) and loaded from conf.DAEval with
 * dae_model_set_params (DAEs.py:164-171); it must outlive the title object.  Batches of <= 256 rows. */
typedef struct dae_title dae_title;
typedef struct dae_title_config {
    int32_t charsize;          /* conf.charsize   (Char_CNN.py:11)  */
    int32_t strmaxlen;         /* conf.strmaxlen  (Char_CNN.py:9), <= 32 */
    int32_t char_emb;          /* conf.char_emb   (Char_CNN.py:8), > 0 */
    int32_t filter_num;        /* conf.filter_num (title_get.py:20) */
    int32_t n_filter_sizes;    /* len(conf.filter_size), <= 8; filter_num * n_filter_sizes <= 512 */
    int32_t filter_size[8];    /* conf.filter_size (title_get.py:19) */
    float lr;                  /* [TITLE] lr (main.py:60) */
    int32_t trainable;         /* 0: inference only */
} dae_title_config;
int32_t dae_title_create(dae_model* constant_dae, const dae_title_config* cfg, dae_title** out);
void dae_title_destroy(dae_title* t);
/* xavier_initializer(uniform=False) on every title variable                 Char_CNN.py:19, :45-47, :71-73 */
int32_t dae_title_init(dae_title* t, uint64_t seed);
/* Host fp32 arrays in the order [char_embedding, Conv_W0, Conv_b0, ..., Output_W, Output_b] with the
 * reference's shapes: [charsize, char_emb]; [fs_i, char_emb, filter_num]; [filter_num]; [D, n_output]; [n_output]. */
int32_t dae_title_param_count(dae_title* t);
int32_t dae_title_param_size(dae_title* t, int32_t idx, int64_t* n_elem);
int32_t dae_title_set_params(dae_title* t, const float* const* arrays);
int32_t dae_title_get_params(dae_title* t, float* const* arrays);
/* sess.run([optimizer, cost], {x, y, titles, keep_prob, title keep_prob, input_keep_prob, titles_use}) -> cost.
 * titles: int64 [batch, strmaxlen] char ids (-1 = pad); titles_use: fp32 [batch].        main_train.py:214-221 */
int32_t dae_title_train_step(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                             const int64_t* y_pos, const float* y_val, int64_t nnz_y, const int64_t* titles,
                             const float* titles_use, int32_t batch, float keep_prob, float input_keep_prob,
                             float title_keep_prob, float* cost_out);
/* The same step pipelined like dae_model_train_step_async: stages this batch while the previous step runs, returns the
 * PREVIOUS step's cost (*has_prev = 0 on the first call); dae_title_train_flush returns the last pending one. */
int32_t dae_title_train_step_async(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                   const int64_t* y_pos, const float* y_val, int64_t nnz_y, const int64_t* titles,
                                   const float* titles_use, int32_t batch, float keep_prob, float input_keep_prob,
                                   float title_keep_prob, float* prev_cost_out, int32_t* has_prev);
int32_t dae_title_train_flush(dae_title* t, float* cost_out, int32_t* has_cost);
/* sess.run(y_pred, {..., titles, titles_use, keep probabilities 1}) -> [batch, n_cols] fp32.
 *                                                                   main_train.py:69-79; main_challenge.py:80-85 */
int32_t dae_title_predict(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x, const int64_t* titles,
                          const float* titles_use, int32_t batch, int32_t n_cols, float* y_pred_out);
/* y_pred[:, :n_tracks] + cand_generate on the device.                       main_challenge.py:26-36, :87-90 */
int32_t dae_title_recommend(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                            const int64_t* titles, const float* titles_use, int32_t batch, const int32_t* seed_ptr,
                            const int32_t* seed_idx, int32_t k, int32_t* out_idx, float* out_score);
/* dae_title_recommend + the metrics of every list on the device (see dae_model_evaluate).       main_train.py:69-100 */
int32_t dae_title_evaluate(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x, const int64_t* titles,
                           const float* titles_use, int32_t batch, const int32_t* seed_ptr, const int32_t* seed_idx,
                           const int32_t* ans_ptr, const int32_t* ans_idx, int32_t k, double* metrics_out);
int64_t dae_title_launch_count(dae_title* t);
/* named device buffers for parity tests ("W_out", its moments and "g_W_out" are two dense column blocks, [Np, h0]
 * followed by [Np, h1], Np = n_output rounded up to 128, h0 = min(256, D rounded up to 64), h1 = D - 256 rounded up to 64; the bf16 operand copy
 * "W_out_bf16" is [n_output, 512]): "feat" "argpos" "feat_d" "w_t" "w_p" "dzT" "d" "g_emb" "g_conv_W"
 * "g_conv_b" "g_W_out" (debug bit 14 only) "g_b_out" "W_out" "W_out_bf16" "m_W_out" "v_W_out" */
int32_t dae_title_buffer(dae_title* t, const char* name, void** dev_ptr, int64_t* n_elem, int32_t* elem_size);

/* ---- kernel-level entry points on caller-owned DEVICE memory (parity tests, other hosts) ------- */

/* met.single_eval ranking on an existing score matrix.              metrics.py:58-68 */
int32_t dae_topk_device(const float* scores_dev, int64_t ld, int32_t batch, int32_t n_tracks, int32_t k,
                        const int32_t* seed_ptr_dev, const int32_t* seed_idx_dev, int32_t idx_base,
                        int32_t* out_idx_dev, float* out_score_dev, void* stream);
/* met.get_r_precision / get_ndcg / get_rsc of ranked lists already on the device: cand_dev [batch, ld] int32 (k ranks used,
 * -1 padded), answers CSR on the device, out_dev [batch, 3] doubles.              metrics.py:20-27, :29-42, :44-49 */
int32_t dae_metrics_device(const int32_t* cand_dev, int64_t ld, int32_t batch, int32_t k, const int32_t* ans_ptr_dev,
                           const int32_t* ans_idx_dev, double* out_dev, void* stream);
/* merge of per-shard top-k lists: row r = n (score, global id) pairs, -inf / -1 padded -> first k by (score desc, id asc) */
int32_t dae_topk_merge_device(const float* scores_dev, const int32_t* idx_dev, int32_t n, int32_t batch, int32_t k,
                              int32_t* out_idx_dev, float* out_score_dev, void* stream);
/* ApplyAdam on one variable.                                         DAEs.py:102 [TF1] */
int32_t dae_adam_device(float* w_dev, float* m_dev, float* v_dev, const float* g_dev, uint16_t* w_bf16_dev,
                        int64_t n, float lr, float beta1_power, float beta2_power, float reg_lambda, void* stream);
/* tf.sparse_tensor_to_dense semantics kept sparse: COO -> per-row unique sorted columns, last value wins.
 * row_ptr_dev [batch+1] raw offsets, row_len_dev [batch], col_dev/val_dev [nnz].     DAEs.py:33-38 */
int32_t dae_coo_to_csr_device(const int64_t* pos_dev, const float* val_dev, int64_t nnz, int32_t batch,
                              int32_t n_input, int32_t* row_ptr_dev, int32_t* row_len_dev, int32_t* col_dev,
                              float* val_out_dev, void* stream);
/* The three tensor-core contractions on caller-owned bf16 operands (descriptor / layout tests):
 *   op 0: out[b, item] = sigmoid(W[item,:].h_d[b,:] + bias[item])          (decode, DAEs.py:75)
 *   op 1: out[item, :] = sum_b dzT[item,b] h_dT[:,b]                        (dW_dec)
 *   op 2: out[split, b, :] = partial sums of sum_item dzT[item,b] W[item,:] (dh, split-K)
 * lbo/sbo: MN-major descriptor strides for op 2 (0 = defaults). */
int32_t dae_dh_nsplit(int32_t n_items); /* number of split-K partials op 2 writes */
int32_t dae_gemm_test_device(int32_t op, const uint16_t* a_dev, const uint16_t* b_dev, const float* bias_dev,
                             float* out_dev, int32_t n_items, int32_t n_hidden, int32_t batch, int32_t bpad,
                             int32_t lbo, int32_t sbo, int32_t* nsplit_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAE_B200_H */
