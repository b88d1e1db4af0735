// C-ABI of the title branch (include/dae_b200.h, "title branch"): the character CNN + output layer
// trained on top of a constant DAE (models/DAEs.py:153-201, models/title_models/Char_CNN.py:16-75),
// and the mixed prediction used by --title evaluation and --challenge inference
// (main_runner/main_train.py:69-79, main_runner/main_challenge.py:80-90).
#include <math.h>

#include "model.h"
#include "philox.cuh"

#define TRY(x) do { if (int rc_ = (x)) return rc_; } while (0)

struct dae_title {
    dae_model* dae = nullptr;          // constant DAE (trainable 0 or 2); owned by the caller
    dae_title_config cfg{};
    CnnShape shape{};
    int D = 0, N = 0, H = 0, Bmax = 0, n_conv_w = 0, nsplit = 0;
    // The output layer's master / moments are kept as TWO DENSE column blocks, [N, h0] (feature columns 0..255) followed by
    // [N, h1] (columns 256..D-1 rounded up to 64: D = 400 -> 256 + 192), block b at offset blk_off[b]: the fused dW + Adam
    // kernel streams dense row groups with 1-D bulk copies and the dead columns of a [N, 512] layout are never touched.
    // The bf16 tensor-core operand (W_out_bf16) stays [N, 512].
    int hb[2] = {0, 0};
    size_t blk_off[2] = {0, 0};
    bool trainable = true;
    cudaStream_t st = nullptr;
    cudaStream_t st_cnn = nullptr;     // the CNN's backward + small Adam updates run here, under the output layer's HBM-bound update
    cudaEvent_t ev_dfeat = nullptr, ev_cnn = nullptr;
    // variables (fp32 masters), TF1-Adam moments, gradients
    float *emb = nullptr, *conv_W = nullptr, *conv_b = nullptr, *W_out = nullptr, *b_out = nullptr;
    float *m_emb = nullptr, *m_conv_W = nullptr, *m_conv_b = nullptr, *m_W_out = nullptr, *m_b_out = nullptr;
    float *v_emb = nullptr, *v_conv_W = nullptr, *v_conv_b = nullptr, *v_W_out = nullptr, *v_b_out = nullptr;
    float *g_emb = nullptr, *g_conv_W = nullptr, *g_conv_b = nullptr, *g_W_out = nullptr, *g_b_out = nullptr;
    __nv_bfloat16* W_out_bf16 = nullptr;          // [N, 512] item-major operand copy
    float b1_pow = kBeta1, b2_pow = kBeta2;
    long long step = 0;
    // activations / workspaces
    long long *titles = nullptr, *h_titles = nullptr;      // slot of the call in flight (points into the pairs below)
    float *titles_use = nullptr, *h_titles_use = nullptr;
    long long *titles2[2] = {nullptr, nullptr}, *h_titles2[2] = {nullptr, nullptr};   // per staging slot
    float *titles_use2[2] = {nullptr, nullptr}, *h_titles_use2[2] = {nullptr, nullptr};
    cudaEvent_t ev_titles[2] = {nullptr, nullptr};        // the slot's pinned mirrors have been copied to the device
    cudaEvent_t ev_cost[2] = {nullptr, nullptr};
    float* cost_ring = nullptr;                            // pinned [2]
    int* err_ring = nullptr;                               // pinned [2]
    int async_slot = 1;
    bool async_pending = false;
    float *dx = nullptr, *conv_WT = nullptr;
    float *feat = nullptr, *d = nullptr, *w_t = nullptr, *w_p = nullptr, *dh_partial = nullptr, *scratch = nullptr;
    unsigned char* argpos = nullptr;
    __nv_bfloat16 *feat_d = nullptr, *feat_dT = nullptr, *dzT = nullptr;
    float *loss_partial = nullptr, *cost = nullptr, *cost_host = nullptr;
    int last_batch = 0;
    long long launches = 0;
    std::vector<void*> dev_allocs, host_allocs;
};

template <typename T>
static int dalloc(dae_title* t, T** p, size_t n) {
    CK(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T) + 16));
    t->dev_allocs.push_back(*p);
    CK(cudaMemsetAsync(*p, 0, n * sizeof(T), t->st));
    return 0;
}

extern "C" int32_t dae_title_create(dae_model* dae, const dae_title_config* cfg, dae_title** out) {
    if (!dae || !cfg || !out) return fail("null argument");
    *out = nullptr;
    if (dae->trainable) return fail("the DAE under a title model is a constant (create it with trainable = 0 or 2)");
    if (dae->world != 1) return fail("the title branch runs on one GPU (world = 1)");
    if (cfg->n_filter_sizes <= 0 || cfg->n_filter_sizes > kTitleMaxWidths) return fail("1..%d filter sizes", kTitleMaxWidths);
    if (cfg->strmaxlen <= 0 || cfg->strmaxlen > kTitleMaxLen) return fail("strmaxlen must be in [1,%d]", kTitleMaxLen);
    if (cfg->char_emb <= 0) return fail("char_emb must be > 0 (the one-hot variant is not used by any shipped config)");
    if (cfg->filter_num <= 0 || cfg->filter_num * cfg->n_filter_sizes > kTitleFpad)
        return fail("filter_num * len(filter_size) must be in [1,%d]", kTitleFpad);
    if (dae->Bmax > kMaxBpad) return fail("the title branch handles batches of at most %d rows", kMaxBpad);
    ensure_loaded();
    dae_title* t = new dae_title();
    t->dae = dae; t->cfg = *cfg; t->st = dae->st;
    CK(cudaStreamCreateWithFlags(&t->st_cnn, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&t->ev_dfeat, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&t->ev_cnn, cudaEventDisableTiming));
    t->N = dae->N; t->H = dae->H; t->Bmax = dae->Bmax;
    t->trainable = cfg->trainable != 0;
    if (t->trainable && !dae->needs_y) { delete t; return fail("training the title branch needs a DAE created with trainable = 2"); }
    CnnShape& s = t->shape;
    s.C = cfg->charsize; s.L = cfg->strmaxlen; s.E = cfg->char_emb; s.F = cfg->filter_num; s.n_widths = cfg->n_filter_sizes;
    int off = 0, maxw = 0;
    for (int i = 0; i < s.n_widths; ++i) {
        const int w = cfg->filter_size[i];
        if (w <= 0 || w > s.L) { delete t; return fail("filter size %d outside [1, strmaxlen]", w); }
        s.width[i] = w; s.w_off[i] = off; off += w * s.E * s.F;
        maxw = w > maxw ? w : maxw;
    }
    if (maxw * s.E > 512) { delete t; return fail("filter size x char_emb must be <= 512"); }
    if (s.E > 128) { delete t; return fail("char_emb must be <= 128"); }
    t->n_conv_w = off;
    t->D = s.F * s.n_widths;
    t->hb[0] = t->D >= 256 ? 256 : round_up(t->D, 64);
    t->hb[1] = t->D > 256 ? round_up(t->D - 256, 64) : 0;
    // (rows padded to whole 128-item tiles: the fused kernel moves the last tile's row groups as a whole)
    t->blk_off[0] = 0; t->blk_off[1] = (size_t)round_up(dae->N, kTileItems) * t->hb[0];
    const size_t NF = (size_t)t->N * kTitleFpad;
    const size_t NB = (size_t)round_up(t->N, kTileItems) * (t->hb[0] + t->hb[1]);       // the two dense column blocks
    const int B = t->Bmax, D = t->D, N = t->N;
    TRY(dalloc(t, &t->emb, (size_t)s.C * s.E)); TRY(dalloc(t, &t->conv_W, off)); TRY(dalloc(t, &t->conv_b, D));
    TRY(dalloc(t, &t->W_out, NB)); TRY(dalloc(t, &t->b_out, N)); TRY(dalloc(t, &t->W_out_bf16, NF));
    TRY(dalloc(t, &t->scratch, NF));                                  // [D, N] <-> [N, 512] transposes; dW_out during training
    if (t->trainable) {
        TRY(dalloc(t, &t->m_emb, (size_t)s.C * s.E)); TRY(dalloc(t, &t->v_emb, (size_t)s.C * s.E));
        TRY(dalloc(t, &t->m_conv_W, off)); TRY(dalloc(t, &t->v_conv_W, off));
        TRY(dalloc(t, &t->m_conv_b, D)); TRY(dalloc(t, &t->v_conv_b, D));
        TRY(dalloc(t, &t->m_W_out, NB)); TRY(dalloc(t, &t->v_W_out, NB));
        TRY(dalloc(t, &t->m_b_out, N)); TRY(dalloc(t, &t->v_b_out, N));
        TRY(dalloc(t, &t->g_emb, (size_t)s.C * s.E)); TRY(dalloc(t, &t->g_conv_W, off)); TRY(dalloc(t, &t->g_conv_b, D));
        TRY(dalloc(t, &t->g_b_out, N));
        t->g_W_out = t->scratch;
        TRY(dalloc(t, &t->dzT, (size_t)N * kMaxBpad));
        t->nsplit = dh_nsplit(N);
        TRY(dalloc(t, &t->dh_partial, (size_t)2 * t->nsplit * kMaxBpad * 256));
        TRY(dalloc(t, &t->d, (size_t)B * D));
        TRY(dalloc(t, &t->dx, (size_t)B * s.L * s.E));
        TRY(dalloc(t, &t->conv_WT, off));
        TRY(dalloc(t, &t->loss_partial, 148 * 2));
        TRY(dalloc(t, &t->cost, 1));
    }
    for (int k = 0; k < 2; ++k) {
        TRY(dalloc(t, &t->titles2[k], (size_t)B * s.L)); TRY(dalloc(t, &t->titles_use2[k], B));
        CK(cudaMallocHost(reinterpret_cast<void**>(&t->h_titles2[k]), sizeof(long long) * B * s.L)); t->host_allocs.push_back(t->h_titles2[k]);
        CK(cudaMallocHost(reinterpret_cast<void**>(&t->h_titles_use2[k]), sizeof(float) * B)); t->host_allocs.push_back(t->h_titles_use2[k]);
        CK(cudaEventCreateWithFlags(&t->ev_titles[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&t->ev_cost[k], cudaEventDisableTiming));
    }
    t->titles = t->titles2[0]; t->titles_use = t->titles_use2[0]; t->h_titles = t->h_titles2[0]; t->h_titles_use = t->h_titles_use2[0];
    CK(cudaMallocHost(reinterpret_cast<void**>(&t->cost_ring), sizeof(float) * 2)); t->host_allocs.push_back(t->cost_ring);
    CK(cudaMallocHost(reinterpret_cast<void**>(&t->err_ring), sizeof(int) * 2)); t->host_allocs.push_back(t->err_ring);
    TRY(dalloc(t, &t->feat, (size_t)B * D)); TRY(dalloc(t, &t->argpos, (size_t)B * D));
    TRY(dalloc(t, &t->feat_d, (size_t)kMaxBpad * kTitleFpad)); TRY(dalloc(t, &t->feat_dT, (size_t)kTitleFpad * kMaxBpad));
    TRY(dalloc(t, &t->w_t, kMaxBpad)); TRY(dalloc(t, &t->w_p, kMaxBpad));
    CK(cudaMallocHost(reinterpret_cast<void**>(&t->cost_host), sizeof(float))); t->host_allocs.push_back(t->cost_host);
    CK(cudaStreamSynchronize(t->st));
    *out = t;
    return 0;
}

extern "C" void dae_title_destroy(dae_title* t) {
    if (!t) return;
    cudaStreamSynchronize(t->st);
    if (t->st_cnn) { cudaStreamSynchronize(t->st_cnn); cudaStreamDestroy(t->st_cnn); }
    if (t->ev_dfeat) cudaEventDestroy(t->ev_dfeat);
    if (t->ev_cnn) cudaEventDestroy(t->ev_cnn);
    for (int k = 0; k < 2; ++k) {
        if (t->ev_titles[k]) cudaEventDestroy(t->ev_titles[k]);
        if (t->ev_cost[k]) cudaEventDestroy(t->ev_cost[k]);
    }
    for (void* p : t->dev_allocs) cudaFree(p);
    for (void* p : t->host_allocs) cudaFreeHost(p);
    delete t;
}

// Parameter arrays in the order [emb, conv_W0, conv_b0, ..., out_W, out_b] with the reference's shapes:
// emb [charsize, char_emb]; conv_Wi [fs_i, char_emb, filter_num] (= [fs, E, 1, F]); conv_bi [filter_num];
// out_W [D, n_output]; out_b [n_output].                              Char_CNN.py:20, :45-47, :72-73
extern "C" int32_t dae_title_param_count(dae_title* t) { return t ? 3 + 2 * t->shape.n_widths : 0; }
extern "C" int32_t dae_title_param_size(dae_title* t, int32_t idx, int64_t* n_elem) {
    if (!t || !n_elem) return fail("null argument");
    const CnnShape& s = t->shape;
    const int n = s.n_widths;
    if (idx == 0) *n_elem = (int64_t)s.C * s.E;
    else if (idx >= 1 && idx <= 2 * n) *n_elem = ((idx - 1) % 2 == 0) ? (int64_t)s.width[(idx - 1) / 2] * s.E * s.F : s.F;
    else if (idx == 2 * n + 1) *n_elem = (int64_t)t->D * t->N;
    else if (idx == 2 * n + 2) *n_elem = t->N;
    else return fail("parameter index %d out of range", idx);
    return 0;
}

static void refresh_out_shadow(dae_title* t) {
    for (int b = 0; b < 2; ++b)
        launch_cast_block_bf16(t->W_out + t->blk_off[b], t->W_out_bf16, t->N, t->hb[b], kTitleFpad, 256 * b, t->st);
    t->launches += 2;
}
// [D, N] host layout (Char_CNN.py:72) <-> the two dense column blocks
static void out_layer_transpose(dae_title* t, float* dn, float* blocks, bool to_blocks) {
    for (int b = 0; b < 2; ++b) {
        if (t->hb[b] == 0) continue;
        const int d_live = b == 0 ? (t->D < 256 ? t->D : 256) : t->D - 256;
        if (to_blocks) launch_transpose_pad(dn + (size_t)256 * b * t->N, blocks + t->blk_off[b], d_live, t->N, t->hb[b], 1, t->st);
        else launch_transpose_pad(blocks + t->blk_off[b], dn + (size_t)256 * b * t->N, d_live, t->N, t->hb[b], 0, t->st);
    }
}

extern "C" int32_t dae_title_set_params(dae_title* t, const float* const* arrays) {
    if (!t || !arrays) return fail("null argument");
    const CnnShape& s = t->shape;
    CK(cudaMemcpyAsync(t->emb, arrays[0], sizeof(float) * s.C * s.E, cudaMemcpyHostToDevice, t->st));
    for (int i = 0; i < s.n_widths; ++i) {
        CK(cudaMemcpyAsync(t->conv_W + s.w_off[i], arrays[1 + 2 * i], sizeof(float) * s.width[i] * s.E * s.F, cudaMemcpyHostToDevice, t->st));
        CK(cudaMemcpyAsync(t->conv_b + i * s.F, arrays[2 + 2 * i], sizeof(float) * s.F, cudaMemcpyHostToDevice, t->st));
    }
    CK(cudaMemcpyAsync(t->scratch, arrays[1 + 2 * s.n_widths], sizeof(float) * (size_t)t->D * t->N, cudaMemcpyHostToDevice, t->st));
    out_layer_transpose(t, t->scratch, t->W_out, true);
    CK(cudaMemcpyAsync(t->b_out, arrays[2 + 2 * s.n_widths], sizeof(float) * t->N, cudaMemcpyHostToDevice, t->st));
    refresh_out_shadow(t);
    if (t->conv_WT) launch_conv_transpose(t->conv_W, t->conv_WT, t->shape, t->st);
    CK(cudaStreamSynchronize(t->st));
    return 0;
}

extern "C" int32_t dae_title_get_params(dae_title* t, float* const* arrays) {
    if (!t || !arrays) return fail("null argument");
    const CnnShape& s = t->shape;
    CK(cudaMemcpyAsync(arrays[0], t->emb, sizeof(float) * s.C * s.E, cudaMemcpyDeviceToHost, t->st));
    for (int i = 0; i < s.n_widths; ++i) {
        CK(cudaMemcpyAsync(arrays[1 + 2 * i], t->conv_W + s.w_off[i], sizeof(float) * s.width[i] * s.E * s.F, cudaMemcpyDeviceToHost, t->st));
        CK(cudaMemcpyAsync(arrays[2 + 2 * i], t->conv_b + i * s.F, sizeof(float) * s.F, cudaMemcpyDeviceToHost, t->st));
    }
    out_layer_transpose(t, t->scratch, t->W_out, false);
    CK(cudaMemcpyAsync(arrays[1 + 2 * s.n_widths], t->scratch, sizeof(float) * (size_t)t->D * t->N, cudaMemcpyDeviceToHost, t->st));
    CK(cudaMemcpyAsync(arrays[2 + 2 * s.n_widths], t->b_out, sizeof(float) * t->N, cudaMemcpyDeviceToHost, t->st));
    CK(cudaStreamSynchronize(t->st));
    return 0;
}

// xavier_initializer(uniform=False) on every title variable, biases included (Char_CNN.py:19, :45-47, :71-73)
extern "C" int32_t dae_title_init(dae_title* t, uint64_t seed) {
    if (!t) return fail("null model");
    const CnnShape& s = t->shape;
    auto sd = [](double fan_in, double fan_out) { return (float)sqrt(2.6 / (fan_in + fan_out)); };
    unsigned stream = kStreamInit + 8;
    launch_trunc_normal(t->emb, s.C, s.E, s.E, sd(s.C, s.E), seed, stream++, t->st);
    for (int i = 0; i < s.n_widths; ++i) {
        const double rf = (double)s.width[i] * s.E;                                       // receptive field x in-channels (1)
        launch_trunc_normal(t->conv_W + s.w_off[i], (long long)s.width[i] * s.E, s.F, s.F, sd(rf, rf * s.F), seed, stream++, t->st);
        launch_trunc_normal(t->conv_b + i * s.F, 1, s.F, s.F, sd(s.F, s.F), seed, stream++, t->st);
    }
    for (int b = 0; b < 2; ++b) {                                                                  // padding columns stay zero
        const int d_live = b == 0 ? (t->D < 256 ? t->D : 256) : t->D - 256;
        if (t->hb[b] > 0)
            launch_trunc_normal(t->W_out + t->blk_off[b], t->N, t->D, t->hb[b], sd(t->D, t->N), seed, stream, t->st, 256 * b, d_live);
    }
    ++stream;
    launch_trunc_normal(t->b_out, 1, t->N, t->N, sd(t->N, t->N), seed, stream++, t->st);
    refresh_out_shadow(t);
    if (t->conv_WT) launch_conv_transpose(t->conv_W, t->conv_WT, t->shape, t->st);
    t->launches += 3 + 2 * s.n_widths;
    CK(cudaStreamSynchronize(t->st));
    return 0;
}

// stage the batch, run the constant DAE's encoder and the character CNN; leaves h_d, feat_d, w_t, w_p ready
static int title_forward(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x, const int64_t* y_pos,
                         const float* y_val, int64_t nnz_y, const int64_t* titles, const float* titles_use, int32_t batch,
                         float kp, float kp_in, float kp_t, bool with_y, int slot = 0) {
    dae_model* m = t->dae;
    if (!titles || !titles_use) return fail("null titles");
    if (batch <= 0 || batch > t->Bmax) return fail("batch %d outside (0, %d]", batch, t->Bmax);
    if (!(kp > 0.f) || !(kp_in > 0.f) || !(kp_t > 0.f)) return fail("keep probabilities must be > 0");
    TRY(stage_impl(m, slot, x_pos, x_val, nnz_x, y_pos, y_val, nnz_y, batch, with_y));
    const CnnShape& s = t->shape;
    t->titles = t->titles2[slot]; t->titles_use = t->titles_use2[slot];
    t->h_titles = t->h_titles2[slot]; t->h_titles_use = t->h_titles_use2[slot];
    CK(cudaEventSynchronize(t->ev_titles[slot]));         // this slot's pinned title mirrors have left for the device
    memcpy(t->h_titles, titles, sizeof(long long) * batch * s.L);
    memcpy(t->h_titles_use, titles_use, sizeof(float) * batch);
    // (on the main stream: the slot's device copies are only overwritten after the step that read them two calls ago)
    CK(cudaMemcpyAsync(t->titles, t->h_titles, sizeof(long long) * batch * s.L, cudaMemcpyHostToDevice, t->st));
    CK(cudaMemcpyAsync(t->titles_use, t->h_titles_use, sizeof(float) * batch, cudaMemcpyHostToDevice, t->st));
    CK(cudaEventRecord(t->ev_titles[slot], t->st));
    const int bpad = round_up(batch, 64);
    if (t->last_batch != batch) {                         // padding rows / columns of the operand copies must read zero
        CK(cudaMemsetAsync(t->feat_d, 0, sizeof(__nv_bfloat16) * (size_t)kMaxBpad * kTitleFpad, t->st));
        CK(cudaMemsetAsync(t->feat_dT, 0, sizeof(__nv_bfloat16) * (size_t)kTitleFpad * kMaxBpad, t->st));
    }
    m->step = t->step;                                    // dropout masks are keyed by the title model's step
    CK(cudaStreamWaitEvent(t->st, m->slots[slot].prepared, 0));
    run_encode(m, slot, bpad, bpad, kp, kp_in, 0, false);
    if (with_y) build_ybits(m, slot, batch, bpad);
    CnnFwdArgs c{};
    c.titles = t->titles; c.emb = t->emb; c.conv_W = t->conv_W; c.conv_b = t->conv_b; c.shape = s;
    c.feat = t->feat; c.argpos = t->argpos; c.feat_d = t->feat_d; c.feat_dT = t->feat_dT; c.B = batch; c.bpad = bpad;
    c.kp_t = kp_t; c.seed = m->cfg.seed; c.step = (unsigned long long)t->step; c.row_offset = 0;
    launch_charcnn_fwd(c, t->st);
    launch_mix_weights(m->rowsum, t->titles_use, kp_in, batch, bpad, t->w_t, t->w_p, t->st);
    t->launches += 2;
    t->last_batch = batch;
    return 0;
}

static TitleTileArgs tile_args(dae_title* t, int batch) {
    dae_model* m = t->dae;
    TitleTileArgs a{};
    a.W_dec = m->shadow_full; a.h_d = m->h_d; a.b_dec = m->b_dec;
    a.W_out = t->W_out_bf16; a.feat_d = t->feat_d; a.b_out = t->b_out; a.w_t = t->w_t; a.w_p = t->w_p;
    a.N = t->N; a.H = t->H; a.batch = batch; a.bpad = round_up(batch, 64); a.kf = (t->D + 63) / 64;
    return a;
}

// sess.run([model.optimizer, model.cost], {x, y, titles, keep_prob, title keep_prob, input_keep_prob, titles_use})
//                                                                        main_train.py:214-221
extern "C" int32_t dae_title_train_flush(dae_title* t, float* cost_out, int32_t* has_cost);
// one title-mode train step from staging slot `slot`; sync: wait for it and return its cost, else leave cost / error flag
// in the slot's ring entries behind ev_cost[slot]
static int title_train_impl(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                            const int64_t* y_pos, const float* y_val, int64_t nnz_y, const int64_t* titles,
                            const float* titles_use, int32_t batch, float keep_prob, float input_keep_prob,
                            float title_keep_prob, int slot, bool sync, float* cost_out) {
    dae_model* m = t->dae;
    TRY(title_forward(t, x_pos, x_val, nnz_x, y_pos, y_val, nnz_y, titles, titles_use, batch, keep_prob, input_keep_prob,
                      title_keep_prob, true, slot));
    const int bpad = round_up(batch, 64), N = t->N;
    const Slot& sl = m->slots[slot];
    TitleTileArgs a = tile_args(t, batch);
    a.ybits = m->ybits; a.ywords = bpad / 32; a.dzT = t->dzT; a.db_out = t->g_b_out; a.loss_partial = t->loss_partial;
    a.inv_batch = 1.0f / (float)batch;
    ph_begin(m, PH_T_FWD, t->st);
    launch_title_train(a, t->st);
    ph_end(m, PH_T_FWD, t->st);
    CK(cudaEventRecord(sl.consumed, t->st));
    launch_reduce_loss2(t->loss_partial, decode_grid(N, 1), nullptr, 0, 0.f, 1.0f / (float)batch, t->cost, t->st);   // no l2 term (DAEs.py:196)
    t->launches += 2;

    // dense TF1 Adam on every title variable (DAEs.py:198; the DAE's are constants)
    AdamArgs ad{};
    ad.alpha = t->cfg.lr * sqrtf(1.0f - t->b2_pow) / (1.0f - t->b1_pow);
    ad.one_minus_b1 = 1.0f - kBeta1; ad.one_minus_b2 = 1.0f - kBeta2; ad.eps = kAdamEps; ad.lambda = 0.f;
    ad.row_touched = nullptr; ad.w_bf16 = nullptr;

    // d cost / d feat_d = dz_t . W_out FIRST (it reads the operand copy the update below rewrites), as two 256-column halves
    ph_begin(m, PH_T_DFEAT, t->st);
    for (int half = 0; half < 2; ++half) {
        DhArgs q{};
        q.dzT = t->dzT; q.W = t->W_out_bf16 + half * 256; q.ldW = kTitleFpad; q.N = N; q.H = 256; q.bpad = bpad;
        q.nsplit = t->nsplit; q.partial = t->dh_partial + (size_t)half * t->nsplit * bpad * 256;
        launch_dh(q, t->st);
        t->launches += 1;
    }
    ph_end(m, PH_T_DFEAT, t->st);
    // fork: everything behind d cost / d feat_d (CNN backward, the small Adam updates) is latency-bound CUDA-core work that
    // runs on its own stream UNDER the output layer's HBM-bound update; profiled steps keep it on the main stream
    cudaStream_t sc = m->profiling ? t->st : t->st_cnn;
    if (sc != t->st) {
        CK(cudaEventRecord(t->ev_dfeat, t->st));
        CK(cudaStreamWaitEvent(sc, t->ev_dfeat, 0));
    }
    // Output layer: dW_out = dz_t^T . feat_d per dense column block (feature columns [0, 256) and [256, 256 + hb[1])).
    // Default: the block's dW tile stays in tensor memory and the dense TF1 Adam is applied from there (k_dw_adam_fused):
    // dW_out never exists in HBM.  Debug bit 14: dW_out through HBM (buffer "g_W_out", same block layout) +
    // k_adam_rows_vec4 + a cast of the blocks into the operand copy.
    const bool two_kernel = (m->debug & 16384) != 0;
    ph_begin(m, PH_T_DW_ADAM, t->st);
    for (int half = 0; half < 2; ++half) {
        const int hw = t->hb[half];
        if (hw == 0) continue;
        DwArgs w{};
        w.dzT = t->dzT; w.h_dT = t->feat_dT + (size_t)half * 256 * bpad; w.n_local = N; w.N = N; w.H = hw;
        w.K = bpad; w.pt.world = 1;
        if (two_kernel) w.g = t->g_W_out + t->blk_off[half];
        else {
            w.w = t->W_out + t->blk_off[half]; w.m = t->m_W_out + t->blk_off[half]; w.v = t->v_W_out + t->blk_off[half];
            w.shadow = t->W_out_bf16; w.shadow_ld = kTitleFpad; w.shadow_col0 = 256 * half;
            w.adam = AdamConst{ad.alpha, ad.one_minus_b1, ad.one_minus_b2, ad.eps, ad.lambda};
        }
        launch_dw(w, t->st);
        t->launches += 1;
        if (two_kernel) {
            ad.w = t->W_out + t->blk_off[half]; ad.m = t->m_W_out + t->blk_off[half]; ad.v = t->v_W_out + t->blk_off[half];
            ad.g = t->g_W_out + t->blk_off[half]; ad.n = (long long)N * hw; ad.row_len = hw;
            launch_adam_rows(ad, nullptr, nullptr, t->st);
            t->launches += 1;
        }
    }
    if (two_kernel) refresh_out_shadow(t);
    ph_end(m, PH_T_DW_ADAM, t->st);
    ph_begin(m, PH_T_CNN_BWD, sc);
    CnnBwdArgs b{};
    b.titles = t->titles; b.emb = t->emb; b.conv_W = t->conv_W; b.shape = t->shape; b.dh_partial = t->dh_partial;
    b.nsplit = t->nsplit; b.bpad = bpad; b.B = batch; b.feat = t->feat; b.argpos = t->argpos; b.kp_t = title_keep_prob;
    b.seed = m->cfg.seed; b.step = (unsigned long long)t->step; b.row_offset = 0; b.d = t->d; b.dx = t->dx; b.conv_WT = t->conv_WT;
    b.g_emb = t->g_emb; b.g_conv_W = t->g_conv_W; b.g_conv_b = t->g_conv_b;
    launch_charcnn_bwd(b, sc);
    ph_end(m, PH_T_CNN_BWD, sc);
    t->launches += 4;

    ph_begin(m, PH_T_ADAM_SMALL, sc);
    ad.row_len = 1; ad.g = nullptr;
    const CnnShape& s = t->shape;
    ad.w = t->emb; ad.m = t->m_emb; ad.v = t->v_emb; ad.g = t->g_emb; ad.n = (long long)s.C * s.E; launch_adam(ad, sc);
    ad.w = t->conv_W; ad.m = t->m_conv_W; ad.v = t->v_conv_W; ad.g = t->g_conv_W; ad.n = t->n_conv_w; launch_adam(ad, sc);
    ad.w = t->conv_b; ad.m = t->m_conv_b; ad.v = t->v_conv_b; ad.g = t->g_conv_b; ad.n = t->D; launch_adam(ad, sc);
    launch_conv_transpose(t->conv_W, t->conv_WT, t->shape, sc);
    ph_end(m, PH_T_ADAM_SMALL, sc);
    if (sc != t->st) CK(cudaEventRecord(t->ev_cnn, sc));
    ad.w = t->b_out; ad.m = t->m_b_out; ad.v = t->v_b_out; ad.g = t->g_b_out; ad.n = N; launch_adam(ad, t->st);
    if (sc != t->st) CK(cudaStreamWaitEvent(t->st, t->ev_cnn, 0));     // join: the next step's CNN forward reads the updated variables
    t->launches += 5;
    t->b1_pow *= kBeta1; t->b2_pow *= kBeta2; t->step += 1;

    if (!sync) {
        CK(cudaMemcpyAsync(t->cost_ring + slot, t->cost, sizeof(float), cudaMemcpyDeviceToHost, t->st));
        CK(cudaMemcpyAsync(t->err_ring + slot, m->err, sizeof(int), cudaMemcpyDeviceToHost, t->st));
        CK(cudaEventRecord(t->ev_cost[slot], t->st));
        return 0;
    }
    CK(cudaMemcpyAsync(t->cost_host, t->cost, sizeof(float), cudaMemcpyDeviceToHost, t->st));
    TRY(check_device_flag(m));
    ph_collect(m);
    if (cost_out) *cost_out = *t->cost_host;
    return 0;
}

extern "C" int32_t dae_title_train_step(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                        const int64_t* y_pos, const float* y_val, int64_t nnz_y, const int64_t* titles,
                                        const float* titles_use, int32_t batch, float keep_prob, float input_keep_prob,
                                        float title_keep_prob, float* cost_out) {
    if (!t || !t->trainable) return fail("title model is not trainable");
    if (t->async_pending) TRY(dae_title_train_flush(t, nullptr, nullptr));       // drain a pipelined step first
    return title_train_impl(t, x_pos, x_val, nnz_x, y_pos, y_val, nnz_y, titles, titles_use, batch, keep_prob, input_keep_prob,
                            title_keep_prob, 0, true, cost_out);
}

static int title_collect(dae_title* t, int slot, float* cost_out) {
    CK(cudaEventSynchronize(t->ev_cost[slot]));
    const int e = t->err_ring[slot];
    if (e != 0) {
        cudaMemsetAsync(t->dae->err, 0, sizeof(int), t->st);
        return fail("invalid sparse batch (flag %d)", e);
    }
    if (cost_out) *cost_out = t->cost_ring[slot];
    return 0;
}

// The same step pipelined for the training loop (main_train.py:214-223 only accumulates the cost): the batch is staged into
// the slot the device is not using while the previous step still runs, the step is enqueued, and the PREVIOUS step's cost
// comes back (*has_prev = 0 on the first call).  dae_title_train_flush returns the last pending cost.
extern "C" int32_t dae_title_train_step_async(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                              const int64_t* y_pos, const float* y_val, int64_t nnz_y, const int64_t* titles,
                                              const float* titles_use, int32_t batch, float keep_prob, float input_keep_prob,
                                              float title_keep_prob, float* prev_cost_out, int32_t* has_prev) {
    if (!t || !t->trainable) return fail("title model is not trainable");
    const int slot = t->async_slot ^ 1;
    TRY(title_train_impl(t, x_pos, x_val, nnz_x, y_pos, y_val, nnz_y, titles, titles_use, batch, keep_prob, input_keep_prob,
                         title_keep_prob, slot, false, nullptr));
    const bool had = t->async_pending;
    t->async_slot = slot;
    t->async_pending = true;
    if (has_prev) *has_prev = had ? 1 : 0;
    if (had) return title_collect(t, slot ^ 1, prev_cost_out);
    return 0;
}

extern "C" int32_t dae_title_train_flush(dae_title* t, float* cost_out, int32_t* has_cost) {
    if (!t) return fail("null model");
    if (has_cost) *has_cost = t->async_pending ? 1 : 0;
    if (!t->async_pending) return 0;
    t->async_pending = false;
    return title_collect(t, t->async_slot, cost_out);
}

static int title_scores(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x, const int64_t* titles,
                        const float* titles_use, int32_t batch, int32_t n_cols) {
    dae_model* m = t->dae;
    TRY(title_forward(t, x_pos, x_val, nnz_x, nullptr, nullptr, 0, titles, titles_use, batch, 1.f, 1.f, 1.f, false));
    CK(cudaEventRecord(m->slots[0].consumed, t->st));
    const size_t need = (size_t)batch * n_cols;
    if (m->scores_elems < need) {
        if (m->scores) { CK(cudaStreamSynchronize(t->st)); CK(cudaFree(m->scores)); m->scores = nullptr; }
        CK(cudaMalloc(reinterpret_cast<void**>(&m->scores), need * sizeof(float)));
        m->scores_elems = need;
    }
    TitleTileArgs a = tile_args(t, batch);
    a.out = m->scores; a.ld_out = n_cols; a.n_out = n_cols;
    launch_title_predict(a, t->st);
    t->launches += 1;
    return 0;
}

// sess.run(model.y_pred, {..., titles, titles_use, all keep probabilities 1})   main_train.py:69-79, main_challenge.py:80-85
extern "C" int32_t dae_title_predict(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                     const int64_t* titles, const float* titles_use, int32_t batch, int32_t n_cols,
                                     float* y_pred_out) {
    if (!t || !y_pred_out) return fail("null argument");
    if (n_cols <= 0 || n_cols > t->N) return fail("n_cols must be in (0, n_input]");
    TRY(title_scores(t, x_pos, x_val, nnz_x, titles, titles_use, batch, n_cols));
    CK(cudaMemcpyAsync(y_pred_out, t->dae->scores, (size_t)batch * n_cols * sizeof(float), cudaMemcpyDeviceToHost, t->st));
    return check_device_flag(t->dae);
}

// y_pred[:, :n_tracks] + cand_generate (argsort, seed removal, first k)           main_challenge.py:26-36, :87-90
// (+ the metrics of the lists against the answers CSR when metrics_out != NULL: main_train.py:88-100 on the device)
static int title_rank(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x, const int64_t* titles,
                      const float* titles_use, int32_t batch, const int32_t* seed_ptr, const int32_t* seed_idx, int32_t k,
                      int32_t* out_idx, float* out_score, const int32_t* ans_ptr, const int32_t* ans_idx, double* metrics_out) {
    if (k <= 0 || k > 1024) return fail("k must be in [1,1024]");
    dae_model* m = t->dae;
    const int T = m->T;
    TRY(title_scores(t, x_pos, x_val, nnz_x, titles, titles_use, batch, T));
    int *d_idx = nullptr, *d_sp = nullptr, *d_si = nullptr;
    float* d_sc = nullptr;
    const int nseed = seed_ptr ? seed_ptr[batch] : 0;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_idx), sizeof(int) * batch * k, t->st));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_sc), sizeof(float) * batch * k, t->st));
    if (seed_ptr) {
        CK(cudaMallocAsync(reinterpret_cast<void**>(&d_sp), sizeof(int) * (batch + 1), t->st));
        CK(cudaMallocAsync(reinterpret_cast<void**>(&d_si), sizeof(int) * (nseed > 0 ? nseed : 1), t->st));
        CK(cudaMemcpyAsync(d_sp, seed_ptr, sizeof(int) * (batch + 1), cudaMemcpyHostToDevice, t->st));
        if (nseed > 0) CK(cudaMemcpyAsync(d_si, seed_idx, sizeof(int) * nseed, cudaMemcpyHostToDevice, t->st));
    }
    TopkArgs a{};
    a.scores = m->scores; a.ld = T; a.B = batch; a.T = T; a.k = k; a.seed_ptr = d_sp; a.seed_idx = d_si; a.idx_base = 0;
    a.out_idx = d_idx; a.out_score = d_sc;
    launch_topk(a, t->st);
    t->launches += 1;
    int rc = 0;
    if (metrics_out) { rc = run_metrics(m, d_idx, batch, k, ans_ptr, ans_idx, metrics_out); t->launches += 1; }
    if (out_idx) CK(cudaMemcpyAsync(out_idx, d_idx, sizeof(int) * batch * k, cudaMemcpyDeviceToHost, t->st));
    if (out_score) CK(cudaMemcpyAsync(out_score, d_sc, sizeof(float) * batch * k, cudaMemcpyDeviceToHost, t->st));
    cudaFreeAsync(d_idx, t->st); cudaFreeAsync(d_sc, t->st);
    if (d_sp) cudaFreeAsync(d_sp, t->st);
    if (d_si) cudaFreeAsync(d_si, t->st);
    if (rc) return rc;
    return check_device_flag(m);
}

extern "C" int32_t dae_title_recommend(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                       const int64_t* titles, const float* titles_use, int32_t batch,
                                       const int32_t* seed_ptr, const int32_t* seed_idx, int32_t k, int32_t* out_idx,
                                       float* out_score) {
    if (!t || !out_idx) return fail("null argument");
    return title_rank(t, x_pos, x_val, nnz_x, titles, titles_use, batch, seed_ptr, seed_idx, k, out_idx, out_score, nullptr,
                      nullptr, nullptr);
}

// dae_title_recommend + met.single_eval of every playlist on the device -> [batch, 3] (r-precision, ndcg, clicks)
extern "C" int32_t dae_title_evaluate(dae_title* t, const int64_t* x_pos, const float* x_val, int64_t nnz_x,
                                      const int64_t* titles, const float* titles_use, int32_t batch,
                                      const int32_t* seed_ptr, const int32_t* seed_idx, const int32_t* ans_ptr,
                                      const int32_t* ans_idx, int32_t k, double* metrics_out) {
    if (!t || !metrics_out) return fail("null argument");
    return title_rank(t, x_pos, x_val, nnz_x, titles, titles_use, batch, seed_ptr, seed_idx, k, nullptr, nullptr, ans_ptr,
                      ans_idx, metrics_out);
}

extern "C" int64_t dae_title_launch_count(dae_title* t) { return t ? t->launches : 0; }

extern "C" int32_t dae_title_buffer(dae_title* t, const char* name, void** dev_ptr, int64_t* n_elem, int32_t* elem_size) {
    if (!t || !name || !dev_ptr) return fail("null argument");
    const CnnShape& s = t->shape;
    const int64_t NF = (int64_t)t->N * kTitleFpad, NB = (int64_t)round_up(t->N, kTileItems) * (t->hb[0] + t->hb[1]);
    struct E { const char* n; void* p; int64_t c; int32_t s; };
    const E table[] = {
        {"feat", t->feat, (int64_t)t->Bmax * t->D, 4}, {"argpos", t->argpos, (int64_t)t->Bmax * t->D, 1},
        {"feat_d", t->feat_d, (int64_t)kMaxBpad * kTitleFpad, 2}, {"w_t", t->w_t, kMaxBpad, 4}, {"w_p", t->w_p, kMaxBpad, 4},
        {"dzT", t->dzT, (int64_t)t->N * kMaxBpad, 2}, {"d", t->d, (int64_t)t->Bmax * t->D, 4},
        {"g_emb", t->g_emb, (int64_t)s.C * s.E, 4}, {"g_conv_W", t->g_conv_W, t->n_conv_w, 4},
        {"g_conv_b", t->g_conv_b, t->D, 4}, {"g_W_out", t->g_W_out, NB, 4}, {"g_b_out", t->g_b_out, t->N, 4},
        {"W_out", t->W_out, NB, 4}, {"W_out_bf16", t->W_out_bf16, NF, 2}, {"m_W_out", t->m_W_out, NB, 4}, {"v_W_out", t->v_W_out, NB, 4},
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
