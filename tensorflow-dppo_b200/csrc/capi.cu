// C ABI of libdppo_b200.so (include/dppo_b200.h): context, workspace carving and the
// stream-ordered orchestration of the update path
//     GAE (Worker.py:82-92) -> pi & oldpi forward (Model.py:11-15, PPO.py:21-22)
//     -> loss + d_out (PPO.py:29-40,46) -> dense backward -> split-K finalize
//     -> allreduce (PPO.py:55-65) -> Adam (PPO.py:53).
#include "common.cuh"
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <stddef.h>
#include <new>

namespace dppo {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- launch accounting: one ProfState per context.  Every C-ABI entry point binds the calling thread to its
// context's state (CTX_GUARD); the kernel launchers, which only see a stream, account through that binding. ----
struct ProfState {
    bool timing = false;
    long long launches[KK_COUNT] = {0};
    double ms[KK_COUNT] = {0};
    static constexpr int kMaxSlots = 8192;
    cudaEvent_t ev0[kMaxSlots], ev1[kMaxSlots];
    int kind_of[kMaxSlots];
    int created = 0, used = 0;
};
static thread_local ProfState* t_prof = nullptr;

ProfScope::ProfScope(int k, cudaStream_t s) : kind(k), st(s), slot(-1) {
    ProfState* p = t_prof;
    if (!p) return;
    p->launches[k]++;
    if (!p->timing || p->used >= ProfState::kMaxSlots) return;
    if (p->used >= p->created) {
        if (cudaEventCreate(&p->ev0[p->created]) != cudaSuccess) return;
        if (cudaEventCreate(&p->ev1[p->created]) != cudaSuccess) return;
        p->created++;
    }
    slot = p->used++;
    p->kind_of[slot] = k;
    cudaEventRecord(p->ev0[slot], st);
}
ProfScope::~ProfScope() {
    if (slot >= 0 && t_prof) cudaEventRecord(t_prof->ev1[slot], st);
}

bool peer_available(const dppo_ctx* ctx);
float* peer_exchange_buffer(dppo_ctx* ctx);
int peer_allreduce_adam(dppo_ctx* ctx, float* params, float* m, float* v, int step, float lr, float b1, float b2, float eps,
                        float* grad_out, cudaStream_t st, const StepParams* sp = nullptr, const int* counter = nullptr);
float* peer_exchange_half(dppo_ctx* ctx, int sel);
int comm_unique_id(void* out);
int comm_init(dppo_ctx* ctx, int rank, int world, const void* id_bytes);
int comm_destroy(dppo_ctx* ctx);
int comm_peer_export(dppo_ctx* ctx, int rank, int world, void* handle_out);
int comm_peer_import(dppo_ctx* ctx, const void* all_handles);
int allreduce_avg(dppo_ctx* ctx, float* buf, int64_t n, cudaStream_t st);

int make_shape(const dppo_config* cfg, NetShape* s) {
    DPPO_REQUIRE(cfg != nullptr, "null config");
    DPPO_REQUIRE(cfg->obs_dim >= 1 && cfg->act_dim >= 1, "obs_dim=%d act_dim=%d must be >= 1", cfg->obs_dim, cfg->act_dim);
    DPPO_REQUIRE(cfg->n_hidden >= 0 && cfg->n_hidden <= kMaxLayers, "n_hidden=%d out of [0,%d]", cfg->n_hidden, kMaxLayers);
    DPPO_REQUIRE(cfg->pd_type == DPPO_PD_DIAG_GAUSSIAN || cfg->pd_type == DPPO_PD_CATEGORICAL,
                 "pd_type=%d unknown (DPPO_PD_DIAG_GAUSSIAN=0, DPPO_PD_CATEGORICAL=1)", cfg->pd_type);
    s->obs = cfg->obs_dim; s->act = cfg->act_dim; s->L = cfg->n_hidden;
    s->pd = cfg->pd_type;
    s->pdw = (s->pd == DPPO_PD_CATEGORICAL) ? cfg->act_dim : 2 * cfg->act_dim;      // pdtype.param_shape()[0], Model.py:14
    s->acols = (s->pd == DPPO_PD_CATEGORICAL) ? 1 : cfg->act_dim;
    s->P = 1 + s->pdw;
    s->Pld = (s->P + 3) & ~3;
    int64_t off = 0;
    int fan_in = cfg->obs_dim;
    for (int l = 0; l < s->L; ++l) {
        DPPO_REQUIRE(cfg->hidden[l] >= 1, "hidden[%d]=%d must be >= 1", l, cfg->hidden[l]);
        s->h[l] = cfg->hidden[l];
        s->w_off[l] = off; off += (int64_t)fan_in * s->h[l];
        s->b_off[l] = off; off += s->h[l];
        fan_in = s->h[l];
    }
    s->w_off[s->L] = off; off += (int64_t)fan_in * s->P;
    s->b_off[s->L] = off; off += s->P;
    s->n_params = off;
    // split-K plan: as many slices as keep (output tiles x slices) within one wave of 2 CTAs/SM
    int64_t poff = 0;
    for (int l = 0; l <= s->L; ++l) {
        const int in = s->in_dim(l), out = s->out_dim(l);
        const int bn = out > 64 ? 128 : 64;
        const int tiles = ((in + 127) / 128) * ((out + bn - 1) / bn);
        int sl = (2 * kPlanSMs) / tiles;
        if (tc_mode(cfg->gemm_mode)) {
            // tensor-core accumulators truncate on every add: bound the rows one CTA reduces (<= 512, i.e. 64
            // MMA steps); the cross-slice sum in finalize_kernel is round-to-nearest fp32
            // (+25 % head-room: active_slices() may round the count up to whole waves of the dW grid)
            sl = (cfg->max_rows + 511) / 512;
            sl = sl + sl / 4 + 1;
            if (sl > 640) sl = 640;
        } else {
            if (sl < 1) sl = 1;
            if (sl > kMaxSplitK) sl = kMaxSplitK;
            if (cfg->gemm_mode == DPPO_GEMM_BF16 && cfg->max_rows >= 65536) {
                // the persistent dW kernel (one-term mode of tc_dw_h.cu) walks tiles x slices work items on 148 CTAs: pick the
                // smallest slice count (<= 40, >= 2048 rows each) whose last wave is within 3 % of the fullest (config 4: 32 tiles x 9
                // slices = 288 items = 1.95 waves)
                const int t2 = ((in + 127) / 128) * ((out + 255) / 256);
                int best = sl; double best_w = 1e30;
                for (int c2 = 4; c2 <= 40 && (long long)c2 * 2048 <= cfg->max_rows; ++c2) {
                    const long long items = (long long)t2 * c2;
                    const double waste = (double)((items + kPlanSMs - 1) / kPlanSMs * kPlanSMs) / (double)items;
                    if (waste < best_w - 0.03) { best_w = waste; best = c2; }      // (fewer slices unless > 3 % fuller: finalize streams them all)
                }
                sl = best;
            }
        }
        s->slices[l] = sl;
        s->region[l] = (((int64_t)in * out + out) + 3) & ~3LL;
        s->part_off[l] = poff;
        poff += (int64_t)sl * s->region[l];
    }
    s->partial_floats = poff;
    return 0;
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// one pass computes the size (base == nullptr) or carves the pointers
static size_t carve(const dppo_config* cfg, const NetShape& s, char* base, Workspace* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* {
        char* p = base ? base + off : nullptr;
        off += align_up(bytes);
        return p;
    };
    const size_t R = (size_t)cfg->max_rows;
    const int H = s.in_dim(s.L);
    const size_t head = ((size_t)H * s.P + s.P) * sizeof(float);
    Workspace w{};
    w.head_pi = (float*)take(head);
    w.head_old = (float*)take(head);
    if (cfg->gemm_mode == DPPO_GEMM_BF16) {
        w.wt_pi = (float*)take((size_t)s.n_params * sizeof(float));
        w.wt_old = (float*)take((size_t)s.n_params * sizeof(float));
    }
    if (cfg->gemm_mode == DPPO_GEMM_TF32X3) {
        for (int l = 0; l <= s.L; ++l) {
            const size_t fwd = (size_t)tc_image_floats(s.out_dim(l), s.in_dim(l)) * sizeof(float);
            w.img_fwd_pi[l] = (float*)take(fwd);
            w.img_fwd_old[l] = (float*)take(fwd);
            w.img_dx_pi[l] = l > 0 ? (float*)take((size_t)tc_image_floats(s.in_dim(l), s.out_dim(l)) * sizeof(float)) : nullptr;
        }
    }
    for (int l = 0; l < s.L; ++l) w.act_pi[l] = (float*)take(R * s.h[l] * sizeof(float));
    if (cfg->gemm_mode == DPPO_GEMM_TF32X3)
        for (int l = 0; l < s.L; ++l) w.act_bits[l] = (uint32_t*)take((size_t)tc_bits_words((long long)R, s.h[l]) * sizeof(uint32_t));
    if (cfg->gemm_mode == DPPO_GEMM_FP16X2 || cfg->gemm_mode == DPPO_GEMM_BF16) {     // (bf16: the one-term mode of the same kernels)
        for (int l = 0; l <= s.L; ++l) {
            const size_t fwd = (size_t)th_image_floats(s.out_dim(l), s.in_dim(l)) * sizeof(float);
            w.img_fwd_pi[l] = (float*)take(fwd);
            w.img_fwd_old[l] = (float*)take(fwd);
            w.img_dx_pi[l] = l > 0 ? (float*)take((size_t)th_image_floats(s.in_dim(l), s.out_dim(l)) * sizeof(float)) : nullptr;
        }
        for (int l = 0; l < s.L; ++l) w.act_bits[l] = (uint32_t*)take((size_t)th_bits_words((long long)R, s.h[l]) * sizeof(uint32_t));
    }
    const size_t hm = (size_t)(s.hmax() > 0 ? s.hmax() : 1);
    w.scratch[0] = (float*)take(R * hm * sizeof(float));
    w.scratch[1] = (float*)take(R * hm * sizeof(float));
    w.out_pi = (float*)take(R * s.Pld * sizeof(float));
    w.out_old = (float*)take(R * s.Pld * sizeof(float));
    w.d_out = (float*)take(R * s.Pld * sizeof(float));
    w.partial = (float*)take((size_t)s.partial_floats * sizeof(float));
    const size_t tiles = (R + kLossTileRows - 1) / kLossTileRows + 64;
    w.loss_part = (double*)take(tiles * kLossSlots * sizeof(double));
    w.gae_adv = (float*)take(R * sizeof(float));
    w.gae_ret = (float*)take(R * sizeof(float));
    if (cfg->gemm_mode == DPPO_GEMM_BF16 && s.obs % 8 == 0)
        // one-term mode: the observations are rounded to bf16 ONCE per row block (they feed pi's and oldpi's first layer and
        // layer 1's dW) -- as a plane they take the TMA-fed, pair, 64-element k-block path like every hidden activation
        w.obs_pl = (float*)take(R * (size_t)s.obs * 2);
    if (cfg->gemm_mode == DPPO_GEMM_FP16X2 && s.obs % 4 != 0) {
        // observation rows whose pitch TMA cannot address (obs_dim 3, 17, ...): the dense layers read a re-pitched copy
        w.obs_ld = (s.obs + 3) & ~3;
        w.obs_pad = (float*)take(R * (size_t)w.obs_ld * sizeof(float));
    }
    w.stage_floats = R * ((size_t)s.obs + s.acols + 3);
    w.stage = (float*)take(w.stage_floats * sizeof(float));
    w.step_params = (StepParams*)take((size_t)kMaxGraphSteps * sizeof(StepParams));
    w.step_counter = (int*)take(256);
    if (ws) *ws = w;
    return off;
}

// fp16x2 mode: hidden activations and their gradients are stored as fp16 hi | lo planes (same bytes as fp32) whenever every
// hidden width keeps the plane rows 16-byte aligned (DPPO_H_PLANES=0 at dppo_create keeps fp32 matrices + converter warps)
static bool h_planes(const dppo_ctx* c) {
    const bool bf16 = c->cfg.gemm_mode == DPPO_GEMM_BF16;
    if ((c->cfg.gemm_mode != DPPO_GEMM_FP16X2 && !bf16) || !c->use_h_planes || !c->use_h_dw) return false;
    for (int l = 0; l < c->shape.L; ++l)
        if (c->shape.h[l] % 8 != 0) return false;
    // DPPO_GEMM_BF16 runs on these kernels in their one-term mode (one bf16 plane per activation, one MMA per product) only
    // when every operand can be read by TMA; other shapes keep the round-1 bf16 kernels (tc_gemm_bf16.cu)
    if (bf16 && c->shape.obs % 4 != 0) return false;
    return c->shape.L > 0;
}
// the persistent kernel family of tc_gemm_h.cu / tc_dw_h.cu: 2 = fp16 hi + lo operands (DPPO_GEMM_FP16X2), 1 = bf16 operands
// (DPPO_GEMM_BF16 on TMA-addressable shapes), 0 = not on this family
static int th_terms(const dppo_ctx* c) {
    if (c->cfg.gemm_mode == DPPO_GEMM_FP16X2) return 2;
    if (c->cfg.gemm_mode == DPPO_GEMM_BF16 && h_planes(c)) return 1;
    return 0;
}

// ---- forward of pi (and oldpi) over rows [0, rows) of obs ------------------------------------
static int forward_nets(dppo_ctx* c, const float* params_pi, const float* params_old, const float* obs,
                        int rows, cudaStream_t st) {
    const NetShape& s = c->shape;
    Workspace& w = c->ws;
    const int nets = params_old ? 2 : 1;
    if (int r = launch_pack_heads(s, params_pi, params_old, w.head_pi, w.head_old, st)) return r;
    const float* in0 = obs; const float* in1 = obs;
    int ld0 = s.obs;                                     // pitch of the layer-0 operand
    if (w.obs_pad && th_terms(c) && c->use_h_dw) {
        // obs_dim % 4 != 0: re-pitch the rows once so that TMA feeds layer 1's forward and dW like any other shape (the pad
        // columns are never read: the tensor maps end at obs_dim and zero-fill beyond it)
        DPPO_CHECK_CUDA(cudaMemcpy2DAsync(w.obs_pad, (size_t)w.obs_ld * 4, obs, (size_t)s.obs * 4, (size_t)s.obs * 4, (size_t)rows,
                                          cudaMemcpyDeviceToDevice, st));
        in0 = in1 = w.obs_pad; ld0 = w.obs_ld;
    }
    if (c->cfg.gemm_mode == DPPO_GEMM_BF16 && th_terms(c) == 0) {
        // bf16 tensor cores: B operand K-major = transposed kernels (fp32 in HBM, rounded to bf16 by the movers)
        if (int r = tc_transpose_weights(s, params_pi, w.head_pi, w.wt_pi, params_old, w.head_old, w.wt_old, st)) return r;
        for (int l = 0; l <= s.L; ++l) {
            const bool head = (l == s.L);
            float* o0 = head ? w.out_pi : w.act_pi[l];
            float* o1 = head ? w.out_old : w.scratch[l & 1];
            const int in = s.in_dim(l), out = s.out_dim(l);
            const float* b0 = head ? w.head_pi + (size_t)in * out : params_pi + s.b_off[l];
            const float* b1 = head ? w.head_old + (size_t)in * out : (params_old ? params_old + s.b_off[l] : nullptr);
            if (int r = bf16_gemm_kk(0, rows, out, in, in0, in1, in, w.wt_pi + s.w_off[l], w.wt_old + s.w_off[l], in, o0, o1,
                                     head ? s.Pld : out, b0, b1, head ? 0 : 1, nullptr, nets, st)) return r;
            in0 = o0; in1 = o1;
        }
        return 0;
    }
    if (th_terms(c)) {
        // persistent fp16-split kernels: weight images (fp16 hi | lo tiles; bf16 tiles in one-term mode) refreshed from the
        // current parameters
        const int terms = th_terms(c);
        if (int r = th_pack_images(s, params_pi, w.head_pi, params_old, w.head_old, w.img_fwd_pi, w.img_fwd_old,
                                   w.img_dx_pi, st, terms, rows, h_planes(c) ? (terms == 1 && w.obs_pl ? 2 : 1) : 0)) return r;
        // hidden activations live as fp16 hi | lo planes (h_planes): the epilogue that produces them applies the split the
        // consumer's converter warps would apply, so every later product reads its operand tiles straight from TMA
        const bool planes = h_planes(c);
        const long long R = c->cfg.max_rows;
        const bool obs_plane = planes && terms == 1 && w.obs_pl != nullptr;
        if (obs_plane) {
            if (int r = th_split_planes(obs, (long long)rows * s.obs, w.obs_pl, nullptr, st, 1)) return r;
            in0 = in1 = w.obs_pl;
        }
        for (int l = 0; l <= s.L; ++l) {
            const bool head = (l == s.L);
            float* o0 = head ? w.out_pi : w.act_pi[l];
            float* o1 = head ? w.out_old : w.scratch[l & 1];
            const int in = s.in_dim(l), out = s.out_dim(l);
            const float* b0 = head ? w.head_pi + (size_t)in * out : params_pi + s.b_off[l];
            const float* b1 = head ? w.head_old + (size_t)in * out : (params_old ? params_old + s.b_off[l] : nullptr);
            ThPlanes pl{};
            pl.terms = terms;
            pl.a_planes = planes && (l > 0 || obs_plane); pl.a_lo_off = R * in;
            pl.c_planes = planes && !head; pl.c_lo_off = R * out;
            if (int r = th_gemm(0, rows, out, in, in0, in1, l == 0 ? ld0 : in, w.img_fwd_pi[l], w.img_fwd_old[l], o0, o1,
                                head ? s.Pld : out, b0, b1, head ? 0 : 1, nullptr, nets, st,
                                head ? nullptr : w.act_bits[l], nullptr, nullptr, &pl)) return r;
            in0 = o0; in1 = o1;
        }
        return 0;
    }
    if (c->cfg.gemm_mode == DPPO_GEMM_TF32X3) {
        // tensor cores: B operand K-major = transposed kernels, refreshed from the current parameters
        if (int r = tc_pack_images(s, params_pi, w.head_pi, params_old, w.head_old, w.img_fwd_pi, w.img_fwd_old,
                                   w.img_dx_pi, st)) return r;
        for (int l = 0; l <= s.L; ++l) {
            const bool head = (l == s.L);
            float* o0 = head ? w.out_pi : w.act_pi[l];
            float* o1 = head ? w.out_old : w.scratch[l & 1];
            const int in = s.in_dim(l), out = s.out_dim(l);
            const float* b0 = head ? w.head_pi + (size_t)in * out : params_pi + s.b_off[l];
            const float* b1 = head ? w.head_old + (size_t)in * out : (params_old ? params_old + s.b_off[l] : nullptr);
            if (int r = tc_gemm_kk_img(0, rows, out, in, in0, in1, in, w.img_fwd_pi[l], w.img_fwd_old[l], o0, o1,
                                       head ? s.Pld : out, b0, b1, head ? 0 : 1, nullptr, nets, st,
                                       head ? nullptr : w.act_bits[l], nullptr)) return r;
            in0 = o0; in1 = o1;
        }
        return 0;
    }
    for (int l = 0; l < s.L; ++l) {
        float* o0 = w.act_pi[l];
        float* o1 = w.scratch[l & 1];
        if (int r = gemm_nn_dual(rows, s.h[l], s.in_dim(l), in0, in1, s.in_dim(l),
                                 params_pi + s.w_off[l], params_old ? params_old + s.w_off[l] : nullptr, s.h[l],
                                 o0, o1, s.h[l], params_pi + s.b_off[l],
                                 params_old ? params_old + s.b_off[l] : nullptr, 1, nets, st)) return r;
        in0 = o0; in1 = o1;
    }
    const int H = s.in_dim(s.L);
    return gemm_nn_dual(rows, s.P, H, in0, in1, H, w.head_pi, w.head_old, s.P, w.out_pi, w.out_old, s.Pld,
                        w.head_pi + (size_t)H * s.P, w.head_old + (size_t)H * s.P, 0, nets, st);
}

// Split-K slices one update really needs: the plan (NetShape::slices) is sized for max_rows; a call over fewer rows
// (minibatches, the row chunks of dppo_update_host) uses proportionally fewer, so the dW kernels launch fewer CTAs and
// finalize_kernel streams fewer partials.  `rows` = the largest row block of the update (all its blocks share the count).
static void active_slices(const dppo_ctx* c, int rows, int* active) {
    const NetShape& s = c->shape;
    for (int l = 0; l <= s.L; ++l) {
        int a = s.slices[l];
        if (tc_mode(c->cfg.gemm_mode)) {
            // <= 512 rows per TMEM accumulation chain (64 truncating adds of a K = 8 tf32 MMA); the fp16 dW kernel covers
            // K = 16 per MMA, so 1024 rows are the same 64 adds -- half the partial slices to write and to sum
            const bool h_dw = c->cfg.gemm_mode == DPPO_GEMM_FP16X2 && c->use_h_dw && (s.in_dim(l) % 4 == 0 || (l == 0 && c->ws.obs_pad)) &&
                              (l == s.L ? s.Pld : s.out_dim(l)) % 4 == 0;
            const int chain = h_dw ? 1024 : 512;
            int want = (rows + chain - 1) / chain;
            // whole waves: the dW grid is tiles x slices CTAs on 2 x 148 slots (two CTAs per SM).  Round the slice count
            // up to the next count that fills its last wave when that costs <= 25 % more slices (cfg3: 256 -> 296, the
            // 6-tile layer-1 grid goes from 5.2 to 6.0 waves of 12 % shorter CTAs), then drop slices left without rows.
            // (the persistent fp16 kernel: 148 CTAs walk tiles x slices work items, 128 x 256 / 128 / 64 tiles)
            const int in = s.in_dim(l), out = s.out_dim(l);
            const int bn = h_dw ? (out > 128 ? 256 : (out > 64 ? 128 : 64)) : (out > 64 ? 128 : 64);
            const int tiles = ((in + 127) / 128) * ((out + bn - 1) / bn);
            const int slots = h_dw ? kPlanSMs : 2 * kPlanSMs;
            int g = slots, t = tiles;
            while (t) { const int r = g % t; g = t; t = r; }            // gcd(slots, tiles)
            const int W = slots / g;
            // short row blocks (minibatches, the H2D chunks of the host-buffer path): at least one work item per CTA slot,
            // as long as an item keeps >= 128 rows
            const int fill = (slots + tiles - 1) / tiles;
            if (want < fill) { const int cap = rows / 128 > 1 ? rows / 128 : 1; want = fill < cap ? fill : cap; }
            const int up = ((want + W - 1) / W) * W;
            if (up * 4 <= want * 5) want = up;
            const int kps = (((rows + want - 1) / want + 31) / 32) * 32;     // whole k-blocks of the fp16 (32-row) and tf32 (16-row) kernels
            want = (rows + kps - 1) / kps;
            if (want < a) a = want;
        } else {
            const int want = (rows + 63) / 64;              // no point slicing below 64 rows per CTA
            if (want < a) a = want;
        }
        active[l] = a < 1 ? 1 : a;
    }
}

// fp16-split mode: the loss kernel emits d_out multiplied by this power of two (the backward operands must sit in fp16's
// range; raw they are ~1/B), every backward product carries the factor and finalize_kernel divides it out (exact in fp32)
static float grad_prescale(const dppo_ctx* c, int B_total) {
    if (c->cfg.gemm_mode != DPPO_GEMM_FP16X2) return 1.0f;
    return exp2f(floorf(log2f((float)(B_total > 1 ? B_total : 1))));
}

// ---- fwd + loss + bwd over one block of rows; dW/db land in the split-K partial buffer --------
static int loss_grad_rows(dppo_ctx* c, const float* params_pi, const float* params_old, const float* obs,
                          const float* act, const float* adv, const float* ret, int rows, int tile0,
                          int B_total, float l_mul, int accumulate, const int* active, cudaStream_t st,
                          cudaEvent_t adv_ready = nullptr, const float* old_rows = nullptr,
                          const float* ret_all = nullptr, int row_off = 0, int groups = 1) {
    const NetShape& s = c->shape;
    Workspace& w = c->ws;
    // old_rows: the oldpi head outputs [rows, Pld] of exactly these rows, computed earlier (oldpi is frozen while the
    // chief trains, so epochs after the first need not re-evaluate it); nullptr -> evaluate both nets here
    if (int r = forward_nets(c, params_pi, old_rows ? nullptr : params_old, obs, rows, st)) return r;
    if (adv_ready) DPPO_CHECK_CUDA(cudaStreamWaitEvent(st, adv_ready, 0));   // adv/ret come from a side stream
    const float inv_B = grad_prescale(c, B_total) / (float)B_total;     // 1/B of the batch means (x the fp16-mode gradient scale)
    const float clip = c->cfg.clip_param * l_mul;                       // PPO.py:19
    if (int r = launch_loss(s, w.out_pi, old_rows ? old_rows : w.out_old, act, adv, ret, rows, tile0, inv_B, clip, c->cfg.entcoeff,
                            c->cfg.vcoeff, w.d_out, w.loss_part, st)) return r;
    if (c->cfg.value_loss_mode == DPPO_VLOSS_LITERAL_PAIRWISE) {
        // PPO.py:36-39 as written: every v_i against every R_j of its rollout (`groups` time-major rollouts of
        // B_total / groups rows; ret_all = the returns of the whole batch, this call covers rows [row_off, row_off + rows))
        if (int r = launch_value_pairwise(s, w.out_pi, old_rows ? old_rows : w.out_old, ret_all ? ret_all : ret, rows,
                                          row_off, B_total / groups, groups, tile0, inv_B, clip, c->cfg.vcoeff, w.d_out,
                                          w.loss_part, st)) return r;
    }

    GemmEpi tn; tn.accumulate = accumulate;
    const int H = s.in_dim(s.L);
    const float* head_in = s.L ? w.act_pi[s.L - 1] : obs;
    // heads: dWh = h_L^T . d_out, dbh = colsum(d_out)
    const bool tcore = tc_mode(c->cfg.gemm_mode);
    const int terms = th_terms(c);
    const bool fp16x2 = (terms != 0);                                   // the persistent kernel family (two- or one-term)
    const bool bf16 = (c->cfg.gemm_mode == DPPO_GEMM_BF16) && terms == 0;   // round-1 bf16 kernels
    const bool planes = h_planes(c);
    // dW_l (+ db_l) into layer l's split-K partial region
    // (tmp: a free [rows, hmax] buffer -- planes mode joins dY back to fp32 there when layer 0's obs rows cannot be read by TMA)
    // layer 0 of the fp16x2 mode reads the re-pitched observation rows forward_nets prepared (obs_dim % 4 != 0)
    const bool obs_padded = w.obs_pad && terms && c->use_h_dw;
    auto dense_dw = [&](int l, const float* x_in, const float* dY, float* tmp) -> int {
        const int in = s.in_dim(l), out = s.out_dim(l);
        const int ldy = (l == s.L) ? s.Pld : out;
        int ldx = in;
        if (l == 0 && obs_padded) { x_in = w.obs_pad; ldx = w.obs_ld; }
        float* p = w.partial + s.part_off[l];
        if (bf16) {
            if (int r = bf16_gemm_tn_splitk(in, out, rows, x_in, in, dY, ldy, p, out, active[l], s.region[l], accumulate, st))
                return r;
            return tc_colsum(dY, rows, out, ldy, active[l], p + (size_t)in * out, s.region[l], accumulate, st);
        }
        if (planes) {                                       // X = a hidden activation, G = a hidden gradient: fp16 planes
            ThPlanes pl{};
            pl.terms = terms;
            pl.a_planes = l > 0 || (terms == 1 && w.obs_pl != nullptr); pl.a_lo_off = (long long)c->cfg.max_rows * in;
            if (l == 0 && pl.a_planes) x_in = w.obs_pl;       // the bf16 plane forward_nets made of this row block
            pl.c_planes = l < s.L; pl.c_lo_off = (long long)c->cfg.max_rows * out;
            if ((pl.a_planes || th_dw_supported(x_in, ldx, x_in, ldx)) && (pl.c_planes || th_dw_supported(dY, ldy, dY, ldy)))
                return th_gemm_dw(in, out, rows, x_in, ldx, dY, ldy, p, out, p + (size_t)in * out, active[l], s.region[l], accumulate,
                                  st, &pl);
            // obs rows TMA cannot address (obs_dim 3, 17, ...): the tf32 kernels below take fp32 operands
            DPPO_REQUIRE(pl.c_planes && tmp != nullptr, "fp16 planes: layer %d has an fp32 operand TMA cannot address", l);
            const __half* hi = reinterpret_cast<const __half*>(dY);
            if (int r = th_join_planes(hi, hi + pl.c_lo_off, (long long)rows * out, tmp, st, terms)) return r;
            dY = tmp;
        }
        if (fp16x2 && c->use_h_dw && th_dw_supported(x_in, ldx, dY, ldy)) {    // persistent fp16-split kernel, bias column sums fused
            ThPlanes pl{};
            pl.terms = terms;
            return th_gemm_dw(in, out, rows, x_in, ldx, dY, ldy, p, out, p + (size_t)in * out, active[l], s.region[l], accumulate, st,
                              &pl);
        }
        if (l == 0 && obs_padded) x_in = obs;                 // (the fallbacks below take the caller's rows)
        if (tcore) {
            if (tc_tn_tma_supported(x_in, in, dY, ldy))       // TMA-fed kernel, bias column sums fused
                return tc_gemm_tn_tma(in, out, rows, x_in, in, dY, ldy, p, out, p + (size_t)in * out, active[l],
                                      s.region[l], accumulate, st);
            if (int r = tc_gemm_tn_splitk(3, in, out, rows, x_in, in, dY, ldy, p, out, active[l], s.region[l],
                                          accumulate, st)) return r;
            return tc_colsum(dY, rows, out, ldy, active[l], p + (size_t)in * out, s.region[l], accumulate, st);
        }
        tn.slice_stride = s.region[l]; tn.slices = active[l]; tn.bias_out = p + (size_t)in * out;
        return gemm_tn_splitk(in, out, rows, x_in, in, dY, ldy, p, out, tn, st);
    };
    // dX = (dY . W^T) * [h > 0]   (W is [n_out_of_this_product, k] row-major = K-major B operand)
    // mask = activations of hidden layer l_ - 1 (N_ wide); the tensor-core path reads their sign bitmap instead
    auto dense_dx = [&](int l_, int M_, int N_, int K_, const float* dY, int ldy, const float* W, float* dX,
                        const float* mask) -> int {
        if (fp16x2) {
            ThPlanes pl{};                                    // dY of a hidden layer and every dX are planes
            pl.terms = terms;
            pl.a_planes = planes && l_ < s.L; pl.a_lo_off = (long long)c->cfg.max_rows * K_;
            pl.c_planes = planes; pl.c_lo_off = (long long)c->cfg.max_rows * N_;
            return th_gemm(1, M_, N_, K_, dY, nullptr, ldy, w.img_dx_pi[l_], nullptr, dX, nullptr, N_, nullptr, nullptr,
                           0, mask, 1, st, nullptr, w.act_bits[l_ - 1], nullptr, &pl);
        }
        if (tcore) return tc_gemm_kk_img(1, M_, N_, K_, dY, nullptr, ldy, w.img_dx_pi[l_], nullptr, dX, nullptr, N_,
                                         nullptr, nullptr, 0, mask, 1, st, nullptr, w.act_bits[l_ - 1]);
        if (bf16) return bf16_gemm_kk(1, M_, N_, K_, dY, nullptr, ldy, W, nullptr, K_, dX, nullptr, N_, nullptr, nullptr, 0,
                                      mask, 1, st);
        GemmEpi nt; nt.mask = mask;
        return gemm_nt(M_, N_, K_, dY, ldy, W, K_, dX, N_, nt, st);
    };
    if (int r = dense_dw(s.L, head_in, w.d_out, nullptr)) return r;
    if (s.L == 0) return 0;
    // dh_L = (d_out . Wh^T) * [h_L > 0]
    int cur = 0;
    if (int r = dense_dx(s.L, rows, H, s.P, w.d_out, s.Pld, w.head_pi, w.scratch[cur], w.act_pi[s.L - 1])) return r;
    for (int l = s.L - 1; l >= 0; --l) {
        const float* dY = w.scratch[cur];
        const float* x_in = l ? w.act_pi[l - 1] : obs;
        if (int r = dense_dw(l, x_in, dY, w.scratch[cur ^ 1])) return r;
        if (l > 0) {
            if (int r = dense_dx(l, rows, s.h[l - 1], s.h[l], dY, s.h[l], params_pi + s.w_off[l], w.scratch[cur ^ 1],
                                 w.act_pi[l - 1])) return r;
            cur ^= 1;
        }
    }
    return 0;
}

// ---- tail of every update: split-K partials -> flat gradient (+8 loss scalars) -> tower mean over the ranks -> Adam ----
static int reduce_and_apply(dppo_ctx* c, const int* active, int rows, float* params_pi, float* adam_m, float* adam_v,
                            int step, float lr, float beta1, float beta2, float eps, float* grad_out, cudaStream_t st,
                            const StepParams* sp = nullptr, const int* counter = nullptr) {
    const int tiles = (rows + kLossTileRows - 1) / kLossTileRows;
    if (c->world > 1 && peer_available(c)) {
        // NVLink peer path: the rank's gradient goes straight into its exchange buffer; ONE kernel then averages over
        // the ranks through peer loads and applies Adam (PPO.py:53,55-65)
        if (int r = launch_finalize(c->shape, active, c->ws.partial, c->ws.loss_part, tiles, c->cfg.entcoeff, c->cfg.vcoeff,
                                    sp ? peer_exchange_half(c, 0) : peer_exchange_buffer(c), nullptr, st,
                                    sp ? peer_exchange_half(c, 1) : nullptr, sp, counter, 1.0f / grad_prescale(c, rows))) return r;
        return peer_allreduce_adam(c, params_pi, (adam_m && adam_v) ? adam_m : nullptr, adam_v, step, lr, beta1, beta2, eps,
                                   grad_out, st, sp, counter);
    }
    if (!(c->nccl_comm && c->world > 1) && adam_m && adam_v) {
        // one rank: the gradient is final when finalize_kernel has summed the slices -- Adam is applied in the same kernel
        // (sp: the graph-replayed minibatch loop reads its step size from the step record there)
        AdamFused f{params_pi, adam_m, adam_v, step, lr, beta1, beta2, eps};
        return launch_finalize(c->shape, active, c->ws.partial, c->ws.loss_part, tiles, c->cfg.entcoeff, c->cfg.vcoeff,
                               grad_out, nullptr, st, nullptr, sp, counter, 1.0f / grad_prescale(c, rows), &f);
    }
    if (int r = launch_finalize(c->shape, active, c->ws.partial, c->ws.loss_part, tiles, c->cfg.entcoeff, c->cfg.vcoeff,
                                grad_out, nullptr, st, nullptr, nullptr, nullptr, 1.0f / grad_prescale(c, rows))) return r;
    if (c->nccl_comm && c->world > 1)
        if (int r = allreduce_avg(c, grad_out, c->shape.n_params + kLossSlots, st)) return r;
    if (adam_m && adam_v)
        return launch_adam(params_pi, grad_out, adam_m, adam_v, c->shape.n_params, step, lr, beta1, beta2, eps, st, sp, counter);
    return 0;
}

__global__ void reduce_slices_kernel(const float* __restrict__ part, int n_slices, long long stride, long long n,
                                     float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < n_slices; ++k) s += part[(size_t)k * stride + i];
        out[i] = s;
    }
}

}  // namespace dppo

using namespace dppo;

#define CTX_GUARD(ctx)                                                     \
    DPPO_REQUIRE((ctx) != nullptr, "null dppo_ctx");                       \
    t_prof = static_cast<ProfState*>((ctx)->prof);                         \
    DPPO_CHECK_CUDA(cudaSetDevice((ctx)->cfg.device))

extern "C" {

const char* dppo_last_error(void) { return g_err; }
int dppo_version(void) { return 100; }

int64_t dppo_param_count(const dppo_config* cfg) {
    NetShape s;
    if (make_shape(cfg, &s)) return -1;
    return s.n_params;
}

int dppo_param_offsets(const dppo_config* cfg, int64_t* offsets, int capacity) {
    NetShape s;
    if (int r = make_shape(cfg, &s)) return r;
    const int n = 2 * (s.L + 2);
    DPPO_REQUIRE(offsets && capacity >= n, "dppo_param_offsets: need capacity %d", n);
    for (int l = 0; l < s.L; ++l) { offsets[2 * l] = s.w_off[l]; offsets[2 * l + 1] = s.b_off[l]; }
    const int64_t H = s.in_dim(s.L), head = s.w_off[s.L];
    offsets[2 * s.L + 0] = head;                         // Wv [H,1]
    offsets[2 * s.L + 1] = head + H;                     // bv [1]
    offsets[2 * s.L + 2] = head + H + 1;                 // Wp [H,2A]
    offsets[2 * s.L + 3] = head + H + 1 + (int64_t)s.pdw * H; // bp [2A | ncat]
    return n;
}

int64_t dppo_grad_buffer_len(const dppo_config* cfg) {
    const int64_t n = dppo_param_count(cfg);
    return n < 0 ? n : n + kLossSlots;
}

size_t dppo_workspace_bytes(const dppo_config* cfg) {
    NetShape s;
    if (make_shape(cfg, &s)) return 0;
    if (cfg->max_rows < 1) { set_error("max_rows=%d must be >= 1", cfg->max_rows); return 0; }
    return carve(cfg, s, nullptr, nullptr);
}

int dppo_create(const dppo_config* cfg, void* workspace, size_t workspace_bytes, dppo_ctx** out) {
    DPPO_REQUIRE(out != nullptr, "dppo_create: null out");
    *out = nullptr;
    NetShape s;
    if (int r = make_shape(cfg, &s)) return r;
    DPPO_REQUIRE(cfg->max_rows >= 1, "max_rows=%d must be >= 1", cfg->max_rows);
    DPPO_REQUIRE(cfg->value_loss_mode == DPPO_VLOSS_PER_SAMPLE || cfg->value_loss_mode == DPPO_VLOSS_LITERAL_PAIRWISE,
                 "value_loss_mode=%d unknown (DPPO_VLOSS_PER_SAMPLE=0, DPPO_VLOSS_LITERAL_PAIRWISE=1)", cfg->value_loss_mode);
    DPPO_REQUIRE(cfg->gemm_mode == DPPO_GEMM_FP32_SIMT || cfg->gemm_mode == DPPO_GEMM_TF32X3 ||
                     cfg->gemm_mode == DPPO_GEMM_BF16 || cfg->gemm_mode == DPPO_GEMM_FP16X2,
                 "gemm_mode=%d unknown (DPPO_GEMM_FP32_SIMT=0, DPPO_GEMM_TF32X3=1, DPPO_GEMM_BF16=2, DPPO_GEMM_FP16X2=3)",
                 cfg->gemm_mode);
    int ndev = 0;
    DPPO_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    DPPO_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "device %d not present (%d CUDA devices)", cfg->device, ndev);
    DPPO_CHECK_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    DPPO_CHECK_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    DPPO_REQUIRE(prop.major == 10, "libdppo_b200 is built for sm_100a only; device %d is sm_%d%d", cfg->device,
                 prop.major, prop.minor);
    const size_t need = carve(cfg, s, nullptr, nullptr);
    DPPO_REQUIRE(workspace != nullptr && workspace_bytes >= need, "workspace too small: %zu < %zu bytes",
                 workspace_bytes, need);
    DPPO_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    dppo_ctx* c = new (std::nothrow) dppo_ctx();
    DPPO_REQUIRE(c != nullptr, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->cfg = *cfg; c->shape = s; c->workspace = workspace; c->workspace_bytes = workspace_bytes;
    c->num_sms = prop.multiProcessorCount;
    c->prof = new (std::nothrow) ProfState();
    { const char* e = getenv("DPPO_H_DW"); c->use_h_dw = (e && e[0] == '0') ? 0 : 1; }
     // fp16-split dW kernel (tc_dw_h.cu); 0 = tf32 TMA kernel
    { const char* e = getenv("DPPO_H_PLANES"); c->use_h_planes = (e && e[0] == '0') ? 0 : 1; }   // fp16 hi | lo activation planes
    { const char* e = getenv("DPPO_GRAPH"); c->use_graph = (e && e[0] == '0') ? 0 : 1; }
    { const char* e = getenv("DPPO_NO_OLD_CACHE"); c->no_old_cache = (e && e[0] == '1') ? 1 : 0; }   // read once, per context
    carve(cfg, s, (char*)workspace, &c->ws);
    c->world = 1;
    cudaError_t e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 64 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->chunk_ev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->losses_pinned, (size_t)kMaxTrainSteps * kLossSlots * sizeof(float));
    if (e != cudaSuccess) {
        set_error("dppo_create: %s", cudaGetErrorString(e));
        delete c;
        return -2;
    }
    *out = c;
    return 0;
}

int dppo_destroy(dppo_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->cfg.device);
    comm_destroy(c);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 64; ++i) if (c->chunk_ev[i]) cudaEventDestroy(c->chunk_ev[i]);
    if (c->done_ev) cudaEventDestroy(c->done_ev);
    if (c->losses_pinned) cudaFreeHost(c->losses_pinned);
    if (c->step_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)c->step_graph_exec);
    if (ProfState* p = static_cast<ProfState*>(c->prof)) {
        for (int i = 0; i < p->created; ++i) { cudaEventDestroy(p->ev0[i]); cudaEventDestroy(p->ev1[i]); }
        if (t_prof == p) t_prof = nullptr;
        delete p;
    }
    delete c;
    return 0;
}

int dppo_gae(dppo_ctx* c, const float* rew, const float* predv, const float* done, const float* last_v,
             const float* last_done, int T, int N, double gamma, double lam, int mode, int normalize,
             float* adv_out, float* ret_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(rew && predv && done && adv_out && ret_out, "dppo_gae: null buffer");
    return launch_gae(rew, predv, done, last_v, last_done, T, N, gamma, lam, mode, normalize, adv_out, ret_out,
                      (cudaStream_t)stream);
}

int dppo_forward(dppo_ctx* c, const float* params, const float* obs, int B, float* v_out, float* pdflat_out,
                 void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params && obs, "dppo_forward: null buffer");
    DPPO_REQUIRE(B >= 0 && B <= c->cfg.max_rows, "dppo_forward: B=%d exceeds max_rows=%d", B, c->cfg.max_rows);
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = forward_nets(c, params, nullptr, obs, B, st)) return r;
    return launch_split_heads(c->shape, c->ws.out_pi, B, v_out, pdflat_out, st);
}

int dppo_act(dppo_ctx* c, const float* params, const float* obs, int B, uint64_t seed, uint64_t offset,
             float* action_out, float* v_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params && obs && action_out, "dppo_act: null buffer");
    DPPO_REQUIRE(B >= 0 && B <= c->cfg.max_rows, "dppo_act: B=%d exceeds max_rows=%d", B, c->cfg.max_rows);
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = forward_nets(c, params, nullptr, obs, B, st)) return r;
    return launch_sample(c->shape, c->ws.out_pi, B, seed, offset, action_out, v_out, st);
}

int dppo_loss_grad(dppo_ctx* c, const float* params_pi, const float* params_old, const float* obs,
                   const float* act, const float* adv, const float* ret, int B, float l_mul, float* losses_out,
                   float* grad_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params_pi && params_old && obs && act && adv && ret && grad_out, "dppo_loss_grad: null buffer");
    DPPO_REQUIRE(B >= 1 && B <= c->cfg.max_rows, "dppo_loss_grad: B=%d out of [1, max_rows=%d]", B, c->cfg.max_rows);
    cudaStream_t st = (cudaStream_t)stream;
    int active[kMaxLayers + 1];
    active_slices(c, B, active);
    if (int r = loss_grad_rows(c, params_pi, params_old, obs, act, adv, ret, B, 0, B, l_mul, 0, active, st)) return r;
    const int tiles = (B + kLossTileRows - 1) / kLossTileRows;
    return launch_finalize(c->shape, active, c->ws.partial, c->ws.loss_part, tiles, c->cfg.entcoeff, c->cfg.vcoeff, grad_out,
                           losses_out, st, nullptr, nullptr, nullptr, 1.0f / grad_prescale(c, B));
}

int dppo_update(dppo_ctx* c, float* params_pi, const float* params_old, const float* obs, const float* act,
                const float* rew, const float* predv, const float* done, int T, int N, double gamma, double lam,
                int gae_mode, float l_mul, float* adam_m, float* adam_v, int step, float lr, float beta1, float beta2,
                float eps, float* grad_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params_pi && params_old && obs && act && rew && predv && done && grad_out, "dppo_update: null buffer");
    DPPO_REQUIRE(gae_mode == DPPO_GAE_REFERENCE, "dppo_update: only DPPO_GAE_REFERENCE (last_v is not fed)");
    const long long B64 = (long long)T * N;
    DPPO_REQUIRE(T >= 1 && N >= 1 && B64 <= c->cfg.max_rows, "dppo_update: T*N=%lld exceeds max_rows=%d", B64, c->cfg.max_rows);
    const int B = (int)B64;
    cudaStream_t st = (cudaStream_t)stream, side = c->copy_stream;
    // GAE depends only on rew/predv/done: it runs on a side stream underneath the forward GEMMs
    DPPO_CHECK_CUDA(cudaEventRecord(c->done_ev, st));
    DPPO_CHECK_CUDA(cudaStreamWaitEvent(side, c->done_ev, 0));
    if (int r = launch_gae(rew, predv, done, nullptr, nullptr, T, N, gamma, lam, gae_mode, 1, c->ws.gae_adv, c->ws.gae_ret,
                           side)) return r;
    DPPO_CHECK_CUDA(cudaEventRecord(c->chunk_ev[0], side));
    int active[kMaxLayers + 1];
    active_slices(c, B, active);
    if (int r = loss_grad_rows(c, params_pi, params_old, obs, act, c->ws.gae_adv, c->ws.gae_ret, B, 0, B, l_mul, 0, active, st,
                               c->chunk_ev[0], nullptr, c->ws.gae_ret, 0, N)) return r;
    return reduce_and_apply(c, active, B, params_pi, adam_m, adam_v, step, lr, beta1, beta2, eps, grad_out, st);
}

int dppo_update_epochs(dppo_ctx* c, float* params_pi, const float* params_old, const float* obs, const float* act,
                       const float* rew, const float* predv, const float* done, int T, int N, double gamma, double lam,
                       int gae_mode, int n_epochs, int minibatch_rows, int shuffle, uint64_t seed, float l_mul,
                       float* adam_m, float* adam_v, int step0, float lr, float beta1, float beta2, float eps,
                       float* losses_out, float* grad_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params_pi && params_old && obs && act && rew && predv && done && grad_out && adam_m && adam_v,
                 "dppo_update_epochs: null buffer");
    DPPO_REQUIRE(gae_mode == DPPO_GAE_REFERENCE, "dppo_update_epochs: only DPPO_GAE_REFERENCE (last_v is not fed)");
    const long long B64 = (long long)T * N;
    DPPO_REQUIRE(T >= 1 && N >= 1 && B64 <= c->cfg.max_rows, "dppo_update_epochs: T*N=%lld exceeds max_rows=%d", B64,
                 c->cfg.max_rows);
    const int B = (int)B64;
    DPPO_REQUIRE(n_epochs >= 1 && minibatch_rows >= 1, "dppo_update_epochs: n_epochs=%d minibatch_rows=%d", n_epochs,
                 minibatch_rows);
    DPPO_REQUIRE(step0 >= 1, "dppo_update_epochs: step0=%d must be >= 1", step0);
    const int mb = minibatch_rows < B ? minibatch_rows : B;
    const int n_mb = (B + mb - 1) / mb;
    DPPO_REQUIRE(c->cfg.value_loss_mode == DPPO_VLOSS_PER_SAMPLE || (n_mb == 1 && !shuffle),
                 "dppo_update_epochs: DPPO_VLOSS_LITERAL_PAIRWISE pairs the rows of whole rollouts; use whole-buffer minibatches "
                 "without shuffling");
    const NetShape& s = c->shape;
    Workspace& w = c->ws;
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = launch_gae(rew, predv, done, nullptr, nullptr, T, N, gamma, lam, gae_mode, 1, w.gae_adv, w.gae_ret, st))
        return r;
    // gathered minibatch lives in the staging area (also used by dppo_update_host; one call at a time per context)
    float* g_obs = w.stage;
    float* g_act = g_obs + (size_t)mb * s.obs;
    float* g_adv = g_act + (size_t)mb * s.acols;
    float* g_ret = g_adv + mb;
    int active[kMaxLayers + 1];
    active_slices(c, mb, active);
    // oldpi does not change during the call (sync_pis happens in the worker, PPO.py:47 / Worker.py:42): evaluate its heads
    // ONCE over the whole buffer and reuse the rows in every minibatch of every epoch.  The reference re-evaluates oldpi
    // in-graph on every sess.run (PPO.py:22) and gets the same numbers.  DPPO_NO_OLD_CACHE=1 restores the recomputation.
    const bool cache_old = !c->no_old_cache && (long long)n_epochs * n_mb > 1 && (!shuffle || s.hmax() >= s.Pld);
    if (cache_old) {
        if (int r = forward_nets(c, params_old, nullptr, obs, B, st)) return r;          // oldpi through the "pi" slot
        DPPO_CHECK_CUDA(cudaMemcpyAsync(w.out_old, w.out_pi, (size_t)B * s.Pld * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    float* g_old = w.scratch[1];         // gathered oldpi rows of a shuffled minibatch (free until the backward pass)
    const long long total_steps = (long long)n_epochs * n_mb;

    // ---- small shuffled minibatches are launch-bound (~15 kernels of a few microseconds each): ONE minibatch step is captured
    // as a CUDA graph and replayed; what differs from step to step (permutation key, first row, Adam step size, loss slot,
    // exchange epoch) is read by the kernels from a device array of per-step records through a device-side step counter ----
    ProfState* prof = static_cast<ProfState*>(c->prof);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    const bool use_graph = c->use_graph && shuffle && n_mb >= 2 && total_steps <= kMaxGraphSteps && !(prof && prof->timing) &&
                           cap == cudaStreamCaptureStatusNone && !(c->nccl_comm && c->world > 1 && !peer_available(c));
    if (use_graph) {
        StepParams* host_sp = (StepParams*)malloc((size_t)total_steps * sizeof(StepParams));
        DPPO_REQUIRE(host_sp != nullptr, "dppo_update_epochs: out of host memory");
        const bool peer = c->world > 1 && peer_available(c);
        for (int e = 0, i = 0; e < n_epochs; ++e) {
            const PermKey key = make_perm_key((uint32_t)B, seed, (uint64_t)e);
            for (int k = 0; k < n_mb; ++k, ++i) {
                StepParams& q = host_sp[i];
                memset(&q, 0, sizeof(q));
                q.key = key; q.first = k * mb; q.rows = (B - k * mb < mb) ? B - k * mb : mb;
                q.lr_t = (float)adam_lr_t(lr, beta1, beta2, step0 + i);
                q.loss_slot = i;
                if (peer) { c->epoch += 1; q.peer_epoch = c->epoch; q.peer_sel = (int)(c->epoch & 1); }
            }
        }
        // pageable source: the runtime stages it before returning, so host_sp can be freed right away
        cudaError_t e1 = cudaMemcpyAsync(w.step_params, host_sp, (size_t)total_steps * sizeof(StepParams), cudaMemcpyHostToDevice, st);
        free(host_sp);
        DPPO_CHECK_CUDA(e1);
        DPPO_CHECK_CUDA(cudaMemsetAsync(w.step_counter, 0, sizeof(int), st));
        // one graph per (minibatch rows, buffers, hyper-parameters); re-captured when any of them changes
        struct GraphKey { const void* p[10]; int mb, B, cache_old, world; float l_mul, b1, b2, eps; } key;
        memset(&key, 0, sizeof(key));
        key.p[0] = params_pi; key.p[1] = params_old; key.p[2] = obs; key.p[3] = act; key.p[4] = adam_m; key.p[5] = adam_v;
        key.p[6] = grad_out; key.p[7] = losses_out; key.mb = mb; key.B = B; key.cache_old = cache_old ? 1 : 0; key.world = c->world;
        key.l_mul = l_mul; key.b1 = beta1; key.b2 = beta2; key.eps = eps;
        static_assert(sizeof(GraphKey) <= sizeof(c->step_graph_key), "graph key storage");
        if (!c->step_graph_exec || memcmp(&key, c->step_graph_key, sizeof(key)) != 0) {
            if (c->step_graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)c->step_graph_exec); c->step_graph_exec = nullptr; }
            cudaStream_t cs = c->copy_stream;                // the caller's stream may be the legacy default stream: capture elsewhere
            long long before[KK_COUNT];
            for (int k = 0; k < KK_COUNT; ++k) before[k] = prof ? prof->launches[k] : 0;
            DPPO_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
            int rc = launch_gather_rows(make_perm_key((uint32_t)B, seed, 0), 0, mb, s.obs, s.acols, obs, act, w.gae_adv, w.gae_ret, g_obs,
                                        g_act, g_adv, g_ret, cache_old ? w.out_old : nullptr, g_old, s.Pld, cs, w.step_params, w.step_counter);
            if (!rc) rc = loss_grad_rows(c, params_pi, params_old, g_obs, g_act, g_adv, g_ret, mb, 0, mb, l_mul, 0, active, cs, nullptr,
                                         cache_old ? g_old : nullptr);
            if (!rc) rc = reduce_and_apply(c, active, mb, params_pi, adam_m, adam_v, step0, lr, beta1, beta2, eps, grad_out, cs,
                                           w.step_params, w.step_counter);
            if (!rc) rc = launch_step_advance(grad_out + s.n_params, losses_out, w.step_params, w.step_counter, cs);
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            DPPO_CHECK_CUDA(ce);
            cudaGraphExec_t exec = nullptr;
            const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            DPPO_CHECK_CUDA(ie);
            c->step_graph_exec = exec;
            memcpy(c->step_graph_key, &key, sizeof(key));
            for (int k = 0; k < KK_COUNT; ++k) {
                c->step_graph_counts[k] = prof ? prof->launches[k] - before[k] : 0;
                if (prof) prof->launches[k] = before[k];     // captured, not launched
            }
        }
        for (int e = 0, i = 0; e < n_epochs; ++e) {
            const PermKey pkey = make_perm_key((uint32_t)B, seed, (uint64_t)e);
            for (int k = 0; k < n_mb; ++k, ++i) {
                const int r0 = k * mb, rows = (B - r0 < mb) ? B - r0 : mb;
                if (rows == mb) {
                    DPPO_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)c->step_graph_exec, st));
                    if (prof) for (int q = 0; q < KK_COUNT; ++q) prof->launches[q] += c->step_graph_counts[q];
                    continue;
                }
                // the ragged last minibatch of an epoch: direct launches, same per-step records
                if (int r = launch_gather_rows(pkey, r0, rows, s.obs, s.acols, obs, act, w.gae_adv, w.gae_ret, g_obs, g_act, g_adv, g_ret,
                                               cache_old ? w.out_old : nullptr, g_old, s.Pld, st)) return r;
                int active_r[kMaxLayers + 1];
                active_slices(c, rows, active_r);
                if (int r = loss_grad_rows(c, params_pi, params_old, g_obs, g_act, g_adv, g_ret, rows, 0, rows, l_mul, 0, active_r, st,
                                           nullptr, cache_old ? g_old : nullptr)) return r;
                if (int r = reduce_and_apply(c, active_r, rows, params_pi, adam_m, adam_v, step0 + i, lr, beta1, beta2, eps, grad_out, st,
                                             w.step_params, w.step_counter)) return r;
                if (int r = launch_step_advance(grad_out + s.n_params, losses_out, w.step_params, w.step_counter, st)) return r;
            }
        }
        return 0;
    }

    int step = step0, slot = 0;
    for (int e = 0; e < n_epochs; ++e) {
        const PermKey key = make_perm_key((uint32_t)B, seed, (uint64_t)e);
        for (int k = 0; k < n_mb; ++k, ++step, ++slot) {
            const int r0 = k * mb, rows = (B - r0 < mb) ? B - r0 : mb;
            const float *x = obs + (size_t)r0 * s.obs, *a = act + (size_t)r0 * s.acols, *ad = w.gae_adv + r0,
                        *rt = w.gae_ret + r0;
            const float* old_rows = cache_old ? w.out_old + (size_t)r0 * s.Pld : nullptr;
            if (shuffle) {
                if (int r = launch_gather_rows(key, r0, rows, s.obs, s.acols, obs, act, w.gae_adv, w.gae_ret, g_obs, g_act,
                                               g_adv, g_ret, cache_old ? w.out_old : nullptr, g_old, s.Pld, st)) return r;
                x = g_obs; a = g_act; ad = g_adv; rt = g_ret;
                if (cache_old) old_rows = g_old;
            }
            if (int r = loss_grad_rows(c, params_pi, params_old, x, a, ad, rt, rows, 0, rows, l_mul, 0, active, st, nullptr,
                                       old_rows, rt, 0, (n_mb == 1 && !shuffle) ? N : 1)) return r;
            if (int r = reduce_and_apply(c, active, rows, params_pi, adam_m, adam_v, step, lr, beta1, beta2, eps, grad_out, st))
                return r;
            if (losses_out)
                DPPO_CHECK_CUDA(cudaMemcpyAsync(losses_out + (size_t)slot * kLossSlots, grad_out + s.n_params,
                                                kLossSlots * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
    }
    return 0;
}

int dppo_minibatch_perm(dppo_ctx* c, int n, uint64_t seed, uint64_t epoch, int32_t* perm_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(n >= 0 && (n == 0 || perm_out), "dppo_minibatch_perm: n=%d", n);
    if (n == 0) return 0;
    return launch_perm_fill(make_perm_key((uint32_t)n, seed, epoch), perm_out, (cudaStream_t)stream);
}

int dppo_comm_unique_id(void* id_out) {
    DPPO_REQUIRE(id_out != nullptr, "dppo_comm_unique_id: null buffer");
    return comm_unique_id(id_out);
}
int dppo_comm_init(dppo_ctx* c, int rank, int world, const void* id) {
    CTX_GUARD(c);
    DPPO_REQUIRE(id != nullptr, "dppo_comm_init: null id");
    return comm_init(c, rank, world, id);
}
int dppo_comm_peer_export(dppo_ctx* c, int rank, int world, void* handle_out) {
    CTX_GUARD(c);
    DPPO_REQUIRE(handle_out != nullptr, "dppo_comm_peer_export: null buffer");
    return comm_peer_export(c, rank, world, handle_out);
}
int dppo_comm_peer_import(dppo_ctx* c, const void* all_handles) {
    CTX_GUARD(c);
    DPPO_REQUIRE(all_handles != nullptr, "dppo_comm_peer_import: null buffer");
    return comm_peer_import(c, all_handles);
}
int dppo_comm_destroy(dppo_ctx* c) {
    DPPO_REQUIRE(c != nullptr, "null dppo_ctx");
    return comm_destroy(c);
}
int dppo_allreduce_avg(dppo_ctx* c, float* buf, int64_t n, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(buf != nullptr && n >= 0, "dppo_allreduce_avg: bad buffer");
    return allreduce_avg(c, buf, n, (cudaStream_t)stream);
}

int dppo_adam_apply(dppo_ctx* c, float* params, const float* grad, float* m, float* v, int64_t n, int step,
                    float lr, float beta1, float beta2, float eps, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params && grad && m && v && n >= 0, "dppo_adam_apply: bad buffer");
    return launch_adam(params, grad, m, v, n, step, lr, beta1, beta2, eps, (cudaStream_t)stream);
}

int dppo_sync_old(dppo_ctx* c, const float* params_pi, float* params_old, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params_pi && params_old, "dppo_sync_old: null buffer");
    ProfScope ps(KK_MISC, (cudaStream_t)stream);
    DPPO_CHECK_CUDA(cudaMemcpyAsync(params_old, params_pi, (size_t)c->shape.n_params * sizeof(float),
                                    cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int dppo_update_host(dppo_ctx* c, float* params_pi, const float* params_old, const float* obs_h,
                     const float* act_h, const float* rew_h, const float* predv_h, const float* done_h, int T,
                     int N, double gamma, double lam, int gae_mode, float l_mul, float* adam_m, float* adam_v, int step,
                     float lr, float beta1, float beta2, float eps, float* losses_out_h, float* grad_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params_pi && params_old && obs_h && act_h && rew_h && predv_h && done_h && grad_out,
                 "dppo_update_host: null buffer");
    DPPO_REQUIRE(gae_mode == DPPO_GAE_REFERENCE, "dppo_update_host: only DPPO_GAE_REFERENCE (last_v is not fed)");
    const long long B64 = (long long)T * N;
    DPPO_REQUIRE(T >= 1 && N >= 1 && B64 <= c->cfg.max_rows, "dppo_update_host: T*N=%lld exceeds max_rows=%d", B64,
                 c->cfg.max_rows);
    const int B = (int)B64;
    const NetShape& s = c->shape;
    Workspace& w = c->ws;
    cudaStream_t st = (cudaStream_t)stream, cp = c->copy_stream;
    float* d_obs = w.stage;
    float* d_act = d_obs + (size_t)B * s.obs;
    float* d_rew = d_act + (size_t)B * s.acols;
    float* d_pv = d_rew + B;
    float* d_done = d_pv + B;

    // the copy stream must not overwrite staging that earlier work on `st` still reads
    DPPO_CHECK_CUDA(cudaEventRecord(c->done_ev, st));
    DPPO_CHECK_CUDA(cudaStreamWaitEvent(cp, c->done_ev, 0));
    // small arrays first: GAE needs whole rollouts
    DPPO_CHECK_CUDA(cudaMemcpyAsync(d_rew, rew_h, (size_t)B * 4, cudaMemcpyHostToDevice, cp));
    DPPO_CHECK_CUDA(cudaMemcpyAsync(d_pv, predv_h, (size_t)B * 4, cudaMemcpyHostToDevice, cp));
    DPPO_CHECK_CUDA(cudaMemcpyAsync(d_done, done_h, (size_t)B * 4, cudaMemcpyHostToDevice, cp));
    DPPO_CHECK_CUDA(cudaEventRecord(c->chunk_ev[0], cp));
    // row chunks of the big arrays; chunk c's fwd+loss+bwd overlaps chunk c+1's H2D copy
    // chunk = one full wave of forward CTAs (2 CTAs/SM x 148 SMs over the widest layer's column tiles)
    int col_tiles = 1;
    for (int l = 0; l < s.L; ++l) { const int t = (s.h[l] + 127) / 128; if (t > col_tiles) col_tiles = t; }
    int rows_per = (2 * kPlanSMs / col_tiles) * 128;
    if (rows_per < 4096) rows_per = 4096;
    while ((B + rows_per - 1) / rows_per > 32) rows_per *= 2;
    // chunk list: uniform chunks, then the tail tapers (1/2, 1/4, 1/4 of a chunk) so that the work left
    // after the last byte arrives is a quarter chunk instead of a whole one
    int chunk_r0[64], chunk_rows[64], n_chunks = 0;
    {
        int r0 = 0;
        while (B - r0 > rows_per) { chunk_r0[n_chunks] = r0; chunk_rows[n_chunks++] = rows_per; r0 += rows_per; }
        int left = B - r0;
        // every chunk but the last must be a whole number of loss tiles (loss_part is indexed by r0 / kLossTileRows):
        // only the final chunk may be ragged
        const int q = ((left / 4 + kLossTileRows - 1) / kLossTileRows) * kLossTileRows;
        const int first = ((left - 2 * q + kLossTileRows - 1) / kLossTileRows) * kLossTileRows;
        if (n_chunks > 0 && q >= 1024 && first > 0 && left - first - q > 0) {
            const int parts[3] = {first, q, left - first - q};
            for (int i = 0; i < 3; ++i) { chunk_r0[n_chunks] = r0; chunk_rows[n_chunks++] = parts[i]; r0 += parts[i]; }
        } else {
            chunk_r0[n_chunks] = r0; chunk_rows[n_chunks++] = left;
        }
    }
    DPPO_REQUIRE(n_chunks <= 63, "dppo_update_host: too many chunks");
    for (int k = 0; k < n_chunks; ++k) {
        const int r0 = chunk_r0[k], rows = chunk_rows[k];
        DPPO_CHECK_CUDA(cudaMemcpyAsync(d_obs + (size_t)r0 * s.obs, obs_h + (size_t)r0 * s.obs,
                                        (size_t)rows * s.obs * 4, cudaMemcpyHostToDevice, cp));
        DPPO_CHECK_CUDA(cudaMemcpyAsync(d_act + (size_t)r0 * s.acols, act_h + (size_t)r0 * s.acols,
                                        (size_t)rows * s.acols * 4, cudaMemcpyHostToDevice, cp));
        DPPO_CHECK_CUDA(cudaEventRecord(c->chunk_ev[k + 1], cp));
    }
    int active[kMaxLayers + 1];
    active_slices(c, rows_per < B ? rows_per : B, active);
    DPPO_CHECK_CUDA(cudaStreamWaitEvent(st, c->chunk_ev[0], 0));
    if (int r = launch_gae(d_rew, d_pv, d_done, nullptr, nullptr, T, N, gamma, lam, gae_mode, 1, w.gae_adv,
                           w.gae_ret, st)) return r;
    for (int k = 0; k < n_chunks; ++k) {
        const int r0 = chunk_r0[k], rows = chunk_rows[k];
        DPPO_CHECK_CUDA(cudaStreamWaitEvent(st, c->chunk_ev[k + 1], 0));
        if (int r = loss_grad_rows(c, params_pi, params_old, d_obs + (size_t)r0 * s.obs, d_act + (size_t)r0 * s.acols,
                                   w.gae_adv + r0, w.gae_ret + r0, rows, r0 / kLossTileRows, B, l_mul, k > 0, active, st,
                                   nullptr, nullptr, w.gae_ret, r0, N))
            return r;
    }
    if (int r = reduce_and_apply(c, active, B, params_pi, adam_m, adam_v, step, lr, beta1, beta2, eps, grad_out, st)) return r;
    DPPO_CHECK_CUDA(cudaMemcpyAsync(c->losses_pinned, grad_out + s.n_params, kLossSlots * sizeof(float),
                                    cudaMemcpyDeviceToHost, st));
    DPPO_CHECK_CUDA(cudaStreamSynchronize(st));
    if (losses_out_h) memcpy(losses_out_h, c->losses_pinned, kLossSlots * sizeof(float));
    return 0;
}

int dppo_train_host(dppo_ctx* c, float* params_pi, const float* params_old, int n_towers, const float* const* obs_h,
                    const void* const* act_h, const float* const* adv_h, const float* const* ret_h, const int* rows_of,
                    int n_steps, float l_mul, float* adam_m, float* adam_v, int step0, float lr, float beta1, float beta2,
                    float eps, float* losses_out_h, float* grad_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(params_pi && params_old && obs_h && act_h && adv_h && ret_h && rows_of && grad_out && adam_m && adam_v,
                 "dppo_train_host: null argument");
    DPPO_REQUIRE(n_towers >= 1 && n_towers <= 1024, "dppo_train_host: n_towers=%d out of [1, 1024]", n_towers);
    DPPO_REQUIRE(n_steps >= 1 && n_steps <= kMaxTrainSteps, "dppo_train_host: n_steps=%d out of [1, %d]", n_steps, kMaxTrainSteps);
    DPPO_REQUIRE(step0 >= 1, "dppo_train_host: step0=%d must be >= 1 (t after increment)", step0);
    DPPO_REQUIRE(c->cfg.value_loss_mode == DPPO_VLOSS_PER_SAMPLE,
                 "dppo_train_host: DPPO_VLOSS_LITERAL_PAIRWISE needs the [T,N] rollout structure: use dppo_update");
    long long B64 = 0;
    for (int k = 0; k < n_towers; ++k) {
        DPPO_REQUIRE(rows_of[k] >= 1 && obs_h[k] && act_h[k] && adv_h[k] && ret_h[k], "dppo_train_host: tower %d is empty", k);
        B64 += rows_of[k];
    }
    DPPO_REQUIRE(B64 <= c->cfg.max_rows, "dppo_train_host: %lld rows exceed max_rows=%d", B64, c->cfg.max_rows);
    const int B = (int)B64;
    const NetShape& s = c->shape;
    Workspace& w = c->ws;
    cudaStream_t st = (cudaStream_t)stream, cp = c->copy_stream;
    float* d_obs = w.stage;
    float* d_act = d_obs + (size_t)B * s.obs;
    float* d_adv = d_act + (size_t)B * s.acols;
    float* d_ret = d_adv + B;

    DPPO_CHECK_CUDA(cudaEventRecord(c->done_ev, st));          // earlier work on `st` may still read the staging area
    DPPO_CHECK_CUDA(cudaStreamWaitEvent(cp, c->done_ev, 0));
    // the small per-row arrays first, then the big ones in row chunks of one full wave of forward CTAs
    for (int k = 0, r0 = 0; k < n_towers; r0 += rows_of[k], ++k) {
        DPPO_CHECK_CUDA(cudaMemcpyAsync(d_adv + r0, adv_h[k], (size_t)rows_of[k] * 4, cudaMemcpyHostToDevice, cp));
        DPPO_CHECK_CUDA(cudaMemcpyAsync(d_ret + r0, ret_h[k], (size_t)rows_of[k] * 4, cudaMemcpyHostToDevice, cp));
    }
    int col_tiles = 1;
    for (int l = 0; l < s.L; ++l) { const int t = (s.h[l] + 127) / 128; if (t > col_tiles) col_tiles = t; }
    int rows_per = (2 * kPlanSMs / col_tiles) * 128;
    if (rows_per < 4096) rows_per = 4096;
    while ((B + rows_per - 1) / rows_per > 32) rows_per *= 2;
    int chunk_r0[64], chunk_rows[64], n_chunks = 0;
    {
        int r0 = 0;
        while (B - r0 > rows_per) { chunk_r0[n_chunks] = r0; chunk_rows[n_chunks++] = rows_per; r0 += rows_per; }
        const int left = B - r0;
        const int q = ((left / 4 + kLossTileRows - 1) / kLossTileRows) * kLossTileRows;
        const int first = ((left - 2 * q + kLossTileRows - 1) / kLossTileRows) * kLossTileRows;
        if (n_chunks > 0 && q >= 1024 && first > 0 && left - first - q > 0) {
            const int parts[3] = {first, q, left - first - q};
            for (int i = 0; i < 3; ++i) { chunk_r0[n_chunks] = r0; chunk_rows[n_chunks++] = parts[i]; r0 += parts[i]; }
        } else {
            chunk_r0[n_chunks] = r0; chunk_rows[n_chunks++] = left;
        }
    }
    DPPO_REQUIRE(n_chunks <= 63, "dppo_train_host: too many chunks");
    {
        int seg = 0, seg_r0 = 0;                                  // tower that holds global row `r`, and its first global row
        for (int k = 0; k < n_chunks; ++k) {
            int r = chunk_r0[k];
            const int end = r + chunk_rows[k];
            while (r < end) {                                     // a chunk may straddle towers: one copy per piece
                while (r >= seg_r0 + rows_of[seg]) { seg_r0 += rows_of[seg]; ++seg; }
                const int off = r - seg_r0;
                const int n = (end - r < rows_of[seg] - off) ? end - r : rows_of[seg] - off;
                DPPO_CHECK_CUDA(cudaMemcpyAsync(d_obs + (size_t)r * s.obs, obs_h[seg] + (size_t)off * s.obs,
                                                (size_t)n * s.obs * 4, cudaMemcpyHostToDevice, cp));
                DPPO_CHECK_CUDA(cudaMemcpyAsync(d_act + (size_t)r * s.acols,
                                                static_cast<const float*>(act_h[seg]) + (size_t)off * s.acols,
                                                (size_t)n * s.acols * 4, cudaMemcpyHostToDevice, cp));
                r += n;
            }
            DPPO_CHECK_CUDA(cudaEventRecord(c->chunk_ev[k + 1], cp));
        }
    }
    int active[kMaxLayers + 1];
    // step 0: chunk k's forward + loss + backward overlaps chunk k+1's upload
    active_slices(c, rows_per < B ? rows_per : B, active);
    for (int k = 0; k < n_chunks; ++k) {
        const int r0 = chunk_r0[k], rows = chunk_rows[k];
        DPPO_CHECK_CUDA(cudaStreamWaitEvent(st, c->chunk_ev[k + 1], 0));
        if (int r = loss_grad_rows(c, params_pi, params_old, d_obs + (size_t)r0 * s.obs, d_act + (size_t)r0 * s.acols,
                                   d_adv + r0, d_ret + r0, rows, r0 / kLossTileRows, B, l_mul, k > 0, active, st)) return r;
    }
    if (int r = reduce_and_apply(c, active, B, params_pi, adam_m, adam_v, step0, lr, beta1, beta2, eps, grad_out, st)) return r;
    DPPO_CHECK_CUDA(cudaMemcpyAsync(c->losses_pinned, grad_out + s.n_params, kLossSlots * sizeof(float), cudaMemcpyDeviceToHost, st));
    // steps 1..: the whole staged batch in one pass; oldpi's heads are cached once every row of them exists
    bool old_cached = (n_chunks == 1);
    if (n_steps > 1) active_slices(c, B, active);
    for (int i = 1; i < n_steps; ++i) {
        if (int r = loss_grad_rows(c, params_pi, params_old, d_obs, d_act, d_adv, d_ret, B, 0, B, l_mul, 0, active, st, nullptr,
                                   old_cached ? w.out_old : nullptr)) return r;
        old_cached = true;
        if (int r = reduce_and_apply(c, active, B, params_pi, adam_m, adam_v, step0 + i, lr, beta1, beta2, eps, grad_out, st))
            return r;
        DPPO_CHECK_CUDA(cudaMemcpyAsync(c->losses_pinned + (size_t)i * kLossSlots, grad_out + s.n_params,
                                        kLossSlots * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    DPPO_CHECK_CUDA(cudaStreamSynchronize(st));
    if (losses_out_h) memcpy(losses_out_h, c->losses_pinned, (size_t)n_steps * kLossSlots * sizeof(float));
    return 0;
}

// ---- checkpoint / resume (Others/tf_util.py:271-279) -----------------------------------------------------------------
struct StateHeader {
    char magic[8];
    int32_t version, obs_dim, act_dim, pd_type, n_hidden;
    int32_t hidden[DPPO_MAX_HIDDEN];
    int32_t has_adam;
    int64_t n_params, step;
    char pad[128 - 8 - 5 * 4 - DPPO_MAX_HIDDEN * 4 - 4 - 16];
};
static_assert(sizeof(StateHeader) == 128, "checkpoint header is 128 bytes");

static void fill_header(const dppo_ctx* c, StateHeader* h, int has_adam, int64_t step) {
    memset(h, 0, sizeof(*h));
    memcpy(h->magic, "DPPOB200", 8);
    h->version = 1; h->obs_dim = c->cfg.obs_dim; h->act_dim = c->cfg.act_dim; h->pd_type = c->cfg.pd_type;
    h->n_hidden = c->cfg.n_hidden;
    for (int i = 0; i < c->cfg.n_hidden; ++i) h->hidden[i] = c->cfg.hidden[i];
    h->has_adam = has_adam; h->n_params = c->shape.n_params; h->step = step;
}

int dppo_save_state(dppo_ctx* c, const char* path, const float* params_pi, const float* params_old, const float* adam_m,
                    const float* adam_v, int64_t step, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(path && params_pi && params_old, "dppo_save_state: null argument");
    DPPO_REQUIRE((adam_m == nullptr) == (adam_v == nullptr), "dppo_save_state: pass both Adam slots or neither");
    const size_t n = (size_t)c->shape.n_params;
    const int arrays = adam_m ? 4 : 2;
    float* host = (float*)malloc(n * sizeof(float) * arrays);
    DPPO_REQUIRE(host != nullptr, "dppo_save_state: out of host memory");
    const float* src[4] = {params_pi, params_old, adam_m, adam_v};
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < arrays; ++i) {
        const cudaError_t e = cudaMemcpyAsync(host + (size_t)i * n, src[i], n * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) { free(host); set_error("dppo_save_state: %s", cudaGetErrorString(e)); return -2; }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { free(host); set_error("dppo_save_state: stream synchronise failed"); return -2; }
    StateHeader h;
    fill_header(c, &h, adam_m ? 1 : 0, step);
    char tmp[4096];
    snprintf(tmp, sizeof(tmp), "%s.tmp", path);
    FILE* f = fopen(tmp, "wb");
    if (!f) { free(host); set_error("dppo_save_state: cannot open %s", tmp); return -4; }
    const bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(host, sizeof(float), n * arrays, f) == n * arrays;
    const bool closed = fclose(f) == 0;
    free(host);
    if (!ok || !closed || rename(tmp, path) != 0) { remove(tmp); set_error("dppo_save_state: write to %s failed", path); return -4; }
    return 0;
}

int dppo_load_state(dppo_ctx* c, const char* path, float* params_pi, float* params_old, float* adam_m, float* adam_v,
                    int64_t* step_out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(path && params_pi && params_old, "dppo_load_state: null argument");
    DPPO_REQUIRE((adam_m == nullptr) == (adam_v == nullptr), "dppo_load_state: pass both Adam slots or neither");
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("dppo_load_state: cannot open %s", path); return -4; }
    StateHeader h, want;
    if (fread(&h, sizeof(h), 1, f) != 1) { fclose(f); set_error("dppo_load_state: %s is truncated", path); return -4; }
    fill_header(c, &want, h.has_adam, h.step);
    if (memcmp(h.magic, "DPPOB200", 8) != 0 || h.version != 1) { fclose(f); set_error("dppo_load_state: %s is not a dppo_b200 checkpoint", path); return -4; }
    if (memcmp(&h, &want, offsetof(StateHeader, has_adam)) != 0 || h.n_params != want.n_params) {
        fclose(f);
        set_error("dppo_load_state: %s holds a different network (obs %d act %d pd %d layers %d, %lld params; context has %lld)",
                  path, h.obs_dim, h.act_dim, h.pd_type, h.n_hidden, (long long)h.n_params, (long long)want.n_params);
        return -1;
    }
    if (adam_m && !h.has_adam) { fclose(f); set_error("dppo_load_state: %s holds no Adam slots", path); return -1; }
    const size_t n = (size_t)c->shape.n_params;
    const int arrays = adam_m ? 4 : 2;
    float* host = nullptr;
    if (cudaMallocHost((void**)&host, n * sizeof(float) * arrays) != cudaSuccess) { fclose(f); set_error("dppo_load_state: out of pinned host memory"); return -2; }
    const bool ok = fread(host, sizeof(float), n * arrays, f) == n * arrays;
    fclose(f);
    if (!ok) { cudaFreeHost(host); set_error("dppo_load_state: %s is truncated", path); return -4; }
    float* dst[4] = {params_pi, params_old, adam_m, adam_v};
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < arrays && e == cudaSuccess; ++i)
        e = cudaMemcpyAsync(dst[i], host + (size_t)i * n, n * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeHost(host);
    if (e != cudaSuccess) { set_error("dppo_load_state: %s", cudaGetErrorString(e)); return -2; }
    if (step_out) *step_out = h.step;
    return 0;
}

int dppo_diag_gaussian(dppo_ctx* c, int op, const float* flat, const float* arg, int B, int A, uint64_t seed,
                       uint64_t offset, float* out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(B >= 0 && A >= 1, "dppo_diag_gaussian: B=%d A=%d", B, A);
    return launch_diag_gaussian(op, flat, arg, B, A, seed, offset, out, (cudaStream_t)stream);
}

int dppo_categorical(dppo_ctx* c, int op, const float* logits, const void* arg, int B, int ncat, uint64_t seed,
                     uint64_t offset, void* out, void* stream) {
    CTX_GUARD(c);
    DPPO_REQUIRE(B >= 0 && ncat >= 1, "dppo_categorical: B=%d ncat=%d", B, ncat);
    return launch_categorical(op, logits, arg, B, ncat, seed, offset, out, (cudaStream_t)stream);
}

int dppo_profile_reset(dppo_ctx* c, int enable_timing) {
    CTX_GUARD(c);
    ProfState* p = static_cast<ProfState*>(c->prof);
    DPPO_REQUIRE(p != nullptr, "dppo_profile_reset: no accounting state");
    for (int i = 0; i < p->used; ++i) cudaEventSynchronize(p->ev1[i]);
    p->used = 0;
    for (int k = 0; k < KK_COUNT; ++k) { p->launches[k] = 0; p->ms[k] = 0.0; }
    p->timing = enable_timing != 0;
    return 0;
}

int dppo_profile_read(dppo_ctx* c, int64_t* launches, double* ms, int capacity) {
    CTX_GUARD(c);
    ProfState* p = static_cast<ProfState*>(c->prof);
    DPPO_REQUIRE(p != nullptr, "dppo_profile_read: no accounting state");
    DPPO_REQUIRE(launches && ms && capacity >= KK_COUNT, "dppo_profile_read: need capacity %d", (int)KK_COUNT);
    for (int i = 0; i < p->used; ++i) {
        DPPO_CHECK_CUDA(cudaEventSynchronize(p->ev1[i]));
        float t = 0.f;
        DPPO_CHECK_CUDA(cudaEventElapsedTime(&t, p->ev0[i], p->ev1[i]));
        p->ms[p->kind_of[i]] += (double)t;
    }
    p->used = 0;
    for (int k = 0; k < KK_COUNT; ++k) { launches[k] = p->launches[k]; ms[k] = p->ms[k]; }
    return KK_COUNT;
}

int dppo_dbg_gemm(dppo_ctx* c, int mode, int M, int N, int K, const float* A, const float* B, float* C,
                  const float* bias, const float* mask, int relu, float* bias_out, void* stream) {
    CTX_GUARD(c);
    cudaStream_t st = (cudaStream_t)stream;
    GemmEpi e; e.bias = bias; e.relu = relu; e.mask = mask;
    // 192 / 194 / 195 / 196 / 197 / 199 / 201-203: the persistent kernels of modes 92 ... 103 in their ONE-TERM mode (bf16 operands)
    int terms = 2;
    if (mode >= 192 && mode <= 203) { terms = 1; mode -= 100; }
    if (mode >= 64 && mode <= 66) {
        // bf16 tensor-core products: 64 C = relu?(A[M,K].Bt[N,K]^T + bias); 65 same product * (mask>0); 66 C = A[K,M]^T.B[K,N]
        if (mode == 64) return bf16_gemm_kk(0, M, N, K, A, nullptr, K, B, nullptr, K, C, nullptr, N, bias, nullptr, relu, nullptr, 1, st);
        if (mode == 65) return bf16_gemm_kk(1, M, N, K, A, nullptr, K, B, nullptr, K, C, nullptr, N, nullptr, nullptr, 0, mask, 1, st);
        const int64_t stride = ((int64_t)M * N + N + 3) & ~3LL;
        int slices = (int)(c->shape.partial_floats / stride);
        if (slices > 5) slices = 5;
        DPPO_REQUIRE(slices >= 1, "dppo_dbg_gemm: M*N+N exceeds the partial buffer");
        if (int r = bf16_gemm_tn_splitk(M, N, K, A, M, B, N, c->ws.partial, N, slices, stride, 0, st)) return r;
        ProfScope ps(KK_MISC, st);
        reduce_slices_kernel<<<256, 256, 0, st>>>(c->ws.partial, slices, stride, (long long)M * N, C);
        DPPO_LAUNCH_CHECK();
        return 0;
    }
    if (mode == 99 || (mode >= 101 && mode <= 103)) {
        // persistent fp16-split dW product (tc_dw_h.cu): C[M,N] = A[K,M]^T . B[K,N], bias_out[N] = column sums of B
        // 101 / 102 / 103: B / A / both operands first converted to fp16 hi | lo planes (scratch[1] / scratch[0])
        ThPlanes pl{};
        pl.terms = terms;
        const long long cap = (long long)c->cfg.max_rows * (c->shape.hmax() > 0 ? c->shape.hmax() : 1);
        if (mode >= 101) {
            pl.a_planes = (mode & 2) ? 1 : 0; pl.c_planes = (mode & 1) ? 1 : 0;
            pl.a_lo_off = ((long long)K * M + 7) & ~7LL; pl.c_lo_off = ((long long)K * N + 7) & ~7LL;
            DPPO_REQUIRE(2 * pl.a_lo_off <= 2 * cap && 2 * pl.c_lo_off <= 2 * cap, "dppo_dbg_gemm: operand planes exceed the scratch buffers");
            if (pl.a_planes) {
                __half* hi = reinterpret_cast<__half*>(c->ws.scratch[0]);
                if (int r = th_split_planes(A, (long long)K * M, hi, hi + pl.a_lo_off, st, terms)) return r;
                A = c->ws.scratch[0];
            }
            if (pl.c_planes) {
                __half* hi = reinterpret_cast<__half*>(c->ws.scratch[1]);
                if (int r = th_split_planes(B, (long long)K * N, hi, hi + pl.c_lo_off, st, terms)) return r;
                B = c->ws.scratch[1];
            }
        }
        const int64_t stride = ((int64_t)M * N + N + 3) & ~3LL;
        int slices = (int)(c->shape.partial_floats / stride);
        const int want = (K + 511) / 512;
        if (slices > want) slices = want;
        DPPO_REQUIRE(slices >= 1, "dppo_dbg_gemm: M*N+N exceeds the partial buffer");
        DPPO_REQUIRE(th_dw_supported(A, M, B, N), "dppo_dbg_gemm: the fp16 dW kernel needs 16-byte aligned rows");
        if (int r = th_gemm_dw(M, N, K, A, M, B, N, c->ws.partial, N, c->ws.partial + (int64_t)M * N, slices, stride, 0, st, &pl))
            return r;
        ProfScope ps(KK_MISC, st);
        reduce_slices_kernel<<<256, 256, 0, st>>>(c->ws.partial, slices, stride, (long long)M * N, C);
        DPPO_LAUNCH_CHECK();
        if (bias_out) {
            reduce_slices_kernel<<<4, 256, 0, st>>>(c->ws.partial + (int64_t)M * N, slices, stride, N, bias_out);
            DPPO_LAUNCH_CHECK();
        }
        return 0;
    }
    if (mode == 92 || mode == 94 || mode == 95) {        // 92: as 94, bias_out receives the kernel's clock64 time line (128 int64)
        // as 96 / 97 with A given to the kernel as fp16 hi | lo planes and C written as planes (then joined back to fp32):
        // the operand path of the update's hidden layers
        const long long need = th_image_floats(N, K);
        DPPO_REQUIRE(c->shape.partial_floats >= need, "dppo_dbg_gemm: the partial buffer cannot hold the %lld-float weight image", need);
        const long long cap = (long long)c->cfg.max_rows * (c->shape.hmax() > 0 ? c->shape.hmax() : 1);
        ThPlanes pl{};
        pl.terms = terms;
        pl.a_planes = pl.c_planes = 1;
        pl.a_lo_off = ((long long)M * K + 7) & ~7LL; pl.c_lo_off = ((long long)M * N + 7) & ~7LL;
        DPPO_REQUIRE(pl.a_lo_off <= cap && pl.c_lo_off <= cap, "dppo_dbg_gemm: planes exceed the scratch buffers");
        __half* a_hi = reinterpret_cast<__half*>(c->ws.scratch[0]);
        __half* c_hi = reinterpret_cast<__half*>(c->ws.scratch[1]);
        if (int r = th_split_planes(A, (long long)M * K, a_hi, a_hi + pl.a_lo_off, st, terms)) return r;
        if (int r = th_pack_one(B, c->ws.partial, N, K, K, 0, st, terms, M, 1)) return r;
        if (mode == 95) {
            if (int r = th_gemm(1, M, N, K, c->ws.scratch[0], nullptr, K, c->ws.partial, nullptr, c->ws.scratch[1], nullptr, N, nullptr,
                                nullptr, 0, mask, 1, st, nullptr, nullptr, nullptr, &pl)) return r;
        } else {
            if (int r = th_gemm(0, M, N, K, c->ws.scratch[0], nullptr, K, c->ws.partial, nullptr, c->ws.scratch[1], nullptr, N, bias,
                                nullptr, relu, nullptr, 1, st, nullptr, nullptr,
                                mode == 92 ? reinterpret_cast<long long*>(bias_out) : nullptr, &pl)) return r;
        }
        return th_join_planes(c_hi, c_hi + pl.c_lo_off, (long long)M * N, C, st, terms);
    }
    if (mode == 96 || mode == 97 || mode == 98) {
        // persistent fp16-split product (tc_gemm_h.cu): Bt[N,K] is packed into an fp16 hi|lo tile image first (scratch from
        // the caller's workspace).  96: C = relu?(A.Bt^T + bias); 97: C = (A.Bt^T) * (mask > 0);
        // 98: as 96, and bias_out receives the kernel's clock64 time line (bias_out must hold 128 int64)
        const long long need = th_image_floats(N, K);
        float* scratch = nullptr;
        if (c->shape.partial_floats >= need) scratch = c->ws.partial;
        else if ((long long)c->ws.stage_floats >= need) scratch = c->ws.stage;
        else if ((long long)c->cfg.max_rows * (c->shape.hmax() > 0 ? c->shape.hmax() : 1) >= need) scratch = c->ws.scratch[0];
        DPPO_REQUIRE(scratch != nullptr, "dppo_dbg_gemm: the workspace has no %lld-float scratch for the weight image", need);
        if (int r = th_pack_one(B, scratch, N, K, K, 0, st, terms, M)) return r;
        ThPlanes pl{};
        pl.terms = terms;
        if (mode == 97)
            return th_gemm(1, M, N, K, A, nullptr, K, scratch, nullptr, C, nullptr, N, nullptr, nullptr, 0, mask, 1, st, nullptr, nullptr,
                           nullptr, &pl);
        return th_gemm(0, M, N, K, A, nullptr, K, scratch, nullptr, C, nullptr, N, bias, nullptr, relu, nullptr, 1, st,
                       nullptr, nullptr, mode == 98 ? reinterpret_cast<long long*>(bias_out) : nullptr, &pl);
    }
    if (mode == 48 || mode == 49) {
        // weight-image product (the tf32 path of the update): Bt[N,K] is packed into a smem-tile image first (scratch from
        // the caller's workspace)
        const long long need = tc_image_floats(N, K);
        float* scratch = nullptr;
        if (c->shape.partial_floats >= need) scratch = c->ws.partial;
        else if ((long long)c->ws.stage_floats >= need) scratch = c->ws.stage;
        DPPO_REQUIRE(scratch != nullptr, "dppo_dbg_gemm: the workspace has no %lld-float scratch for the weight image", need);
        if (int r = tc_pack_one(B, scratch, N, K, K, 0, st)) return r;
        return tc_gemm_kk_img(mode == 49, M, N, K, A, nullptr, K, scratch, nullptr, C, nullptr, N, bias, nullptr, relu,
                              mode == 49 ? mask : nullptr, 1, st);
    }
    if (mode >= 16) {
        // tensor-core products: terms = 3 (modes 16..18) or 1 (modes 32..34).
        // 16/32: C = relu?(A[M,K].Bt[N,K]^T + bias); 17/33: same product, C *= (mask>0); 18/34: C = A[K,M]^T.B[K,N]
        const int terms = mode >= 32 ? 1 : 3, sub = mode & 15;
        if (sub == 0) return tc_gemm_kk(terms, 0, M, N, K, A, nullptr, K, B, nullptr, K, C, nullptr, N, bias, nullptr, relu,
                                        nullptr, 1, st);
        if (sub == 1) return tc_gemm_kk(terms, 1, M, N, K, A, nullptr, K, B, nullptr, K, C, nullptr, N, nullptr, nullptr, 0,
                                        mask, 1, st);
        if (sub == 3) {      // TMA-fed dW product with fused bias column sums
            const int64_t stride = ((int64_t)M * N + N + 3) & ~3LL;
            int slices = (int)(c->shape.partial_floats / stride);
            const int want = (K + 511) / 512;
            if (slices > want) slices = want;
            DPPO_REQUIRE(slices >= 1, "dppo_dbg_gemm: M*N+N exceeds the partial buffer");
            DPPO_REQUIRE(tc_tn_tma_supported(A, M, B, N), "dppo_dbg_gemm: TMA path needs 16-byte aligned rows");
            if (int r = tc_gemm_tn_tma(M, N, K, A, M, B, N, c->ws.partial, N, c->ws.partial + (int64_t)M * N, slices, stride,
                                       0, st)) return r;
            ProfScope ps(KK_MISC, st);
            reduce_slices_kernel<<<256, 256, 0, st>>>(c->ws.partial, slices, stride, (long long)M * N, C);
            DPPO_LAUNCH_CHECK();
            if (bias_out) {
                reduce_slices_kernel<<<4, 256, 0, st>>>(c->ws.partial + (int64_t)M * N, slices, stride, N, bias_out);
                DPPO_LAUNCH_CHECK();
            }
            return 0;
        }
        if (sub == 2) {
            const int64_t stride = ((int64_t)M * N + N + 3) & ~3LL;
            int slices = (int)(c->shape.partial_floats / stride);
            if (slices > 5) slices = 5;
            DPPO_REQUIRE(slices >= 1, "dppo_dbg_gemm: M*N+N exceeds the partial buffer");
            if (int r = tc_gemm_tn_splitk(terms, M, N, K, A, M, B, N, c->ws.partial, N, slices, stride, 0, st)) return r;
            if (bias_out)
                if (int r = tc_colsum(B, K, N, N, slices, c->ws.partial + (int64_t)M * N, stride, 0, st)) return r;
            ProfScope ps(KK_MISC, st);
            reduce_slices_kernel<<<256, 256, 0, st>>>(c->ws.partial, slices, stride, (long long)M * N, C);
            DPPO_LAUNCH_CHECK();
            if (bias_out) {
                reduce_slices_kernel<<<4, 256, 0, st>>>(c->ws.partial + (int64_t)M * N, slices, stride, N, bias_out);
                DPPO_LAUNCH_CHECK();
            }
            return 0;
        }
        DPPO_REQUIRE(false, "dppo_dbg_gemm: bad tensor-core mode %d", mode);
    }
    if (mode == 0) return gemm_nn(M, N, K, A, K, B, N, C, N, e, st);
    if (mode == 1) return gemm_nt(M, N, K, A, K, B, K, C, N, e, st);
    if (mode == 2) {
        // C[M,N] = A[K,M]^T . B[K,N]; partial slices live in the ctx workspace
        const int64_t need = (int64_t)M * N + N;
        const int64_t stride = (need + 3) & ~3LL;
        int slices = (int)(c->shape.partial_floats / stride);
        if (slices > 37) slices = 37;
        DPPO_REQUIRE(slices >= 1, "dppo_dbg_gemm: M*N+N=%lld exceeds the partial buffer (%lld floats)",
                     (long long)need, (long long)c->shape.partial_floats);
        e.slice_stride = stride; e.slices = slices; e.bias_out = bias_out ? c->ws.partial + (int64_t)M * N : nullptr;
        if (int r = gemm_tn_splitk(M, N, K, A, M, B, N, c->ws.partial, N, e, st)) return r;
        ProfScope ps(KK_MISC, st);
        reduce_slices_kernel<<<256, 256, 0, st>>>(c->ws.partial, slices, stride, (long long)M * N, C);
        DPPO_LAUNCH_CHECK();
        if (bias_out) {
            reduce_slices_kernel<<<4, 256, 0, st>>>(c->ws.partial + (int64_t)M * N, slices, stride, N, bias_out);
            DPPO_LAUNCH_CHECK();
        }
        return 0;
    }
    DPPO_REQUIRE(false, "dppo_dbg_gemm: bad mode %d", mode);
}

}  // extern "C"
