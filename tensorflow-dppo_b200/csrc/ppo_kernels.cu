// K4: DiagGaussianPd + PPO clipped loss forward/backward on the head outputs, plus the small
// elementwise kernels around it (head packing, split-K finalize, Adam, sampling).
//
// Reference math: /root/reference/Others/distributions.py:184-208 (DiagGaussianPd),
// /root/reference/PPO.py:19,29-40 (losses), PPO.py:46 (gradient; closed form per SURVEY
// Appendix B, with TensorFlow's tie rules for minimum / maximum / clip_by_value).
#include "common.cuh"

namespace dppo {

// ---------------------------------------------------------------------------------------------
// head packing: public [Wv(H) | bv | Wp(H,2A) | bp(2A)]  ->  internal [Wh(H,P) | bh(P)]
// (Model.py:13-14 creates two dense layers on the same input; one GEMM serves both.)
// ---------------------------------------------------------------------------------------------
__global__ void pack_heads_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                  float* __restrict__ h0, float* __restrict__ h1, int H, int W,
                                  long long head_off) {
    const int P = 1 + W;                     // W = policy-head width (2A or ncat)
    const int total = H * P + P;
    const float* src = (blockIdx.y == 0 ? p0 : p1);
    float* dst = (blockIdx.y == 0 ? h0 : h1);
    if (src == nullptr) return;
    src += head_off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float v;
        if (i < H * P) {
            const int k = i / P, c = i % P;
            v = (c == 0) ? src[k] : src[H + 1 + k * W + (c - 1)];
        } else {
            const int c = i - H * P;
            v = (c == 0) ? src[H] : src[H + 1 + W * H + (c - 1)];
        }
        dst[i] = v;
    }
}

int launch_pack_heads(const NetShape& s, const float* params_pi, const float* params_old,
                      float* head_pi, float* head_old, cudaStream_t st) {
    const int H = s.in_dim(s.L);
    const int total = H * s.P + s.P;
    dim3 grid((total + 255) / 256, 2);
    ProfScope ps(KK_MISC, st);
    pack_heads_kernel<<<grid, 256, 0, st>>>(params_pi, params_old, head_pi, head_old, H, s.pdw,
                                            (long long)s.w_off[s.L]);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// loss forward + d(total)/d(head outputs)
// ---------------------------------------------------------------------------------------------
// One CTA = kLossTileRows rows.  The [rows,P] head tiles and the [rows,A] action tile are
// staged through shared memory with fully coalesced float loads (P = 1+2A is odd, so the
// row-strided per-thread reads that follow are bank-conflict free), one thread then owns one
// row, and d_out goes back through the same smem tile to a coalesced store.  The seven batch sums
// (surrogate, value, entropy, kl, clip counters) are reduced with warp shuffles, then across the
// 4 warps, and land as fp64 partials per tile; finalize_kernel adds the tiles in a fixed order.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kLossTileRows)
ppo_loss_kernel(const float* __restrict__ out_pi, const float* __restrict__ out_old,
                const float* __restrict__ act, const float* __restrict__ adv,
                const float* __restrict__ ret, int rows, int A, int Pld, int tile0, float inv_B, float clip,
                float entcoeff, float vcoeff, float* __restrict__ d_out,
                double* __restrict__ loss_part) {
    extern __shared__ float smem[];
    const int P = 1 + 2 * A;
    float* s_pi = smem;                              // [R][P]   (becomes d_out)
    float* s_old = s_pi + kLossTileRows * P;         // [R][P]
    float* s_act = s_old + kLossTileRows * P;        // [R][A]
    __shared__ double s_red[kLossTileRows / 32][kLossSlots];

    const int row0 = blockIdx.x * kLossTileRows;
    const int nrows = min(kLossTileRows, rows - row0);
    const int tid = threadIdx.x;

    {   // coalesced 16-B tile loads: global rows are Pld wide (Pld % 4 == 0, 16-B aligned), smem rows P wide
        const float4* gp = reinterpret_cast<const float4*>(out_pi + (size_t)row0 * Pld);
        const float4* go = reinterpret_cast<const float4*>(out_old + (size_t)row0 * Pld);
        const int pld4 = Pld >> 2;
        for (int i = tid; i < nrows * pld4; i += kLossTileRows) {
            const int r = i / pld4, c0 = (i - r * pld4) * 4;
            const float4 a4 = __ldg(gp + i), b4 = __ldg(go + i);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (c0 + e < P) { s_pi[r * P + c0 + e] = av[e]; s_old[r * P + c0 + e] = bv[e]; }
        }
        const float* ga = act + (size_t)row0 * A;
        const int na = nrows * A;
        if (((na & 3) == 0) && ((reinterpret_cast<uintptr_t>(ga) & 15) == 0)) {
            const float4* ga4 = reinterpret_cast<const float4*>(ga);
            for (int i = tid; i < (na >> 2); i += kLossTileRows) reinterpret_cast<float4*>(s_act)[i] = __ldg(ga4 + i);
        } else {
            for (int i = tid; i < na; i += kLossTileRows) s_act[i] = __ldg(ga + i);
        }
    }
    __syncthreads();

    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    if (tid < nrows) {
        float* o = s_pi + tid * P;
        const float* q = s_old + tid * P;
        const float* a = s_act + tid * A;
        const float advb = __ldg(adv + row0 + tid);
        const float R = __ldg(ret + row0 + tid);

        // log ratio = logp_pi(a) - logp_old(a), accumulated per component so that the shared
        // 0.5*log(2pi)*A term cancels exactly (distributions.py:195-198, PPO.py:31)
        float logratio = 0.f, sum_ls = 0.f, kl = 0.f;
        for (int k = 0; k < A; ++k) {
            const float mu = o[1 + k], ls = o[1 + A + k];
            const float muo = q[1 + k], lso = q[1 + A + k];
            const float z = (a[k] - mu) * expf(-ls);
            const float zo = (a[k] - muo) * expf(-lso);
            logratio += 0.5f * (zo * zo - z * z) + (lso - ls);
            sum_ls += ls;
            // KL(old || new), distributions.py:199-201 (diagnostic only)
            const float dm = muo - mu;
            kl += (ls - lso) + (expf(2.f * lso) + dm * dm) * 0.5f * expf(-2.f * ls) - 0.5f;
        }
        const float ratio = expf(logratio);
        const float ent = sum_ls + (float)A * 1.4189385332046727f;     // 0.5*log(2*pi*e), :202-203
        const float surr1 = ratio * advb;                                // PPO.py:32
        const float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
        const float surr2 = rc * advb;                                   // PPO.py:33
        const bool first = (surr1 <= surr2);                             // tf.minimum -> x iff x<=y
        const float g_logp = first ? (-advb * inv_B) * ratio : 0.f;

        const float v = o[0], vo = q[0];
        const float dv = v - vo;
        const float vclip = vo + fminf(fmaxf(dv, -clip), clip);         // PPO.py:37
        const float e1 = (v - R) * (v - R);                              // PPO.py:36
        const float e2 = (vclip - R) * (vclip - R);                      // PPO.py:38
        const bool vfirst = (e1 >= e2);                                  // tf.maximum -> x iff x>=y
        const bool inclip = (dv >= -clip) && (dv <= clip);
        const float g_v = vcoeff * inv_B * (vfirst ? 2.f * (v - R) : (inclip ? 2.f * (vclip - R) : 0.f));

        acc[0] = -(double)fminf(surr1, surr2);                           // PPO.py:34
        acc[1] = (double)fmaxf(e1, e2);                                  // PPO.py:39
        acc[2] = (double)ent;                                            // PPO.py:29-30
        acc[3] = (double)kl;
        acc[4] = first ? 0.0 : 1.0;
        acc[5] = (!vfirst && !inclip) ? 1.0 : 0.0;                       // clipped value branch selected, clip active
        acc[6] = 1.0;

        // d total / d [v | mean | logstd]  
        const float ent_g = entcoeff * inv_B;
        for (int k = 0; k < A; ++k) {
            const float mu = o[1 + k], ls = o[1 + A + k];
            const float inv_std = expf(-ls);
            const float z = (a[k] - mu) * inv_std;
            o[1 + k] = g_logp * z * inv_std;
            o[1 + A + k] = g_logp * (z * z - 1.f) - ent_g;
        }
        o[0] = g_v;
    }
    __syncthreads();
    {   // coalesced 16-B d_out store (pad columns get zeros)
        float4* gd = reinterpret_cast<float4*>(d_out + (size_t)row0 * Pld);
        const int pld4 = Pld >> 2;
        for (int i = tid; i < nrows * pld4; i += kLossTileRows) {
            const int r = i / pld4, c0 = (i - r * pld4) * 4;
            float o4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o4[e] = (c0 + e < P) ? s_pi[r * P + c0 + e] : 0.f;
            gd[i] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
    }
    // block reduction of the batch sums
    const int warp = tid / 32, lane = tid % 32;
#pragma unroll
    for (int s = 0; s < 7; ++s) {
        const double w = warp_sum(acc[s]);
        if (lane == 0) s_red[warp][s] = w;
    }
    __syncthreads();
    if (tid < 7) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kLossTileRows / 32; ++w) t += s_red[w][tid];
        loss_part[(size_t)(tile0 + blockIdx.x) * kLossSlots + tid] = t;
    }
}

// Same tile scheme for CategoricalPd heads (distributions.py:124-159): the row's [v | logits(ncat)] for pi and oldpi,
// one int32 class index per row.  logp = log_softmax(logits)[a] (:131-138; an out-of-range a gives a zero one_hot row,
// hence logp = 0 and no policy gradient, as TF's op does), entropy :148-153, KL(old || new) :139-147 as a diagnostic.
//   d total / d logit_k = g_logp * ([k == a] - p_k) + (ENTCOEFF / B) * p_k * (log p_k + H)
__global__ void __launch_bounds__(kLossTileRows)
ppo_loss_cat_kernel(const float* __restrict__ out_pi, const float* __restrict__ out_old,
                    const int* __restrict__ act, const float* __restrict__ adv,
                    const float* __restrict__ ret, int rows, int ncat, int Pld, int tile0, float inv_B, float clip,
                    float entcoeff, float vcoeff, float* __restrict__ d_out,
                    double* __restrict__ loss_part) {
    extern __shared__ float smem[];
    const int P = 1 + ncat;
    const int Ps = P | 1;                            // odd smem row stride: the row-strided per-thread reads stay conflict free
    float* s_pi = smem;                              // [R][Ps]  (becomes d_out)
    float* s_old = s_pi + kLossTileRows * Ps;        // [R][Ps]
    __shared__ double s_red[kLossTileRows / 32][kLossSlots];

    const int row0 = blockIdx.x * kLossTileRows;
    const int nrows = min(kLossTileRows, rows - row0);
    const int tid = threadIdx.x;
    {
        const float* gp = out_pi + (size_t)row0 * Pld;
        const float* go = out_old + (size_t)row0 * Pld;
        for (int i = tid; i < nrows * Pld; i += kLossTileRows) {
            const int r = i / Pld, cc = i - r * Pld;
            if (cc < P) { s_pi[r * Ps + cc] = __ldg(gp + i); s_old[r * Ps + cc] = __ldg(go + i); }
        }
    }
    __syncthreads();

    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    if (tid < nrows) {
        float* o = s_pi + tid * Ps;
        const float* q = s_old + tid * Ps;
        const int a = __ldg(act + row0 + tid);
        const bool valid = (a >= 0) && (a < ncat);
        const float advb = __ldg(adv + row0 + tid);
        const float R = __ldg(ret + row0 + tid);

        float m = o[1], mo = q[1];
        for (int k = 1; k < ncat; ++k) { m = fmaxf(m, o[1 + k]); mo = fmaxf(mo, q[1 + k]); }
        float z = 0.f, zo = 0.f;
        for (int k = 0; k < ncat; ++k) { z += expf(o[1 + k] - m); zo += expf(q[1 + k] - mo); }
        const float lz = logf(z), lzo = logf(zo), inv_z = 1.f / z, inv_zo = 1.f / zo;
        float ent = 0.f, kl = 0.f;
        for (int k = 0; k < ncat; ++k) {
            const float a0 = o[1 + k] - m, a1 = q[1 + k] - mo;
            ent += expf(a0) * inv_z * (lz - a0);                          // :148-153
            kl += expf(a1) * inv_zo * (a1 - lzo - a0 + lz);               // :139-147, self = oldpi
        }
        const float logp = valid ? (o[1 + a] - m) - lz : 0.f;
        const float logpo = valid ? (q[1 + a] - mo) - lzo : 0.f;
        const float ratio = expf(logp - logpo);                          // PPO.py:31
        const float surr1 = ratio * advb;
        const float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
        const float surr2 = rc * advb;
        const bool first = (surr1 <= surr2);
        const float g_logp = (first && valid) ? (-advb * inv_B) * ratio : 0.f;

        const float v = o[0], vo = q[0];
        const float dv = v - vo;
        const float vclip = vo + fminf(fmaxf(dv, -clip), clip);
        const float e1 = (v - R) * (v - R);
        const float e2 = (vclip - R) * (vclip - R);
        const bool vfirst = (e1 >= e2);
        const bool inclip = (dv >= -clip) && (dv <= clip);
        const float g_v = vcoeff * inv_B * (vfirst ? 2.f * (v - R) : (inclip ? 2.f * (vclip - R) : 0.f));

        acc[0] = -(double)fminf(surr1, surr2);
        acc[1] = (double)fmaxf(e1, e2);
        acc[2] = (double)ent;
        acc[3] = (double)kl;
        acc[4] = first ? 0.0 : 1.0;
        acc[5] = (!vfirst && !inclip) ? 1.0 : 0.0;
        acc[6] = 1.0;

        const float ent_g = entcoeff * inv_B;
        for (int k = 0; k < ncat; ++k) {
            const float a0 = o[1 + k] - m;
            const float pk = expf(a0) * inv_z;
            o[1 + k] = g_logp * ((k == a ? 1.f : 0.f) - pk) + ent_g * pk * ((a0 - lz) + ent);
        }
        o[0] = g_v;
    }
    __syncthreads();
    {
        float* gd = d_out + (size_t)row0 * Pld;
        for (int i = tid; i < nrows * Pld; i += kLossTileRows) {
            const int r = i / Pld, cc = i - r * Pld;
            gd[i] = cc < P ? s_pi[r * Ps + cc] : 0.f;
        }
    }
    const int warp = tid / 32, lane = tid % 32;
#pragma unroll
    for (int s = 0; s < 7; ++s) {
        const double w = warp_sum(acc[s]);
        if (lane == 0) s_red[warp][s] = w;
    }
    __syncthreads();
    if (tid < 7) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kLossTileRows / 32; ++w) t += s_red[w][tid];
        loss_part[(size_t)(tile0 + blockIdx.x) * kLossSlots + tid] = t;
    }
}

int launch_loss(const NetShape& s, const float* out_pi, const float* out_old, const float* act,
                const float* adv, const float* ret, int rows, int tile0, float inv_B, float clip,
                float entcoeff, float vcoeff, float* d_out, double* loss_part, cudaStream_t st) {
    if (rows == 0) return 0;
    const int grid = (rows + kLossTileRows - 1) / kLossTileRows;
    if (s.pd == DPPO_PD_CATEGORICAL) {
        const size_t smem = (size_t)kLossTileRows * 2 * (s.P | 1) * sizeof(float);
        DPPO_REQUIRE(smem <= 200 * 1024, "loss kernel: ncat=%d needs %zu B of shared memory", s.act, smem);
        static PerDeviceMax configured_cat;
        if (smem > 48 * 1024 && configured_cat.need(smem))
            DPPO_CHECK_CUDA(cudaFuncSetAttribute(ppo_loss_cat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ProfScope ps(KK_LOSS, st);
        ppo_loss_cat_kernel<<<grid, kLossTileRows, smem, st>>>(out_pi, out_old, reinterpret_cast<const int*>(act), adv, ret,
                                                               rows, s.act, s.Pld, tile0, inv_B, clip, entcoeff, vcoeff,
                                                               d_out, loss_part);
        DPPO_LAUNCH_CHECK();
        return 0;
    }
    const size_t smem = (size_t)kLossTileRows * (2 * s.P + s.act) * sizeof(float);
    DPPO_REQUIRE(smem <= 200 * 1024, "loss kernel: act_dim=%d needs %zu B of shared memory", s.act, smem);
    static PerDeviceMax configured;
    if (smem > 48 * 1024 && configured.need(smem))
        DPPO_CHECK_CUDA(cudaFuncSetAttribute(ppo_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(KK_LOSS, st);
    ppo_loss_kernel<<<grid, kLossTileRows, smem, st>>>(out_pi, out_old, act, adv, ret, rows, s.act, s.Pld, tile0,
                                                       inv_B, clip, entcoeff, vcoeff, d_out, loss_part);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Literal value loss: PPO.py:36-39 as committed broadcasts v [B,1,1] against etr [B,1] to
// [B,B,1], i.e. every value of a tower (= one rollout) is paired with every return of that rollout.  Runs AFTER the
// loss kernel on the same 128-row tiles: thread = row i, loop over the T returns of its rollout (rows are time-major,
// row b = t*groups + n, so a warp reads consecutive rollouts' R_j: coalesced; groups = 1 broadcasts), fp64 sums, then
// replaces d_out[:,0] and the value-loss / clipped-fraction slots of the tile's loss partials.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLossTileRows)
value_pairwise_kernel(const float* __restrict__ out_pi, const float* __restrict__ out_old, const float* __restrict__ ret_all,
                      int rows, int row_off, int T, int groups, int Pld, int tile0, float inv_B, float clip, float vcoeff,
                      float* __restrict__ d_out, double* __restrict__ loss_part) {
    __shared__ double s_red[kLossTileRows / 32][2];
    const int tid = threadIdx.x;
    const int r = blockIdx.x * kLossTileRows + tid;
    double e_sum = 0.0, g_sum = 0.0, c_sum = 0.0;
    if (r < rows) {
        const int b = row_off + r;
        const int g = b % groups;
        const float v = out_pi[(size_t)r * Pld], vo = out_old[(size_t)r * Pld];
        const float dv = v - vo;
        const float vclip = vo + fminf(fmaxf(dv, -clip), clip);
        const bool inclip = (dv >= -clip) && (dv <= clip);
        const float* R = ret_all + g;
#pragma unroll 4
        for (int j = 0; j < T; ++j) {
            const float Rj = __ldg(R + (size_t)j * groups);
            const float a1 = v - Rj, a2 = vclip - Rj;
            const float e1 = a1 * a1, e2 = a2 * a2;
            const bool first = (e1 >= e2);                                   // tf.maximum -> x iff x >= y
            e_sum += (double)fmaxf(e1, e2);
            g_sum += (double)(first ? 2.f * a1 : (inclip ? 2.f * a2 : 0.f));
            c_sum += (!first && !inclip) ? 1.0 : 0.0;
        }
        const double invT = 1.0 / (double)T;
        e_sum *= invT; g_sum *= invT; c_sum *= invT;
        d_out[(size_t)r * Pld] = (float)((double)vcoeff * (double)inv_B * g_sum);
    }
    const int warp = tid / 32, lane = tid % 32;
    const double we = warp_sum(e_sum), wc = warp_sum(c_sum);
    if (lane == 0) { s_red[warp][0] = we; s_red[warp][1] = wc; }
    __syncthreads();
    if (tid < 2) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kLossTileRows / 32; ++w) t += s_red[w][tid];
        loss_part[(size_t)(tile0 + blockIdx.x) * kLossSlots + (tid == 0 ? 1 : 5)] = t;
    }
}

int launch_value_pairwise(const NetShape& s, const float* out_pi, const float* out_old, const float* ret_all, int rows,
                          int row_off, int T, int groups, int tile0, float inv_B, float clip, float vcoeff, float* d_out,
                          double* loss_part, cudaStream_t st) {
    if (rows == 0) return 0;
    const int grid = (rows + kLossTileRows - 1) / kLossTileRows;
    ProfScope ps(KK_LOSS, st);
    value_pairwise_kernel<<<grid, kLossTileRows, 0, st>>>(out_pi, out_old, ret_all, rows, row_off, T, groups, s.Pld, tile0,
                                                          inv_B, clip, vcoeff, d_out, loss_part);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// finalize: deterministic split-K reduction of dW/db partials (internal layout) into the public
// flat gradient (PPO.py:46 order), + the loss scalars at the tail.
// ---------------------------------------------------------------------------------------------
struct FinalizePlan {
    int n_regions;
    long long w_off[kMaxLayers + 1];      // internal offset of region l (= [W_l | b_l], contiguous)
    long long region[kMaxLayers + 1];     // slice stride of region l
    long long part_off[kMaxLayers + 1];   // where region l's slices start in the partial buffer
    int slices[kMaxLayers + 1];
    int few;                      // every region has <= 20 slices: thread-per-parameter summation (finalize_kernel)
};

// one TF-1.x Adam step of one parameter (PPO.py:20,53), shared by adam_kernel and the fused finalize + Adam tail
__device__ __forceinline__ void adam_update(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, long long i, float gi,
                                            float lr_t, float b1, float b2, float eps) {
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}
struct AdamTail { float* p; float* m; float* v; float lr_t, b1, b2, eps; };     // p == nullptr: finalize only

__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ partial, FinalizePlan plan, long long n_params,
                                long long head_off, int H, int W, const double* __restrict__ loss_part, int n_tiles,
                                float entcoeff, float vcoeff, float* __restrict__ grad_out,
                                float* __restrict__ losses_out, float* __restrict__ grad_out_alt,
                                const StepParams* __restrict__ sp, const int* __restrict__ counter, float grad_scale, AdamTail adam) {
    const int P = 1 + W;                     // W = policy-head width (2A or ncat)
    if (sp && grad_out_alt && sp[*counter].peer_sel) grad_out = grad_out_alt;      // graph replay: the exchange buffer half of this step
    if (adam.p && sp) adam.lr_t = sp[*counter].lr_t;
    // A block owns 32 consecutive parameters.  Warp w streams slices w, w+8, ... of them: every load instruction
    // is one full 128-byte line (32 lanes x consecutive parameters of ONE slice), 8 lines in flight per warp; the 8
    // warps' sums are then combined in a fixed order through shared memory -> deterministic and DRAM-friendly.
    __shared__ float red[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // publish parameter i (head rows are stored transposed in the partials: see pack_heads_kernel) + the optional Adam step
    auto publish = [&](long long i, float t) {
        long long pub = i;
        if (i >= head_off) {
            const long long r = i - head_off;
            if (r < (long long)H * P) {
                const int kk = (int)(r / P), c = (int)(r % P);
                pub = head_off + (c == 0 ? kk : (long long)H + 1 + (long long)kk * W + (c - 1));
            } else {
                const int c = (int)(r - (long long)H * P);
                pub = head_off + (c == 0 ? (long long)H : (long long)H + 1 + (long long)W * H + (c - 1));
            }
        }
        const float gi = t * grad_scale;     // 1 / the power of two the loss kernel put on d_out (fp16-split mode), else 1
        grad_out[pub] = gi;
        // single-rank updates: the optimiser step of this parameter right here (no second pass over the gradient)
        if (adam.p) adam_update(adam.p, adam.m, adam.v, pub, gi, adam.lr_t, adam.b1, adam.b2, adam.eps);
    };
    if (plan.few) {
        // few slices (config 4: 9 - 18 of a 3.6 M-parameter net): thread = parameter, the slices summed in order -- every load
        // instruction of a warp is one 128-byte line and 256 of them are in flight per block (the layout below would leave
        // seven of its eight warps with a single line each)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_params; i += (long long)gridDim.x * blockDim.x) {
            int l = 0;
            while (l + 1 < plan.n_regions && i >= plan.w_off[l + 1]) ++l;
            const float* src = partial + plan.part_off[l] + (i - plan.w_off[l]);
            const long long rg = plan.region[l];
            const int ns = plan.slices[l];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int k = 0;
            for (; k + 3 < ns; k += 4) {
                a0 += __ldg(src + (size_t)(k + 0) * rg); a1 += __ldg(src + (size_t)(k + 1) * rg);
                a2 += __ldg(src + (size_t)(k + 2) * rg); a3 += __ldg(src + (size_t)(k + 3) * rg);
            }
            for (; k < ns; ++k) a0 += __ldg(src + (size_t)k * rg);
            publish(i, (a0 + a1) + (a2 + a3));
        }
    } else
    for (long long i0 = (long long)blockIdx.x * 32; i0 < n_params; i0 += (long long)gridDim.x * 32) {
        const long long i = i0 + lane;
        float s = 0.f;
        if (i < n_params) {
            int l = 0;
            while (l + 1 < plan.n_regions && i >= plan.w_off[l + 1]) ++l;
            const float* src = partial + plan.part_off[l] + (i - plan.w_off[l]);
            const long long rg = plan.region[l];
            const int ns = plan.slices[l];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
            int k = warp;
            for (; k + 56 < ns; k += 64) {           // 8 independent 128-B lines in flight per warp
                a0 += __ldg(src + (size_t)(k + 0) * rg);  a1 += __ldg(src + (size_t)(k + 8) * rg);
                a2 += __ldg(src + (size_t)(k + 16) * rg); a3 += __ldg(src + (size_t)(k + 24) * rg);
                a4 += __ldg(src + (size_t)(k + 32) * rg); a5 += __ldg(src + (size_t)(k + 40) * rg);
                a6 += __ldg(src + (size_t)(k + 48) * rg); a7 += __ldg(src + (size_t)(k + 56) * rg);
            }
            for (; k + 24 < ns; k += 32) {
                a0 += __ldg(src + (size_t)(k + 0) * rg);  a1 += __ldg(src + (size_t)(k + 8) * rg);
                a2 += __ldg(src + (size_t)(k + 16) * rg); a3 += __ldg(src + (size_t)(k + 24) * rg);
            }
            for (; k < ns; k += 8) a0 += __ldg(src + (size_t)k * rg);
            s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
        }
        red[warp][lane] = s;
        __syncthreads();
        if (warp == 0 && i < n_params) {
            const float t = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) +
                            ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
            publish(i, t);
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        double acc[7];
#pragma unroll
        for (int s = 0; s < 7; ++s) acc[s] = 0.0;
        for (int t = threadIdx.x; t < n_tiles; t += 32)
#pragma unroll
            for (int s = 0; s < 7; ++s) acc[s] += loss_part[(size_t)t * kLossSlots + s];
#pragma unroll
        for (int s = 0; s < 7; ++s) acc[s] = warp_sum(acc[s]);
        if (threadIdx.x == 0) {
            const double n = acc[6] > 0 ? acc[6] : 1.0;
            float o[kLossSlots];
            o[0] = (float)(acc[0] / n);                       // policyLoss   PPO.py:34
            o[1] = (float)((double)vcoeff * acc[1] / n);      // valueLoss    PPO.py:39
            o[2] = (float)(-(double)entcoeff * acc[2] / n);   // entropyLoss  PPO.py:35
            o[3] = o[0] + o[2] + o[1];                        // total_loss   PPO.py:40
            o[4] = (float)(acc[3] / n);                       // mean KL(old||new)
            o[5] = (float)(acc[4] / n);                       // fraction of rows on the clipped surrogate
            o[6] = (float)(acc[5] / n);                       // fraction of rows on the clipped value branch
            o[7] = (float)acc[6];                             // rows
#pragma unroll
            for (int s = 0; s < kLossSlots; ++s) {
                grad_out[n_params + s] = o[s];
                if (losses_out) losses_out[s] = o[s];
            }
        }
    }
}

double adam_lr_t(float lr, float b1, float b2, int step);

int launch_finalize(const NetShape& s, const int* active, const float* partial, const double* loss_part, int n_loss_tiles,
                    float entcoeff, float vcoeff, float* grad_out, float* losses_out, cudaStream_t st, float* grad_out_alt,
                    const StepParams* sp, const int* counter, float grad_scale, const AdamFused* fused) {
    FinalizePlan plan;
    plan.n_regions = s.L + 1;
    for (int l = 0; l <= s.L; ++l) {
        plan.w_off[l] = s.w_off[l]; plan.region[l] = s.region[l];
        plan.part_off[l] = s.part_off[l]; plan.slices[l] = active ? active[l] : s.slices[l];
    }
    plan.few = 1;
    for (int l = 0; l <= s.L; ++l)
        if (plan.slices[l] > 20) plan.few = 0;
    if (s.n_params < (1 << 18)) plan.few = 0;            // (small nets: either way is a few microseconds)
    const int H = s.in_dim(s.L);
    int grid = (int)((s.n_params + 31) / 32);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    AdamTail adam{};
    if (fused && fused->params) {
        DPPO_REQUIRE(fused->step >= 1, "dppo_adam_apply: step=%d must be >= 1 (t after increment)", fused->step);
        adam.p = fused->params; adam.m = fused->m; adam.v = fused->v;
        adam.lr_t = (float)adam_lr_t(fused->lr, fused->b1, fused->b2, fused->step);
        adam.b1 = fused->b1; adam.b2 = fused->b2; adam.eps = fused->eps;
    }
    ProfScope ps(KK_FINALIZE, st);
    finalize_kernel<<<grid, 256, 0, st>>>(partial, plan, s.n_params, (long long)s.w_off[s.L], H, s.pdw, loss_part,
                                          n_loss_tiles, entcoeff, vcoeff, grad_out, losses_out, grad_out_alt, sp, counter, grad_scale, adam);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// out[rows,P] -> v[rows], pdflat[rows,2A]
// ---------------------------------------------------------------------------------------------
__global__ void split_heads_kernel(const float* __restrict__ out, long long total, int P, int Pld,
                                   float* __restrict__ v_out, float* __restrict__ pd_out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / Pld; const int c = (int)(i - r * Pld);
        if (c >= P) continue;
        const float x = __ldg(out + i);
        if (c == 0) { if (v_out) v_out[r] = x; }
        else if (pd_out) pd_out[r * (P - 1) + (c - 1)] = x;
    }
}

int launch_split_heads(const NetShape& s, const float* out, int rows, float* v_out, float* pd_out,
                       cudaStream_t st) {
    if (rows == 0) return 0;
    const long long total = (long long)rows * s.Pld;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    ProfScope ps(KK_MISC, st);
    split_heads_kernel<<<grid, 256, 0, st>>>(out, total, s.P, s.Pld, v_out, pd_out);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K7: tf.train.AdamOptimizer apply (PPO.py:20,53), TF-1.x form:
//   lr_t = lr*sqrt(1-b2^t)/(1-b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   theta -= lr_t * m / (sqrt(v) + eps)
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                            const StepParams* __restrict__ sp, const int* __restrict__ counter) {
    if (sp) lr_t = sp[*counter].lr_t;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        adam_update(p, m, v, i, g[i], lr_t, b1, b2, eps);
}

double adam_lr_t(float lr, float b1, float b2, int step) {
    return (double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step));
}

__global__ void step_advance_kernel(const float* __restrict__ grad_tail, float* __restrict__ losses_out,
                                    const StepParams* __restrict__ sp, int* __restrict__ counter) {
    const int i = *counter;
    if (losses_out && threadIdx.x < kLossSlots) losses_out[(size_t)sp[i].loss_slot * kLossSlots + threadIdx.x] = grad_tail[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) *counter = i + 1;
}

int launch_step_advance(const float* grad_tail, float* losses_out, const StepParams* sp, int* counter, cudaStream_t st) {
    ProfScope ps(KK_MISC, st);
    step_advance_kernel<<<1, 32, 0, st>>>(grad_tail, losses_out, sp, counter);
    DPPO_LAUNCH_CHECK();
    return 0;
}

int launch_adam(float* params, const float* grad, float* m, float* v, int64_t n, int step, float lr,
                float b1, float b2, float eps, cudaStream_t st, const StepParams* sp, const int* counter) {
    DPPO_REQUIRE(step >= 1, "dppo_adam_apply: step=%d must be >= 1 (t after increment)", step);
    if (n == 0) return 0;
    const double lr_t = adam_lr_t(lr, b1, b2, step);
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope ps(KK_ADAM, st);
    adam_kernel<<<grid, 256, 0, st>>>(params, grad, m, v, (long long)n, (float)lr_t, b1, b2, eps, sp, counter);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Minibatch scheduler: row gather through the keyed permutation of perm.cuh.
// Row j of the minibatch is source row perm(first + j); every warp moves one row (obs as float4 when obs % 4 == 0).
// HBM-bound: 2 * 4 * (obs + A + 2) bytes per row.
// ---------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(PermKey key, int first, int rows, int obs_dim, int act_dim,
                                   const float* __restrict__ obs, const float* __restrict__ act,
                                   const float* __restrict__ adv, const float* __restrict__ ret,
                                   float* __restrict__ obs_o, float* __restrict__ act_o, float* __restrict__ adv_o,
                                   float* __restrict__ ret_o, const float* __restrict__ old_src, float* __restrict__ old_dst,
                                   int pld, const StepParams* __restrict__ sp, const int* __restrict__ counter) {
    if (sp) { const StepParams& q = sp[*counter]; key = q.key; first = q.first; rows = q.rows; }
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < rows; j += warps) {
        const size_t src = perm_index(key, (uint32_t)(first + j));
        const float* so = obs + src * obs_dim;
        float* dst = obs_o + (size_t)j * obs_dim;
        if ((obs_dim & 3) == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(so);
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int c = lane; c < (obs_dim >> 2); c += 32) d4[c] = __ldg(s4 + c);
        } else {
            for (int c = lane; c < obs_dim; c += 32) dst[c] = __ldg(so + c);
        }
        for (int c = lane; c < act_dim; c += 32) act_o[(size_t)j * act_dim + c] = __ldg(act + src * act_dim + c);
        if (lane == 0) { adv_o[j] = __ldg(adv + src); ret_o[j] = __ldg(ret + src); }
        if (old_src) {                              // cached oldpi head outputs [.., pld] (pld % 4 == 0, 16-B aligned rows)
            const float4* s4 = reinterpret_cast<const float4*>(old_src + src * pld);
            float4* d4 = reinterpret_cast<float4*>(old_dst + (size_t)j * pld);
            for (int c = lane; c < (pld >> 2); c += 32) d4[c] = __ldg(s4 + c);
        }
    }
}

int launch_gather_rows(const PermKey& key, int first, int rows, int obs_dim, int act_dim, const float* obs,
                       const float* act, const float* adv, const float* ret, float* obs_o, float* act_o, float* adv_o,
                       float* ret_o, const float* old_src, float* old_dst, int pld, cudaStream_t st, const StepParams* sp,
                       const int* counter) {
    if (rows <= 0) return 0;
    int grid = (rows + 7) / 8;                  // 8 warps per CTA, one row per warp per pass
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope ps(KK_MISC, st);
    gather_rows_kernel<<<grid, 256, 0, st>>>(key, first, rows, obs_dim, act_dim, obs, act, adv, ret, obs_o, act_o, adv_o,
                                             ret_o, old_src, old_dst, pld, sp, counter);
    DPPO_LAUNCH_CHECK();
    return 0;
}

__global__ void perm_fill_kernel(PermKey key, int32_t* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < key.n; i += gridDim.x * blockDim.x)
        out[i] = (int32_t)perm_index(key, i);
}

int launch_perm_fill(const PermKey& key, int32_t* out, cudaStream_t st) {
    if (key.n == 0) return 0;
    int grid = (int)((key.n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope ps(KK_MISC, st);
    perm_fill_kernel<<<grid, 256, 0, st>>>(key, out);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// DiagGaussianPd.sample (distributions.py:204-205): a = mean + exp(logstd) * N(0,1)
// Philox4x32-10 counter RNG (key = seed, counter = (offset, row, quad)), Box-Muller.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void sample_kernel(const float* __restrict__ out, int rows, int A, int Pld, uint64_t seed, uint64_t offset,
                              float* __restrict__ action, float* __restrict__ v_out) {
    const int P = 1 + 2 * A;
    const int quads = (A + 3) / 4;
    const long long total = (long long)rows * quads;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int r = (int)(i / quads), q = (int)(i % quads);
        uint32_t c[4] = {(uint32_t)offset, (uint32_t)(offset >> 32), (uint32_t)r, (uint32_t)q};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float z[4];
        {
            const float r0 = sqrtf(-2.f * logf(u01(c[0]))), r1 = sqrtf(-2.f * logf(u01(c[2])));
            float s0, c0, s1, c1;
            sincosf(6.283185307179586f * u01(c[1]), &s0, &c0);
            sincosf(6.283185307179586f * u01(c[3]), &s1, &c1);
            z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
        }
        const float* o = out + (size_t)r * Pld;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = q * 4 + e;
            if (k < A) action[(size_t)r * A + k] = o[1 + k] + expf(o[1 + A + k]) * z[e];
        }
        if (q == 0 && v_out) v_out[r] = o[0];
    }
}

// ---------------------------------------------------------------------------------------------
// stand-alone DiagGaussianPd methods on a [B,2A] flat parameter tensor (distributions.py:193-205):
// op 0 neglogp(x)  1 entropy  2 kl(other)  3 sample  4 mode.   One thread per row.
// ---------------------------------------------------------------------------------------------
__global__ void diag_gaussian_kernel(int op, const float* __restrict__ flat, const float* __restrict__ arg,
                                     int B, int A, uint64_t seed, uint64_t offset, float* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const float* f = flat + (size_t)r * 2 * A;
    if (op == 0) {            // :195-198
        const float* x = arg + (size_t)r * A;
        float s2 = 0.f, sl = 0.f;
        for (int k = 0; k < A; ++k) { const float z = (x[k] - f[k]) * expf(-f[A + k]); s2 += z * z; sl += f[A + k]; }
        out[r] = 0.5f * s2 + 0.9189385332046727f * (float)A + sl;
    } else if (op == 1) {     // :202-203
        float sl = 0.f;
        for (int k = 0; k < A; ++k) sl += f[A + k];
        out[r] = sl + 1.4189385332046727f * (float)A;
    } else if (op == 2) {     // :199-201   KL(self || other)
        const float* q = arg + (size_t)r * 2 * A;
        float kl = 0.f;
        for (int k = 0; k < A; ++k) {
            const float dm = f[k] - q[k];
            kl += (q[A + k] - f[A + k]) + (expf(2.f * f[A + k]) + dm * dm) * 0.5f * expf(-2.f * q[A + k]) - 0.5f;
        }
        out[r] = kl;
    } else if (op == 3) {     // :204-205
        for (int q4 = 0; q4 < (A + 3) / 4; ++q4) {
            uint32_t c[4] = {(uint32_t)offset, (uint32_t)(offset >> 32), (uint32_t)r, (uint32_t)q4};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            const float r0 = sqrtf(-2.f * logf(u01(c[0]))), r1 = sqrtf(-2.f * logf(u01(c[2])));
            float s0, c0, s1, c1;
            sincosf(6.283185307179586f * u01(c[1]), &s0, &c0);
            sincosf(6.283185307179586f * u01(c[3]), &s1, &c1);
            const float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
            for (int e = 0; e < 4; ++e) {
                const int k = q4 * 4 + e;
                if (k < A) out[(size_t)r * A + k] = f[k] + expf(f[A + k]) * z[e];
            }
        }
    } else {                  // :193-194
        for (int k = 0; k < A; ++k) out[(size_t)r * A + k] = f[k];
    }
}

int launch_diag_gaussian(int op, const float* flat, const float* arg, int B, int A, uint64_t seed, uint64_t offset,
                         float* out, cudaStream_t st) {
    DPPO_REQUIRE(op >= 0 && op <= 4, "dppo_diag_gaussian: bad op %d", op);
    DPPO_REQUIRE(flat && out && (arg || (op != 0 && op != 2)), "dppo_diag_gaussian: null buffer");
    if (B == 0) return 0;
    ProfScope ps(KK_MISC, st);
    diag_gaussian_kernel<<<(B + 127) / 128, 128, 0, st>>>(op, flat, arg, B, A, seed, offset, out);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// CategoricalPd (distributions.py:124-159): sample = argmax(logits - log(-log(u))) (:154-156, Gumbel max; one Philox
// call yields 4 uniforms = 4 classes), mode = argmax (:129-130, first maximum).  One thread per row.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int gumbel_argmax(const float* __restrict__ lg, int ncat, int r, uint64_t seed, uint64_t offset) {
    float best = -INFINITY; int arg = 0;
    for (int q4 = 0; q4 < (ncat + 3) / 4; ++q4) {
        uint32_t c[4] = {(uint32_t)offset, (uint32_t)(offset >> 32), (uint32_t)r, (uint32_t)q4};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int e = 0; e < 4; ++e) {
            const int k = q4 * 4 + e;
            if (k < ncat) {
                const float g = lg[k] - logf(-logf(u01(c[e])));
                if (g > best) { best = g; arg = k; }
            }
        }
    }
    return arg;
}

__global__ void sample_cat_kernel(const float* __restrict__ out, int rows, int ncat, int Pld, uint64_t seed, uint64_t offset,
                                  int* __restrict__ action, float* __restrict__ v_out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* o = out + (size_t)r * Pld;
    action[r] = gumbel_argmax(o + 1, ncat, r, seed, offset);
    if (v_out) v_out[r] = o[0];
}

__global__ void categorical_kernel(int op, const float* __restrict__ logits, const void* __restrict__ arg, int B, int ncat,
                                   uint64_t seed, uint64_t offset, void* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const float* f = logits + (size_t)r * ncat;
    if (op == 3) { reinterpret_cast<int*>(out)[r] = gumbel_argmax(f, ncat, r, seed, offset); return; }
    float m = f[0]; int am = 0;
    for (int k = 1; k < ncat; ++k) if (f[k] > m) { m = f[k]; am = k; }
    if (op == 4) { reinterpret_cast<int*>(out)[r] = am; return; }
    float z = 0.f;
    for (int k = 0; k < ncat; ++k) z += expf(f[k] - m);
    const float lz = logf(z);
    float* o = reinterpret_cast<float*>(out);
    if (op == 0) {            // :131-138
        const int x = reinterpret_cast<const int*>(arg)[r];
        o[r] = (x >= 0 && x < ncat) ? lz - (f[x] - m) : 0.f;
    } else if (op == 1) {     // :148-153
        float e = 0.f;
        for (int k = 0; k < ncat; ++k) { const float a0 = f[k] - m; e += expf(a0) / z * (lz - a0); }
        o[r] = e;
    } else {                  // :139-147   KL(self || other)
        const float* q = reinterpret_cast<const float*>(arg) + (size_t)r * ncat;
        float mq = q[0];
        for (int k = 1; k < ncat; ++k) mq = fmaxf(mq, q[k]);
        float zq = 0.f;
        for (int k = 0; k < ncat; ++k) zq += expf(q[k] - mq);
        const float lzq = logf(zq);
        float kl = 0.f;
        for (int k = 0; k < ncat; ++k) { const float a0 = f[k] - m, a1 = q[k] - mq; kl += expf(a0) / z * (a0 - lz - a1 + lzq); }
        o[r] = kl;
    }
}

int launch_categorical(int op, const float* logits, const void* arg, int B, int ncat, uint64_t seed, uint64_t offset,
                       void* out, cudaStream_t st) {
    DPPO_REQUIRE(op >= 0 && op <= 4, "dppo_categorical: bad op %d", op);
    DPPO_REQUIRE(logits && out && (arg || (op != 0 && op != 2)), "dppo_categorical: null buffer");
    if (B == 0) return 0;
    ProfScope ps(KK_MISC, st);
    categorical_kernel<<<(B + 127) / 128, 128, 0, st>>>(op, logits, arg, B, ncat, seed, offset, out);
    DPPO_LAUNCH_CHECK();
    return 0;
}

int launch_sample(const NetShape& s, const float* out, int rows, uint64_t seed, uint64_t offset,
                  float* action, float* v_out, cudaStream_t st) {
    if (rows == 0) return 0;
    if (s.pd == DPPO_PD_CATEGORICAL) {
        ProfScope ps(KK_MISC, st);
        sample_cat_kernel<<<(rows + 127) / 128, 128, 0, st>>>(out, rows, s.act, s.Pld, seed, offset,
                                                             reinterpret_cast<int*>(action), v_out);
        DPPO_LAUNCH_CHECK();
        return 0;
    }
    const long long total = (long long)rows * ((s.act + 3) / 4);
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    ProfScope ps(KK_MISC, st);
    sample_kernel<<<grid, 256, 0, st>>>(out, rows, s.act, s.Pld, seed, offset, action, v_out);
    DPPO_LAUNCH_CHECK();
    return 0;
}

}  // namespace dppo
