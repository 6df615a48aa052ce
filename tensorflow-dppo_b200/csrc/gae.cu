// K1/K2: GAE(lambda) reverse scan + returns + per-rollout advantage standardisation.
//
// Replaces /root/reference/Worker.py:82-92 (python loop over T per worker thread) for all N
// rollouts of a [T,N] time-major buffer in one launch.
//
// The recurrence A_t = delta_t + (gamma*lam*nt_t) * A_{t+1} is an affine map in A_{t+1}; affine
// maps compose associatively, so a CTA that owns kCols rollouts splits T into kChunks chunks:
//   pass 1  every thread composes the affine map of its (chunk, column)            (parallel)
//   carry   kCols threads walk the kChunks composites backwards (one DFMA per chunk) (serial, short)
//   pass 2  every thread replays its chunk from the true carry-in -> adv, ret       (parallel)
//   pass 3  per-column mean / std (ddof 0, no eps: Worker.py:92) and normalisation  (parallel)
// Loads are sector-coalesced: consecutive threads read consecutive columns of one time row
// (kCols*4 B = one 32 B sector).  The scan arithmetic is fp64, like the reference's python
// scalars (Worker.py:88-90); stores are fp32 (Worker.py:85).  The buffer (12 B/step in, 8 B/step
// out) is L2-resident, so the kernel is latency-bound, not HBM-bound -- see DESIGN.md.
#include "common.cuh"

namespace dppo {

constexpr int kCols = 8;
constexpr int kChunks = 128;

struct GaeArgs {
    const float* rew; const float* predv; const float* done;
    const float* last_v; const float* last_done;
    float* adv; float* ret;
    int T, N, mode, normalize;
    double gamma, gl;   // gamma, gamma*lam
};

__device__ __forceinline__ void gae_step_inputs(const GaeArgs& a, int t, int n, double& delta,
                                                double& coef) {
    const int T = a.T, N = a.N;
    const double v_t = (double)a.predv[(size_t)t * N + n];
    double nt, v_next;
    if (a.mode == DPPO_GAE_REFERENCE) {
        // Worker.py:82-83,87-88: done_ext[t+1], predv_tmp[t+1]; appended tail = (0, v[T-1]*(1-done[T-1]))
        if (t + 1 < T) {
            nt = 1.0 - (double)a.done[(size_t)(t + 1) * N + n];
            v_next = (double)a.predv[(size_t)(t + 1) * N + n];
        } else {
            nt = 1.0;
            v_next = v_t * (1.0 - (double)a.done[(size_t)t * N + n]);
        }
    } else {
        nt = 1.0 - (double)a.done[(size_t)t * N + n];
        if (t + 1 < T) v_next = (double)a.predv[(size_t)(t + 1) * N + n];
        else v_next = (double)a.last_v[n] * (1.0 - (double)a.last_done[n]);
    }
    delta = (double)a.rew[(size_t)t * N + n] + a.gamma * v_next * nt - v_t;   // Worker.py:89
    coef = a.gl * nt;                                                          // Worker.py:90
}

__global__ void __launch_bounds__(kCols* kChunks)
gae_scan_kernel(GaeArgs a) {
    __shared__ double s_alpha[kChunks][kCols];
    __shared__ double s_beta[kChunks][kCols];   // reused: carry-in, then partial sums
    __shared__ double s_mean[kCols], s_rstd[kCols];

    const int j = threadIdx.x % kCols;
    const int c = threadIdx.x / kCols;
    const int n = blockIdx.x * kCols + j;
    const int T = a.T;
    const int Lc = (T + kChunks - 1) / kChunks;
    const int t0 = c * Lc;
    const int t1 = min(T, t0 + Lc);
    const bool live = (n < a.N);

    // pass 1: composite affine map of this chunk
    double alpha = 1.0, beta = 0.0;
    if (live) {
        for (int t = t1 - 1; t >= t0; --t) {
            double delta, coef;
            gae_step_inputs(a, t, n, delta, coef);
            beta = fma(coef, beta, delta);
            alpha *= coef;
        }
    }
    s_alpha[c][j] = alpha;
    s_beta[c][j] = beta;
    __syncthreads();

    // carry: A at the first step AFTER each chunk (A_T = 0)
    if (c == 0) {
        double x = 0.0;
        for (int k = kChunks - 1; k >= 0; --k) {
            const double al = s_alpha[k][j], be = s_beta[k][j];
            s_beta[k][j] = x;                 // carry-in of chunk k
            x = fma(al, x, be);
        }
    }
    __syncthreads();

    // pass 2: replay with the true carry-in
    double psum = 0.0;
    if (live) {
        double x = s_beta[c][j];
        for (int t = t1 - 1; t >= t0; --t) {
            double delta, coef;
            gae_step_inputs(a, t, n, delta, coef);
            x = fma(coef, x, delta);
            const float adv32 = (float)x;                                        // Worker.py:85,90
            const size_t idx = (size_t)t * a.N + n;
            a.ret[idx] = (float)((double)adv32 + (double)a.predv[idx]);          // Worker.py:91
            a.adv[idx] = adv32;
            psum += (double)adv32;
        }
    }
    if (!a.normalize) return;

    // pass 3: (adv - mean) / std over the column, population std, no epsilon (Worker.py:92)
    __syncthreads();
    s_alpha[c][j] = psum;
    __syncthreads();
    if (c == 0) {
        double s = 0.0;
        for (int k = 0; k < kChunks; ++k) s += s_alpha[k][j];
        s_mean[j] = s / (double)T;
    }
    __syncthreads();
    const double mean = s_mean[j];
    double pvar = 0.0;
    if (live) {
        for (int t = t0; t < t1; ++t) {
            const double d = (double)a.adv[(size_t)t * a.N + n] - mean;
            pvar = fma(d, d, pvar);
        }
    }
    s_beta[c][j] = pvar;
    __syncthreads();
    if (c == 0) {
        double s = 0.0;
        for (int k = 0; k < kChunks; ++k) s += s_beta[k][j];
        s_rstd[j] = 1.0 / sqrt(s / (double)T);
    }
    __syncthreads();
    if (live) {
        const double rstd = s_rstd[j];
        for (int t = t0; t < t1; ++t) {
            const size_t idx = (size_t)t * a.N + n;
            a.adv[idx] = (float)(((double)a.adv[idx] - mean) * rstd);
        }
    }
}

int launch_gae(const float* rew, const float* predv, const float* done, const float* last_v,
               const float* last_done, int T, int N, double gamma, double lam, int mode,
               int normalize, float* adv, float* ret, cudaStream_t st) {
    DPPO_REQUIRE(T > 0 && N > 0, "dppo_gae: T=%d N=%d must be positive", T, N);
    DPPO_REQUIRE(mode == DPPO_GAE_REFERENCE || mode == DPPO_GAE_STANDARD, "dppo_gae: bad mode %d", mode);
    DPPO_REQUIRE(mode == DPPO_GAE_REFERENCE || (last_v && last_done),
                 "dppo_gae: STANDARD mode needs last_v and last_done");
    GaeArgs a;
    a.rew = rew; a.predv = predv; a.done = done; a.last_v = last_v; a.last_done = last_done;
    a.adv = adv; a.ret = ret; a.T = T; a.N = N; a.mode = mode; a.normalize = normalize;
    a.gamma = gamma; a.gl = gamma * lam;
    const int grid = (N + kCols - 1) / kCols;
    ProfScope ps(KK_GAE, st);
    gae_scan_kernel<<<grid, kCols * kChunks, 0, st>>>(a);
    DPPO_LAUNCH_CHECK();
    return 0;
}

}  // namespace dppo
