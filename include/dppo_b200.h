/*
 * dppo_b200.h -- C ABI of libdppo_b200.so, the B200 (sm_100a) DPPO update engine.
 *
 * The reference (oswsnqc/Tensorflow-DPPO) has no FFI: its boundary for this path is the
 * Python/TensorFlow graph surface.  Every entry point below names the reference interface it
 * replaces (file:line under /root/reference).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference adds to call them from PPO.py / Chief.py / Worker.py.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; dppo_last_error() gives the message
 *     (thread-local, valid until the next call on that thread);
 *   - all tensor arguments are RAW DEVICE POINTERS owned by the caller, fp32, contiguous,
 *     row-major, 16-byte aligned; nothing is retained after the call except inside dppo_ctx;
 *   - every launch takes the CUDA stream (cudaStream_t as void*) to enqueue on and is
 *     asynchronous; no call allocates device memory (the caller passes the workspace whose size
 *     dppo_workspace_bytes() reports; the test hook dppo_dbg_gemm takes its scratch from it too) and
 *     none synchronises the host, except dppo_create, dppo_comm_init / _peer_export / _peer_import /
 *     _destroy (which own the exchange buffers), dppo_save_state / dppo_load_state and the *_host
 *     entry points;
 *   - one dppo_ctx per GPU / process; a ctx is not thread-safe;
 *   - layouts: rollout buffers are time-major, sample row b = t*N + n; weights are [in,out]
 *     row-major (TF `kernel`); the flat parameter / gradient vector is
 *     [W1,b1,...,WL,bL,Wv,bv,Wp,bp] (Model.py:12-14,17 creation order, trunk generalised to L
 *     hidden layers).  Head outputs: v[B], pdflat[B,2A] = [mean | logstd] (distributions.py:187).
 */
#ifndef DPPO_B200_H
#define DPPO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPPO_MAX_HIDDEN 8
#define DPPO_NUM_LOSS_SLOTS 8          /* policy, value, entropy, total, mean_kl, clip_frac, vclip_frac, rows */
#define DPPO_NCCL_UNIQUE_ID_BYTES 128
#define DPPO_PEER_HANDLE_BYTES 128     /* two cudaIpcMemHandle_t: exchange buffer, arrival flags */

/* GAE modes (Worker.py:82-92) */
#define DPPO_GAE_REFERENCE 0   /* mask 1-done[t+1]; bootstrap predv[T-1]*(1-done[T-1]) -- as written */
#define DPPO_GAE_STANDARD  1   /* mask 1-done[t];   bootstrap last_v*(1-last_done)                     */

/* arithmetic of the dense layers */
#define DPPO_GEMM_FP32_SIMT 0  /* exact fp32 FMA (parity path, every shape)                         */
#define DPPO_GEMM_TF32X3    1  /* tcgen05 kind::tf32, 3-term split (fp32-accurate tensor-core path)  */
#define DPPO_GEMM_BF16      2  /* tcgen05 kind::f16 bf16 operands, fp32 accumulate (config 4): the       */
                               /* persistent kernels of mode 3 in one-term mode (activations as bf16)    */
#define DPPO_GEMM_FP16X2    3  /* tcgen05 kind::f16, 2-term fp16 split x = hi + lo/2^11 (fp32-accurate, */
                               /* persistent kernels; operands must satisfy |x| < 65504)                */

typedef struct dppo_ctx dppo_ctx;

typedef struct dppo_config {
    int32_t obs_dim;                    /* env.observation_space.shape[0]   (PPO.py:13,17)        */
    int32_t act_dim;                    /* env.action_space.shape[0]        (distributions.py:234) */
    int32_t n_hidden;                   /* trunk depth L (Model.py:12 has L=1, width 16)           */
    int32_t hidden[DPPO_MAX_HIDDEN];    /* trunk widths                                            */
    int32_t max_rows;                   /* largest B (= T*N rows) any call will pass               */
    int32_t device;                     /* CUDA device ordinal                                     */
    int32_t gemm_mode;                  /* DPPO_GEMM_*                                             */
    float   entcoeff;                   /* parameter_dict['ENTCOEFF']   (main.py:15, PPO.py:11)    */
    float   vcoeff;                     /* parameter_dict['VCOEFF']     (main.py:16, PPO.py:12)    */
    float   clip_param;                 /* parameter_dict['CLIP_PARAM'] (main.py:17, PPO.py:10,19) */
    int32_t pd_type;                    /* DPPO_PD_*: what make_pdtype(env.action_space) returns    */
                                        /* (distributions.py:231-243)                               */
    int32_t value_loss_mode;            /* DPPO_VLOSS_*: per-sample value loss, or PPO.py:36-39 as    */
                                        /* literally written (pairwise broadcast, SURVEY A-D7)       */
} dppo_config;

/* Value loss (PPO.py:36-39).  As committed, `v` is [B,1,1] (Model.py:11,13) and `etr` [B,1] (PPO.py:28), so vfloss1/2
 * broadcast to [B,B,1]: every value is paired with every return of its tower (one tower = one worker's rollout).
 *   DPPO_VLOSS_PER_SAMPLE        the evident intent, (v_i - R_i)^2                     -- canonical, default;
 *   DPPO_VLOSS_LITERAL_PAIRWISE  the as-written numbers: mean over all pairs (i,j) of one rollout of
 *                                max((v_i - R_j)^2, (vclip_i - R_j)^2), averaged over the rollouts (tower mean).
 *                                dppo_loss_grad treats its B rows as ONE tower; dppo_update / dppo_update_host /
 *                                dppo_update_epochs (whole-buffer minibatches only) as N towers of T rows. */
#define DPPO_VLOSS_PER_SAMPLE       0
#define DPPO_VLOSS_LITERAL_PAIRWISE 1

/* Policy distribution of the pi/oldpi heads (Model.py:9,14-15):
 *   DPPO_PD_DIAG_GAUSSIAN  spaces.Box      -> DiagGaussianPd; act_dim = A, policy head 2A wide (mean | logstd),
 *                          actions are fp32 [B, A];
 *   DPPO_PD_CATEGORICAL    spaces.Discrete -> CategoricalPd (distributions.py:124-159); act_dim = ncat, policy head
 *                          ncat wide (logits), actions are int32 [B, 1] -- the reference's own `a` placeholder
 *                          (PPO.py:26) -- passed through the same `act` / `action_out` pointers (4-byte elements). */
#define DPPO_PD_DIAG_GAUSSIAN 0
#define DPPO_PD_CATEGORICAL   1

/* ---- library ------------------------------------------------------------------------------ */
const char* dppo_last_error(void);
int         dppo_version(void);

/* number of fp32 parameters of one net (pi or oldpi): tf.get_collection(TRAINABLE_VARIABLES,
 * scope) flattened -- Model.py:17 */
int64_t dppo_param_count(const dppo_config* cfg);
/* offsets[i] of the 2*(L+2) tensors in the flat vector, Model.py:17 order; returns their count */
int     dppo_param_offsets(const dppo_config* cfg, int64_t* offsets, int capacity);
/* length of the gradient buffer dppo_loss_grad writes: n_params + DPPO_NUM_LOSS_SLOTS (the loss
 * scalars ride at the tail so that ONE allreduce averages both -- SURVEY 8(e)) */
int64_t dppo_grad_buffer_len(const dppo_config* cfg);
/* bytes of device workspace dppo_create needs for this config */
size_t  dppo_workspace_bytes(const dppo_config* cfg);

/* replaces: tf.Session() + the PPO graph construction (main.py:31, PPO.py:8-53) */
int dppo_create(const dppo_config* cfg, void* workspace, size_t workspace_bytes, dppo_ctx** out);
int dppo_destroy(dppo_ctx* ctx);

/* ---- the hot path -------------------------------------------------------------------------- */

/* replaces Worker.py:82-92 for N rollouts at once.  rew/predv/done/adv_out/ret_out are [T,N]
 * time-major (done is 0/1 as fp32).  last_v/last_done ([N]) are read in STANDARD mode only and
 * may be NULL in REFERENCE mode.  normalize!=0 applies Worker.py:92 per rollout (column). */
int dppo_gae(dppo_ctx* ctx, const float* rew, const float* predv, const float* done,
             const float* last_v, const float* last_done, int T, int N, double gamma, double lam,
             int mode, int normalize, float* adv_out, float* ret_out, void* stream);

/* replaces sess.run([pipredv, piPD.flatparam()], {s: obs}) -- Model.py:11-15 via PPO.py:21,
 * Worker.py:146.  v_out[B], pdflat_out[B,2A]; either may be NULL. */
int dppo_forward(dppo_ctx* ctx, const float* params, const float* obs, int B,
                 float* v_out, float* pdflat_out, void* stream);

/* replaces sess.run([policyLoss, valueLoss, entropyLoss, total_loss, gradient], feed_dict) for
 * one worker tower: pi forward, oldpi forward (PPO.py:21-22), DiagGaussianPd logp/entropy
 * (distributions.py:195-203), the clipped losses (PPO.py:29-40) and compute_gradients w.r.t.
 * pipara (PPO.py:46).  obs[B,obs_dim], act[B,act_dim], adv[B], ret[B].
 * grad_out has dppo_grad_buffer_len() floats: d total_loss / d params, then the
 * DPPO_NUM_LOSS_SLOTS scalars.  losses_out (8 floats) may be NULL. */
int dppo_loss_grad(dppo_ctx* ctx, const float* params_pi, const float* params_old,
                   const float* obs, const float* act, const float* adv, const float* ret,
                   int B, float l_mul, float* losses_out, float* grad_out, void* stream);

/* replaces PPO._average_gradients (PPO.py:55-65; towers fed by Chief.py:54-64): in-place mean
 * over ranks of buf[0:n] with ncclAllReduce(ncclAvg) on NVLink/NVSwitch. */
int dppo_comm_unique_id(void* id_out_128_bytes);
int dppo_comm_init(dppo_ctx* ctx, int rank, int world, const void* id_128_bytes);
int dppo_comm_destroy(dppo_ctx* ctx);
int dppo_allreduce_avg(dppo_ctx* ctx, float* buf, int64_t n, void* stream);
/* The same tower mean without NCCL, for hosts that bring their own bootstrap: the fused peer-memory exchange of
 * dppo_update / dppo_update_host / dppo_update_epochs (one kernel: arrive/wait over NVLink, mean in rank order, Adam --
 * PPO.py:53,55-65) only needs every rank's two cudaIpc handles.  export allocates this rank's exchange buffers and writes
 * DPPO_PEER_HANDLE_BYTES to handle_out; the host gathers the blobs of all ranks in rank order by any means (file, pipe,
 * MPI, a socket) and passes the world * DPPO_PEER_HANDLE_BYTES bytes to import.  Ranks may share one device
 * (two processes on one GPU), which is how the exchange kernel is tested on a single-GPU box.  2 <= world <= 16. */
int dppo_comm_peer_export(dppo_ctx* ctx, int rank, int world, void* handle_out);
int dppo_comm_peer_import(dppo_ctx* ctx, const void* all_handles);

/* replaces Optimizer.apply_gradients (PPO.py:20,53): tf.train.AdamOptimizer step in TF-1.x
 * form, lr = LEARNING_RATE*l_mul, step = t after increment (1-based). */
int dppo_adam_apply(dppo_ctx* ctx, float* params, const float* grad, float* m, float* v,
                    int64_t n, int step, float lr, float beta1, float beta2, float eps,
                    void* stream);

/* replaces sync_pis (PPO.py:47): oldpi <- pi */
int dppo_sync_old(dppo_ctx* ctx, const float* params_pi, float* params_old, void* stream);

/* replaces sess.run([ca, pipredv], {s: obs}) for a batch of states: pi forward +
 * DiagGaussianPd.sample (PPO.py:23, distributions.py:204-205, Worker.py:146).
 * action_out[B,A], v_out[B]; Philox counter = (seed, offset). */
int dppo_act(dppo_ctx* ctx, const float* params, const float* obs, int B, uint64_t seed,
             uint64_t offset, float* action_out, float* v_out, void* stream);

/* one whole update on DEVICE-resident rollout buffers in a single call -- what one iteration of
 * [sess.run(self.ppo.train, feed_dict) for _ in range(UPDATE_STEPS)] (Chief.py:64) does, plus the GAE of
 * Worker.py:82-92: GAE (on a side stream, under the forward GEMMs) -> pi/oldpi forward -> loss -> backward ->
 * allreduce when a communicator exists -> Adam when adam_m/adam_v are non-NULL (step = t after increment).
 * grad_out as in dppo_loss_grad.  Asynchronous on `stream`. */
int dppo_update(dppo_ctx* ctx, float* params_pi, const float* params_old, const float* obs, const float* act,
                const float* rew, const float* predv, const float* done, int T, int N, double gamma, double lam,
                int gae_mode, float l_mul, float* adam_m, float* adam_v, int step, float lr, float beta1,
                float beta2, float eps, float* grad_out, void* stream);

/* replaces the CategoricalPd methods as stand-alone graph ops (distributions.py:124-159) on a [B,ncat] logits tensor:
 * op 0 neglogp(arg = int32 x[B]) -> out[B] (out-of-range x gives 0, as tf.one_hot does); 1 entropy -> out[B];
 * 2 kl(arg = other logits [B,ncat]) -> out[B]; 3 sample (Gumbel-max, Philox seed/offset) -> int32 out[B];
 * 4 mode (argmax, first maximum) -> int32 out[B]. */
int dppo_categorical(dppo_ctx* ctx, int op, const float* logits, const void* arg, int B, int ncat,
                     uint64_t seed, uint64_t offset, void* out, void* stream);

/* ---- epoch / minibatch scheduler on the device ------------------------------------
 * The reference trains on whole rollouts, UPDATE_STEPS times on the same feed (Chief.py:58-64); with
 * n_epochs = UPDATE_STEPS, minibatch_rows = T*N, shuffle = 0 this call IS that loop (GAE once, then
 * UPDATE_STEPS x [forward, loss, backward, tower mean, Adam], gradients at the current parameters).  With
 * minibatch_rows < T*N every epoch is cut into ceil(T*N / minibatch_rows) minibatches (the last one may be short;
 * every loss is the mean over the minibatch's own rows), contiguous row slices when shuffle = 0, rows drawn
 * through the keyed permutation dppo_minibatch_perm(T*N, seed, epoch) when shuffle = 1 (gathered on the device,
 * no host involvement, the whole call is asynchronous on `stream` and graph-capturable).
 * Adam step numbers are step0, step0+1, ... (step0 = t after increment of the first minibatch).
 * losses_out: device [n_epochs * n_minibatches, 8] (same 8 scalars as dppo_loss_grad) or NULL.
 * grad_out [n_params + 8] holds the last minibatch's (all-reduced) gradient on return.
 * oldpi is frozen during the call (sync_pis happens in the worker, PPO.py:47): its head outputs are evaluated once
 * and the rows reused by every minibatch -- the same numbers the reference's in-graph re-evaluation (PPO.py:22) gives.
 * With a communicator every rank must pass the same n_epochs/minibatch_rows/T/N (one exchange per minibatch). */
int dppo_update_epochs(dppo_ctx* ctx, float* params_pi, const float* params_old, const float* obs, const float* act,
                       const float* rew, const float* predv, const float* done, int T, int N, double gamma,
                       double lam, int gae_mode, int n_epochs, int minibatch_rows, int shuffle, uint64_t seed,
                       float l_mul, float* adam_m, float* adam_v, int step0, float lr, float beta1, float beta2,
                       float eps, float* losses_out, float* grad_out, void* stream);

/* perm_out[i] = position i of epoch `epoch`'s row permutation of [0, n): a keyed 4-round Feistel bijection with
 * cycle walking (csrc/perm.cuh; restated in oracle/dppo_oracle.py:minibatch_perm).  Device int32[n]. */
int dppo_minibatch_perm(dppo_ctx* ctx, int n, uint64_t seed, uint64_t epoch, int32_t* perm_out, void* stream);

/* replaces the DiagGaussianPd methods as stand-alone graph ops (distributions.py:193-205) on a
 * [B,2A] flat parameter tensor: op 0 neglogp(arg=x[B,A]) -> out[B]; 1 entropy -> out[B];
 * 2 kl(arg=other flat[B,2A]) -> out[B]; 3 sample (Philox seed/offset) -> out[B,A]; 4 mode -> out[B,A]. */
int dppo_diag_gaussian(dppo_ctx* ctx, int op, const float* flat, const float* arg, int B, int A,
                       uint64_t seed, uint64_t offset, float* out, void* stream);

/* the whole update from HOST buffers (what Chief._train feeds through feed_dict,
 * Chief.py:54-64): pinned host rollout -> chunked H2D overlapped with GAE + loss_grad ->
 * (allreduce if a communicator exists) -> Adam when adam_m/adam_v are non-NULL (PPO.py:53; step = t after
 * increment, as dppo_update) -> losses back to host.  Synchronises `stream`.
 * params_pi, params_old, adam_m, adam_v and grad_out are device pointers; every ..._h argument is host memory. */
int dppo_update_host(dppo_ctx* ctx, float* params_pi, const float* params_old,
                     const float* obs_h, const float* act_h, const float* rew_h,
                     const float* predv_h, const float* done_h, int T, int N,
                     double gamma, double lam, int gae_mode, float l_mul,
                     float* adam_m, float* adam_v, int step, float lr, float beta1, float beta2, float eps,
                     float* losses_out_h, float* grad_out, void* stream);

/* replaces the whole of Chief._train (Chief.py:54-64) for HOST-resident rollout batches: n_towers feeds (one per worker,
 * Chief.py:58-63) of rows[k] samples each -- obs_h[k] [rows,obs_dim], act_h[k] [rows,act cols], adv_h[k] [rows],
 * ret_h[k] [rows], host memory (pinned memory makes the upload asynchronous; pageable works) -- are uploaded ONCE, in
 * row chunks overlapped with the first step's compute, and trained on n_steps (= UPDATE_STEPS, main.py:24) times on the
 * device: [pi forward, loss, backward, tower mean (+ exchange across ranks), Adam] per step, step numbers step0,
 * step0+1, ...; oldpi is evaluated once (it is frozen while the chief trains, PPO.py:47 / Worker.py:42).
 * The tower mean of PPO.py:55-65 is taken as the mean over all rows (towers of equal length, as the reference's are).
 * losses_out_h: host [n_steps, 8] (may be NULL).  grad_out: the last step's averaged gradient.  Synchronises `stream`.
 * Per-sample value loss only (DPPO_VLOSS_LITERAL_PAIRWISE needs the [T,N] structure: use dppo_update). */
int dppo_train_host(dppo_ctx* ctx, float* params_pi, const float* params_old, int n_towers, const float* const* obs_h,
                    const void* const* act_h, const float* const* adv_h, const float* const* ret_h, const int* rows,
                    int n_steps, float l_mul, float* adam_m, float* adam_v, int step0, float lr, float beta1, float beta2,
                    float eps, float* losses_out_h, float* grad_out, void* stream);

/* ---- checkpoint / resume -------------------------------------------------------------------------
 * replaces save_state / load_state (Others/tf_util.py:271-279: tf.train.Saver over the session's variables): the flat
 * parameter vectors of pi and oldpi (Model.py:17 order), the Adam slots m and v (PPO.py:20) and the optimiser step t.
 * File: 128-byte header (magic "DPPOB200", version, obs_dim, act_dim, pd_type, n_hidden, hidden[8], n_params, step,
 * which arrays are present) followed by the fp32 arrays; written to <path>.tmp and renamed.  adam_m / adam_v may be NULL
 * (not saved / not restored).  load fails loudly when the header does not match the context's network.
 * Both calls synchronise `stream`. */
int dppo_save_state(dppo_ctx* ctx, const char* path, const float* params_pi, const float* params_old,
                    const float* adam_m, const float* adam_v, int64_t step, void* stream);
int dppo_load_state(dppo_ctx* ctx, const char* path, float* params_pi, float* params_old, float* adam_m,
                    float* adam_v, int64_t* step_out, void* stream);

/* ---- launch accounting -------------------------------------------------------------------------
 * Per context: every kernel the library launches on behalf of `ctx` is counted per family; with enable_timing != 0 each launch
 * is also bracketed by CUDA events on its own stream (used by bench.py for the roofline line).
 * Families, in order: gae, gemm_nn (forward), gemm_nt (dX), gemm_tn (dW), loss, finalize, misc,
 * adam, allreduce, tc_gemm.  dppo_profile_read synchronises the recorded events and returns the
 * number of families written. */
#define DPPO_PROFILE_KINDS 10
int dppo_profile_reset(dppo_ctx* ctx, int enable_timing);
int dppo_profile_read(dppo_ctx* ctx, int64_t* launches, double* ms, int capacity);

/* ---- unit-test hooks (exported so the parity tests can exercise kernels in isolation) ------ */
/* C[M,N] = epilogue(A . B).  mode 0: A[M,K] B[K,N] (+bias, relu flag);  mode 1: A[M,K] B[N,K]^T,
 * C *= (mask[M,N] > 0) when mask != NULL;  mode 2: C[K?]: A[Kr,M]^T . B[Kr,N] with split-K over Kr
 * (deterministic), bias_out[N] = column sums of B when non-NULL. */
int dppo_dbg_gemm(dppo_ctx* ctx, int mode, int M, int N, int K, const float* A, const float* B,
                  float* C, const float* bias, const float* mask, int relu, float* bias_out,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPPO_B200_H */
