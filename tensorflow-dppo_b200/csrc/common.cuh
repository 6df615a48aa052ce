// Shared declarations of libdppo_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/dppo_b200.h"
#include "perm.cuh"

namespace dppo {

void set_error(const char* fmt, ...);

#define DPPO_CHECK_CUDA(expr)                                                             \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            dppo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                            cudaGetErrorString(_e));                                      \
            return -2;                                                                    \
        }                                                                                 \
    } while (0)

#define DPPO_REQUIRE(cond, ...)                                                           \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            dppo::set_error(__VA_ARGS__);                                                 \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

#define DPPO_LAUNCH_CHECK() DPPO_CHECK_CUDA(cudaGetLastError())

constexpr int kMaxLayers = DPPO_MAX_HIDDEN;
constexpr int kLossSlots = DPPO_NUM_LOSS_SLOTS;
constexpr int kLossTileRows = 128;          // rows per CTA of the loss kernel
constexpr int kMaxSplitK = 148;             // upper bound of dW split-K slices per layer
constexpr int kMaxTrainSteps = 64;          // UPDATE_STEPS bound of dppo_train_host (pinned loss read-back slots)
constexpr int kPlanSMs = 148;               // B200: the split-K plan fills 2 CTAs/SM x 148 SMs in whole waves

// Internal parameter layout: identical to the public one for the trunk; the two heads are packed
// into ONE [H_L, P] matrix (P = 1 + 2A: value | mean | logstd) + bias[P], so that the heads are a
// single GEMM.  Same total size as the public head block.
struct NetShape {
    int obs, act, L, P;
    int pd;                          // DPPO_PD_*
    int pdw;                         // policy-head width: 2A (mean | logstd) or ncat (logits); P = 1 + pdw
    int acols;                       // 4-byte columns of an action row: A (fp32) or 1 (int32 class index)
    int Pld;                         // row stride of the [rows, P] head buffers (P rounded up to 4: 16-B rows)
    int h[kMaxLayers];
    int64_t w_off[kMaxLayers + 1];   // trunk W_l offsets, [L] = packed head W
    int64_t b_off[kMaxLayers + 1];   // trunk b_l offsets, [L] = packed head b
    int64_t n_params;
    // dW/db split-K plan (deterministic two-stage reduction): layer l's partials are slices[l] copies of
    // its contiguous [W_l | b_l] region (region[l] floats, padded to 4), starting at part_off[l]
    int slices[kMaxLayers + 1];
    int64_t region[kMaxLayers + 1];
    int64_t part_off[kMaxLayers + 1];
    int64_t partial_floats;
    int in_dim(int l) const { return l == 0 ? obs : h[l - 1]; }
    int out_dim(int l) const { return l == L ? P : h[l]; }
    int hmax() const { int m = 0; for (int i = 0; i < L; ++i) m = h[i] > m ? h[i] : m; return m; }
};

int make_shape(const dppo_config* cfg, NetShape* s);

// Per-step values of the device-scheduled minibatch loop, read by the kernels of a CUDA-graph replay through
// params[*counter] (dppo_update_epochs): everything else of a minibatch step is identical from step to step.
struct StepParams {
    PermKey key;            // row permutation of this step's epoch
    int first;              // first permuted index of the minibatch
    int rows;
    float lr_t;             // Adam step size with the bias corrections of this step number
    int loss_slot;          // row of losses_out this step reports to
    unsigned peer_epoch;    // arrival epoch of the fused peer exchange
    int peer_sel;           // which half of the double-buffered exchange buffer this step writes / reads
    int pad[4];
};
static_assert(sizeof(StepParams) == 64, "StepParams is one 64-byte record");
constexpr int kMaxGraphSteps = 4096;

struct Workspace {
    // all device pointers into the caller's workspace
    float* head_pi;      // packed head W|b of pi      [H_L*P + P]
    float* head_old;     // packed head W|b of oldpi
    float* wt_pi;        // bf16 path: transposed kernels Wt[out][in] of pi (internal layout offsets)
    float* wt_old;       // ... of oldpi
    // tensor-core path: per-layer weight images (B operand tiles as they sit in smem, hi|lo), see tc_pack_images
    float* img_fwd_pi[kMaxLayers + 1];
    float* img_fwd_old[kMaxLayers + 1];
    float* img_dx_pi[kMaxLayers + 1];
    float* act_pi[kMaxLayers];   // post-ReLU activations of pi, layer l: [max_rows, h[l]]
    uint32_t* act_bits[kMaxLayers];  // tensor-core path: their sign bitmaps (the dX ReLU mask; layout private to tc_gemm.cu)
    float* scratch[2];   // oldpi activations / dh ping-pong: [max_rows, hmax]
    float* out_pi;       // [max_rows, P]
    float* out_old;      // [max_rows, P]
    float* d_out;        // [max_rows, P]
    float* partial;      // split-K partials, see NetShape::part_off
    double* loss_part;   // [loss_tiles_max, kLossSlots]
    float* gae_adv;      // [max_rows] used by dppo_update_host
    float* gae_ret;      // [max_rows]
    float* obs_pl;       // bf16 mode (one-term kernels): the observations of the current row block as ONE bf16 plane [rows, obs]
    float* obs_pad;      // fp16x2 mode, obs_dim % 4 != 0: the rows of the current block re-pitched to a multiple of 4 floats (TMA)
    int obs_ld;          // its pitch in floats (0: not used)
    float* stage;        // device staging for dppo_update_host: obs|act|rew|predv|done
    size_t stage_floats;
    StepParams* step_params;   // [kMaxGraphSteps] per-step records of the graph-replayed minibatch loop
    int* step_counter;         // index of the step the next replay executes
};

}  // namespace dppo

struct dppo_ctx {
    dppo_config cfg;
    dppo::NetShape shape;
    dppo::Workspace ws;
    void* workspace;
    size_t workspace_bytes;
    int num_sms;
    // communicator (comm.cu)
    void* nccl_comm;
    int rank, world;
    // NVLink peer-memory exchange (comm.cu): this rank's gradient exchange buffers (double-buffered), every
    // rank's mapped copy of them, and the arrival flags
    float* peer_self;            // [2][grad_len] owned by this rank (cudaMalloc + cudaIpc)
    float** peer_bufs_dev;       // device array [world] of mapped base pointers (index rank == peer_self)
    void* peer_mapped[16];       // host copy of the mapped pointers (for cudaIpcCloseMemHandle)
    unsigned int* flags_self;    // [world] arrival epochs written by the peers, + [1] local go word
    unsigned int** flags_dev;    // device array [world] of mapped flag arrays
    void* flags_mapped[16];
    unsigned int epoch;
    int64_t peer_len;            // floats per exchange buffer
    // update_host plumbing
    cudaStream_t copy_stream;
    cudaEvent_t chunk_ev[64];
    cudaEvent_t done_ev;
    float* losses_pinned;
    // dppo_update_epochs: one minibatch step captured as a CUDA graph and replayed (small minibatches are launch-bound)
    void* step_graph_exec;       // cudaGraphExec_t
    unsigned char step_graph_key[160];
    long long step_graph_counts[16];   // kernel launches per replay, per family
    int use_graph;               // DPPO_GRAPH != 0 at dppo_create
    int use_h_planes;            // DPPO_H_PLANES: fp16x2 mode keeps hidden activations / gradients as fp16 hi | lo planes (capi.cu: h_planes)
    int use_h_dw;                // DPPO_H_DW: dW of the fp16x2 mode on the persistent fp16-split kernel (else the tf32 TMA kernel)
    int no_old_cache;            // DPPO_NO_OLD_CACHE=1 at dppo_create: dppo_update_epochs re-evaluates oldpi per minibatch
    void* prof;                  // launch accounting of this context (capi.cu: ProfState)
};

namespace dppo {

// ---- launch accounting (dppo_profile_*): every kernel launch of the library passes through a
// ProfScope; counters are always on, CUDA-event timing only when enabled (bench.py's roofline pass)
enum KernelKind { KK_GAE = 0, KK_GEMM_NN, KK_GEMM_NT, KK_GEMM_TN, KK_LOSS, KK_FINALIZE, KK_MISC, KK_ADAM,
                  KK_ALLREDUCE, KK_TC_GEMM, KK_COUNT };
struct ProfScope {
    int kind; cudaStream_t st; int slot;
    ProfScope(int kind, cudaStream_t st);
    ~ProfScope();
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: remember per device ordinal what has been configured,
// so that one process driving several GPUs (one dppo_ctx each) configures every one of them.
struct PerDeviceOnce {
    unsigned long long mask = 0;
    bool need() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if ((mask >> d) & 1ull) return false;
        mask |= 1ull << d;
        return true;
    }
};
struct PerDeviceMax {
    size_t cur[64] = {};
    bool need(size_t bytes) {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (bytes <= cur[d]) return false;
        cur[d] = bytes;
        return true;
    }
};

// ---- kernels' host launchers (one per .cu) ----
int launch_gae(const float* rew, const float* predv, const float* done, const float* last_v,
               const float* last_done, int T, int N, double gamma, double lam, int mode,
               int normalize, float* adv, float* ret, cudaStream_t st);

// C = epi(A.B) family, see sgemm.cu
struct GemmEpi {
    const float* bias = nullptr;   // [N] added before relu (NN)
    int relu = 0;
    const float* mask = nullptr;   // [M,N] (ld = ldc): C *= (mask > 0)   (NT)
    float* bias_out = nullptr;     // TN: column sums of B -> partial slices, stride = slice_stride
    int accumulate = 0;            // TN: partial += instead of =
    int64_t slice_stride = 0;      // TN: floats between split-K slices of C / bias_out
    int slices = 1;                // TN: number of split-K slices
};
int gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const GemmEpi& e, cudaStream_t st);
// nets (1 or 2) independent products sharing shapes (pi and oldpi forward in one launch)
int gemm_nn_dual(int M, int N, int K, const float* A0, const float* A1, int lda, const float* B0,
                 const float* B1, int ldb, float* C0, float* C1, int ldc, const float* bias0,
                 const float* bias1, int relu, int nets, cudaStream_t st);
int gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const GemmEpi& e, cudaStream_t st);
// C[s][M,N] (+)= A[k, M]^T . B[k, N] over k in slice s of [0,K); e.slices slices
int gemm_tn_splitk(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                   int ldc, const GemmEpi& e, cudaStream_t st);

// tensor-core products (tc_gemm.cu)
int tc_gemm_kk(int terms, int epi_nt, int M, int N, int K, const float* A0, const float* A1, int lda, const float* B0,
               const float* B1, int ldb, float* C0, float* C1, int ldc, const float* bias0, const float* bias1,
               int relu, const float* mask, int nets, cudaStream_t st);
int bf16_gemm_kk(int epi_nt, int M, int N, int K, const float* A0, const float* A1, int lda, const float* B0, const float* B1,
                 int ldb, float* C0, float* C1, int ldc, const float* bias0, const float* bias1, int relu, const float* mask,
                 int nets, cudaStream_t st);
int bf16_gemm_tn_splitk(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int slices,
                        long long slice_stride, int accumulate, cudaStream_t st);
long long tc_image_floats(int N, int K);
int tc_pack_images(const NetShape& s, const float* params_pi, const float* head_pi, const float* params_old,
                   const float* head_old, float* const* img_fwd_pi, float* const* img_fwd_old, float* const* img_dx_pi,
                   cudaStream_t st);
int tc_pack_one(const float* src, float* dst, int N, int K, int ld, int transposed, cudaStream_t st);
bool tc_bits_supported();
// A kind::tf32 MMA reads only the upper 19 bits of an fp32 operand (verified on sm_100a: the parity tests pass with raw
// tiles), i.e. it sees trunc(x).  The activation split can therefore be hi = x as it is (no conversion, and for the
// TMA-fed dW kernel no rewrite of the tile at all) and lo = RN-tf32(x - trunc(x)).  DPPO_TC_RAWHI=0 restores hi = RN-tf32(x).
bool tc_raw_hi();
int tc_gemm_kk_img(int epi_nt, int M, int N, int K, const float* A0, const float* A1, int lda, const float* img0,
                   const float* img1, float* C0, float* C1, int ldc, const float* bias0, const float* bias1, int relu,
                   const float* mask, int nets, cudaStream_t st, uint32_t* bits_out0 = nullptr,
                   const uint32_t* bits_in = nullptr);
long long tc_bits_words(long long rows, int N);
int tc_gemm_tn_splitk(int terms, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                      int slices, long long slice_stride, int accumulate, cudaStream_t st);
bool tc_tn_tma_supported(const float* A, int lda, const float* B, int ldb);
int tc_gemm_tn_tma(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, float* bias_out,
                   int slices, long long slice_stride, int accumulate, cudaStream_t st);
// persistent fp16-split tensor-core products (tc_gemm_h.cu, DPPO_GEMM_FP16X2)
long long th_image_floats(int N, int K);
long long th_bits_words(long long rows, int N);
int th_pack_images(const NetShape& s, const float* params_pi, const float* head_pi, const float* params_old,
                   const float* head_old, float* const* img_fwd_pi, float* const* img_fwd_old, float* const* img_dx_pi,
                   cudaStream_t st, int terms = 2, int rows = 1 << 30, int planes = 0);     // rows / planes: of the products these images feed
int th_pack_one(const float* src, float* dst, int N, int K, int ld, int transposed, cudaStream_t st, int terms = 2, int rows = 1 << 30,
                int a_planes = 0);
// Activations kept as two fp16 planes (x ~= hi + lo / 2^11, the split of tc_gemm_h.cu) instead of one fp32 matrix: the
// consumer's TMA writes the MMA operand tiles directly, no converter warps.  A "planes" pointer is the hi plane (passed
// through the float* parameter), the lo plane starts *_lo_off fp16 elements after it, leading dimensions count fp16 elements.
struct ThPlanes { int a_planes, c_planes; long long a_lo_off, c_lo_off; int terms; };   // terms: 0 / 2 = fp16 hi + lo, 1 = one bf16 plane (no lo)
int th_gemm(int epi_nt, int M, int N, int K, const float* A0, const float* A1, int lda, const float* img0, const float* img1,
            float* C0, float* C1, int ldc, const float* bias0, const float* bias1, int relu, const float* mask, int nets,
            cudaStream_t st, uint32_t* bits_out0 = nullptr, const uint32_t* bits_in = nullptr,
            long long* timeline = nullptr, const ThPlanes* planes = nullptr);
bool th_dw_supported(const float* X, int ldx, const float* G, int ldg);
// x_planes / g_planes: that operand is a pair of fp16 planes (ThPlanes convention: a = X, c = G)
int th_gemm_dw(int M, int N, int K, const float* X, int ldx, const float* G, int ldg, float* C, int ldc, float* bias_out,
               int slices, long long slice_stride, int accumulate, cudaStream_t st, const ThPlanes* planes = nullptr);
int th_split_planes(const float* src, long long n, void* hi, void* lo, cudaStream_t st, int terms = 2);
int th_join_planes(const void* hi, const void* lo, long long n, float* dst, cudaStream_t st, int terms = 2);
// tensor-core gemm modes whose dW runs on the tf32 TMA kernel with the 512-row accumulation bound
inline bool tc_mode(int gemm_mode) { return gemm_mode == DPPO_GEMM_TF32X3 || gemm_mode == DPPO_GEMM_FP16X2; }
int tc_colsum(const float* G, int rows, int N, int ld, int slabs, float* out, long long slab_stride, int accumulate,
              cudaStream_t st);
int tc_transpose_weights(const NetShape& s, const float* params_pi, const float* head_pi, float* wt_pi,
                         const float* params_old, const float* head_old, float* wt_old, cudaStream_t st);

int launch_pack_heads(const NetShape& s, const float* params_pi, const float* params_old,
                      float* head_pi, float* head_old, cudaStream_t st);
// out_pi / out_old / d_out are [rows, s.Pld] with P valid columns
int launch_categorical(int op, const float* logits, const void* arg, int B, int ncat, uint64_t seed, uint64_t offset,
                       void* out, cudaStream_t st);
int launch_loss(const NetShape& s, const float* out_pi, const float* out_old, const float* act,
                const float* adv, const float* ret, int rows, int tile0, float inv_B, float clip,
                float entcoeff, float vcoeff, float* d_out, double* loss_part, cudaStream_t st);
// literal pairwise value loss (quirk D7): overwrites d_out[:,0] and the value slots of the loss partials of rows
// [row_off, row_off + rows) of a time-major batch of `groups` rollouts x T rows whose returns are ret_all[T*groups]
int launch_value_pairwise(const NetShape& s, const float* out_pi, const float* out_old, const float* ret_all, int rows,
                          int row_off, int T, int groups, int tile0, float inv_B, float clip, float vcoeff, float* d_out,
                          double* loss_part, cudaStream_t st);
// active[l] = number of split-K slices the dW launches of this update actually wrote (<= s.slices[l])
// fused: single-rank updates apply the TF-Adam step of every parameter inside finalize_kernel (no separate adam_kernel pass)
struct AdamFused { float* params; float* m; float* v; int step; float lr, b1, b2, eps; };
int launch_finalize(const NetShape& s, const int* active, const float* partial, const double* loss_part, int n_loss_tiles,
                    float entcoeff, float vcoeff, float* grad_out, float* losses_out, cudaStream_t st,
                    float* grad_out_alt = nullptr, const StepParams* sp = nullptr, const int* counter = nullptr,
                    float grad_scale = 1.0f, const AdamFused* fused = nullptr);
int launch_split_heads(const NetShape& s, const float* out, int rows, float* v_out, float* pd_out,
                       cudaStream_t st);
int launch_adam(float* params, const float* grad, float* m, float* v, int64_t n, int step, float lr,
                float b1, float b2, float eps, cudaStream_t st, const StepParams* sp = nullptr, const int* counter = nullptr);
double adam_lr_t(float lr, float b1, float b2, int step);
// end of a graph-replayed step: report the 8 loss scalars to losses_out[slot] and advance the step counter
int launch_step_advance(const float* grad_tail, float* losses_out, const StepParams* sp, int* counter, cudaStream_t st);
int launch_diag_gaussian(int op, const float* flat, const float* arg, int B, int A, uint64_t seed, uint64_t offset,
                         float* out, cudaStream_t st);
int launch_sample(const NetShape& s, const float* out, int rows, uint64_t seed, uint64_t offset,
                  float* action, float* v_out, cudaStream_t st);
int launch_gather_rows(const PermKey& key, int first, int rows, int obs_dim, int act_dim, const float* obs,
                       const float* act, const float* adv, const float* ret, float* obs_o, float* act_o, float* adv_o,
                       float* ret_o, const float* old_src, float* old_dst, int pld, cudaStream_t st,
                       const StepParams* sp = nullptr, const int* counter = nullptr);
int launch_perm_fill(const PermKey& key, int32_t* out, cudaStream_t st);

}  // namespace dppo
