// K6: the chief's gradient average (/root/reference/PPO.py:55-65, fed by Chief.py:54-64) as ONE
// ncclAllReduce(ncclAvg) on the flat gradient buffer, one rank per GPU over NVLink 5 / NVSwitch.
//
// NCCL is resolved at run time with dlopen so that libdppo_b200.so loads (and exports every
// symbol) on a box without a GPU; the first communicator call fails loudly if it is missing.
// When PyTorch is already in the process its bundled libnccl.so.2 is the one that is found
// (same SONAME), so the communicator shares NCCL's process-wide state with torch.distributed.
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

namespace dppo {

typedef struct { char internal[DPPO_NCCL_UNIQUE_ID_BYTES]; } nccl_unique_id_t;
typedef void* nccl_comm_t;
// ncclDataType_t / ncclRedOp_t values are ABI-stable since NCCL 2.10
constexpr int kNcclFloat32 = 7;
constexpr int kNcclAvg = 4;

struct NcclApi {
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id_t*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id_t, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    DPPO_REQUIRE(h != nullptr, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (int (*)(nccl_unique_id_t*))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_unique_id_t, int))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (int (*)(int*))dlsym(h, "ncclGetVersion");
    a.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    DPPO_REQUIRE(a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString,
                 "libnccl is missing a required symbol");
    g_nccl = a;
    return 0;
}

#define DPPO_CHECK_NCCL(expr)                                                              \
    do {                                                                                   \
        int _r = (expr);                                                                   \
        if (_r != 0) {                                                                     \
            dppo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                            g_nccl.GetErrorString(_r));                                    \
            return -3;                                                                     \
        }                                                                                  \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// NVLink / NVSwitch peer-memory exchange: the gradient mean of PPO.py:55-65 fused with the Adam apply of
// PPO.py:53 in ONE kernel.  Every rank owns a double-buffered exchange buffer (cudaMalloc, shared with
// cudaIpc); finalize_kernel writes the rank's gradient there; peer_allreduce_adam_kernel then
//   (1) signals arrival in every peer's flag array and waits for all peers (release/acquire at system scope),
//   (2) reads the W buffers over NVLink with 16-byte peer loads, sums them in rank order (bit-identical on
//       every rank), scales by 1/W,
//   (3) applies the TF-form Adam step to the local parameters and stores the averaged gradient + losses.
// One-shot (every rank reads everything): at 0.7 MB the exchange is latency-bound, so this beats a ring.
// ---------------------------------------------------------------------------------------------------------
static int peer_teardown(dppo_ctx* ctx) {
    for (int p = 0; p < ctx->world && p < 16; ++p) {
        if (p != ctx->rank && ctx->peer_mapped[p]) cudaIpcCloseMemHandle(ctx->peer_mapped[p]);
        if (p != ctx->rank && ctx->flags_mapped[p]) cudaIpcCloseMemHandle(ctx->flags_mapped[p]);
        ctx->peer_mapped[p] = nullptr; ctx->flags_mapped[p] = nullptr;
    }
    if (ctx->peer_self) cudaFree(ctx->peer_self);
    if (ctx->flags_self) cudaFree(ctx->flags_self);
    if (ctx->peer_bufs_dev) cudaFree(ctx->peer_bufs_dev);
    if (ctx->flags_dev) cudaFree(ctx->flags_dev);
    ctx->peer_self = nullptr; ctx->flags_self = nullptr; ctx->peer_bufs_dev = nullptr; ctx->flags_dev = nullptr;
    return 0;
}

struct PeerHandles { cudaIpcMemHandle_t buf, flags; };
static_assert(sizeof(PeerHandles) == DPPO_PEER_HANDLE_BYTES, "DPPO_PEER_HANDLE_BYTES must match the two IPC handles");

// step 1: this rank's exchange buffers + their IPC handles
static int peer_alloc(dppo_ctx* ctx, PeerHandles* mine) {
    const int W = ctx->world;
    ctx->peer_len = ((ctx->shape.n_params + kLossSlots + 3) / 4) * 4;
    DPPO_CHECK_CUDA(cudaMalloc((void**)&ctx->peer_self, (size_t)2 * ctx->peer_len * sizeof(float)));
    DPPO_CHECK_CUDA(cudaMalloc((void**)&ctx->flags_self, (size_t)(W + 1) * sizeof(unsigned int)));
    DPPO_CHECK_CUDA(cudaMemset(ctx->flags_self, 0, (size_t)(W + 1) * sizeof(unsigned int)));
    DPPO_CHECK_CUDA(cudaMemset(ctx->peer_self, 0, (size_t)2 * ctx->peer_len * sizeof(float)));
    DPPO_CHECK_CUDA(cudaDeviceSynchronize());          // the zeroed flags must be in memory before a peer maps them
    DPPO_CHECK_CUDA(cudaIpcGetMemHandle(&mine->buf, ctx->peer_self));
    DPPO_CHECK_CUDA(cudaIpcGetMemHandle(&mine->flags, ctx->flags_self));
    return 0;
}

// step 2: map every peer's buffers (all[] in rank order; all[rank] is this rank's own entry and is not opened)
static int peer_map(dppo_ctx* ctx, const PeerHandles* all) {
    const int W = ctx->world, R = ctx->rank;
    float* bufs[16]; unsigned int* flags[16];
    for (int p = 0; p < W; ++p) {
        if (p == R) { bufs[p] = ctx->peer_self; flags[p] = ctx->flags_self; continue; }
        DPPO_CHECK_CUDA(cudaIpcOpenMemHandle(&ctx->peer_mapped[p], all[p].buf, cudaIpcMemLazyEnablePeerAccess));
        DPPO_CHECK_CUDA(cudaIpcOpenMemHandle(&ctx->flags_mapped[p], all[p].flags, cudaIpcMemLazyEnablePeerAccess));
        bufs[p] = (float*)ctx->peer_mapped[p]; flags[p] = (unsigned int*)ctx->flags_mapped[p];
    }
    DPPO_CHECK_CUDA(cudaMalloc((void**)&ctx->peer_bufs_dev, sizeof(float*) * W));
    DPPO_CHECK_CUDA(cudaMalloc((void**)&ctx->flags_dev, sizeof(unsigned int*) * W));
    DPPO_CHECK_CUDA(cudaMemcpy(ctx->peer_bufs_dev, bufs, sizeof(float*) * W, cudaMemcpyHostToDevice));
    DPPO_CHECK_CUDA(cudaMemcpy(ctx->flags_dev, flags, sizeof(unsigned int*) * W, cudaMemcpyHostToDevice));
    ctx->epoch = 0;
    return 0;
}

// both steps with the handles all-gathered through the NCCL communicator we already have
static int peer_setup(dppo_ctx* ctx) {
    const int W = ctx->world, R = ctx->rank;
    PeerHandles mine;
    if (int r = peer_alloc(ctx, &mine)) return r;
    PeerHandles* d_all = nullptr;
    DPPO_CHECK_CUDA(cudaMalloc((void**)&d_all, sizeof(PeerHandles) * W));
    DPPO_CHECK_CUDA(cudaMemcpy(d_all + R, &mine, sizeof(PeerHandles), cudaMemcpyHostToDevice));
    cudaStream_t s0 = 0;
    const int rc = g_nccl.AllGather(d_all + R, d_all, sizeof(PeerHandles), /*ncclInt8*/ 0, (nccl_comm_t)ctx->nccl_comm, s0);
    if (rc != 0) { cudaFree(d_all); DPPO_REQUIRE(false, "ncclAllGather of IPC handles: %s", g_nccl.GetErrorString(rc)); }
    PeerHandles all[16];
    DPPO_CHECK_CUDA(cudaStreamSynchronize(s0));
    DPPO_CHECK_CUDA(cudaMemcpy(all, d_all, sizeof(PeerHandles) * W, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    return peer_map(ctx, all);
}

// ---- the same exchange without NCCL: the host moves the handle blobs (file, pipe, MPI, torch.distributed ...) ----
int comm_peer_export(dppo_ctx* ctx, int rank, int world, void* handle_out) {
    DPPO_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world, "dppo_comm_peer_export: rank %d of %d (2..16 ranks)", rank, world);
    DPPO_REQUIRE(ctx->nccl_comm == nullptr && ctx->peer_self == nullptr, "dppo_comm_peer_export: a communicator already exists");
    ctx->rank = rank; ctx->world = world;
    PeerHandles mine;
    if (int r = peer_alloc(ctx, &mine)) { peer_teardown(ctx); ctx->world = 1; return r; }
    memcpy(handle_out, &mine, sizeof(mine));
    return 0;
}

int comm_peer_import(dppo_ctx* ctx, const void* all_handles) {
    DPPO_REQUIRE(ctx->peer_self != nullptr && ctx->peer_bufs_dev == nullptr, "dppo_comm_peer_import: call dppo_comm_peer_export first");
    PeerHandles all[16];
    memcpy(all, all_handles, sizeof(PeerHandles) * ctx->world);
    if (int r = peer_map(ctx, all)) { peer_teardown(ctx); ctx->world = 1; return r; }
    return 0;
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
peer_allreduce_adam_kernel(float* const* __restrict__ bufs, unsigned int* const* __restrict__ flags, int rank, int world,
                           unsigned int epoch, long long buf_off, long long n_params, long long n_valid, long long n_total,
                           float* __restrict__ params,
                           float* __restrict__ m, float* __restrict__ v, float lr_t, float b1, float b2, float eps,
                           float* __restrict__ grad_out, const StepParams* __restrict__ sp, const int* __restrict__ counter) {
    if (sp) {                                                          // graph replay: this step's epoch / buffer half / step size
        const StepParams& q = sp[*counter];
        epoch = q.peer_epoch; buf_off = (long long)q.peer_sel * n_total; lr_t = q.lr_t;
    }
    unsigned int* my_flags = flags[rank];
    unsigned int* go = my_flags + world;
    if (blockIdx.x == 0) {
        if (threadIdx.x < world) {
            __threadfence_system();                                    // this rank's finalize writes are visible system-wide
            st_release_sys(flags[threadIdx.x] + rank, epoch);          // "rank arrived at `epoch`" into every peer (and self)
            while (ld_acquire_sys(my_flags + threadIdx.x) < epoch) { } // peer `threadIdx.x` arrived
        }
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(go, epoch);
    }
    if (threadIdx.x == 0) while (ld_acquire_sys(go) < epoch) { }
    __syncthreads();
    const float inv = 1.0f / (float)world;
    const long long n4 = n_total / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < world; ++p) {                              // rank order: identical sums on every rank
            const float4 x = ld_peer_f4(bufs[p] + buf_off + i * 4);
            s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
        }
        float g[4] = {s.x * inv, s.y * inv, s.z * inv, s.w * inv};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long j = i * 4 + e;
            if (j < n_valid) grad_out[j] = g[e];
            if (m != nullptr && j < n_params) {
                const float mi = b1 * m[j] + (1.f - b1) * g[e];
                const float vi = b2 * v[j] + (1.f - b2) * g[e] * g[e];
                m[j] = mi; v[j] = vi;
                params[j] = params[j] - lr_t * mi / (sqrtf(vi) + eps);
            }
        }
    }
}

bool peer_available(const dppo_ctx* ctx) { return ctx->peer_bufs_dev != nullptr; }
float* peer_exchange_buffer(dppo_ctx* ctx) {         // where finalize_kernel writes this update's gradient
    return ctx->peer_self + (size_t)((ctx->epoch + 1) & 1) * ctx->peer_len;
}

// grad (already in this rank's exchange buffer for epoch+1) -> mean over ranks in grad_out (+ Adam when m, v given)
float* peer_exchange_half(dppo_ctx* ctx, int sel) { return ctx->peer_self + (size_t)sel * ctx->peer_len; }

int peer_allreduce_adam(dppo_ctx* ctx, float* params, float* m, float* v, int step, float lr, float b1, float b2, float eps,
                        float* grad_out, cudaStream_t st, const StepParams* sp, const int* counter) {
    DPPO_REQUIRE(peer_available(ctx), "peer exchange is not set up");
    if (!sp) ctx->epoch += 1;                      // graph replay: the host advanced the epoch when it filled the step records
    const long long n_total = ctx->peer_len;
    double lr_t = 0.0;
    if (m) {
        DPPO_REQUIRE(step >= 1, "dppo_update: step=%d must be >= 1 (t after increment)", step);
        lr_t = (double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step));
    }
    int grid = (int)((n_total / 4 + 255) / 256);
    if (grid > 64) grid = 64;                      // all CTAs must be co-resident (they spin on the go word)
    if (grid < 1) grid = 1;
    ProfScope ps(KK_ALLREDUCE, st);
    peer_allreduce_adam_kernel<<<grid, 256, 0, st>>>(ctx->peer_bufs_dev, ctx->flags_dev, ctx->rank, ctx->world, ctx->epoch,
                                                     (long long)(ctx->epoch & 1) * ctx->peer_len, ctx->shape.n_params,
                                                     ctx->shape.n_params + kLossSlots, n_total,
                                                     params, m, v, (float)lr_t, b1, b2, eps, grad_out, sp, counter);
    DPPO_LAUNCH_CHECK();
    return 0;
}

int comm_unique_id(void* out) {
    if (int r = load_nccl()) return r;
    nccl_unique_id_t id;
    DPPO_CHECK_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return 0;
}

int comm_init(dppo_ctx* ctx, int rank, int world, const void* id_bytes) {
    DPPO_REQUIRE(world >= 1 && rank >= 0 && rank < world, "dppo_comm_init: rank %d of %d", rank, world);
    DPPO_REQUIRE(ctx->nccl_comm == nullptr, "dppo_comm_init: communicator already initialised");
    if (int r = load_nccl()) return r;
    nccl_unique_id_t id;
    memcpy(&id, id_bytes, sizeof(id));
    DPPO_CHECK_CUDA(cudaSetDevice(ctx->cfg.device));
    nccl_comm_t comm = nullptr;
    DPPO_CHECK_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm; ctx->rank = rank; ctx->world = world;
    if (world > 1 && world <= 16 && g_nccl.AllGather && !getenv("DPPO_NCCL_ONLY")) {
        if (int r = peer_setup(ctx)) {
            // peer mapping is an optimisation: without it the update falls back to ncclAllReduce + adam_kernel
            fprintf(stderr, "[dppo_b200] NVLink peer exchange unavailable (%s); using NCCL allreduce\n", dppo_last_error());
            peer_teardown(ctx);
        }
    }
    return 0;
}

int comm_destroy(dppo_ctx* ctx) {
    peer_teardown(ctx);
    if (ctx->nccl_comm) {
        DPPO_CHECK_NCCL(g_nccl.CommDestroy((nccl_comm_t)ctx->nccl_comm));
        ctx->nccl_comm = nullptr;
    }
    return 0;
}

int allreduce_avg(dppo_ctx* ctx, float* buf, int64_t n, cudaStream_t st) {
    DPPO_REQUIRE(ctx->nccl_comm != nullptr, "dppo_allreduce_avg: no communicator (call dppo_comm_init first)");
    if (n == 0) return 0;
    ProfScope ps(KK_ALLREDUCE, st);
    DPPO_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, kNcclFloat32, kNcclAvg, (nccl_comm_t)ctx->nccl_comm, st));
    return 0;
}

}  // namespace dppo
