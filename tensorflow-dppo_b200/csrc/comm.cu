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

namespace dppo {

typedef struct { char internal[DPPO_NCCL_UNIQUE_ID_BYTES]; } nccl_unique_id_t;
typedef void* nccl_comm_t;
// ncclDataType_t / ncclRedOp_t values are ABI-stable since NCCL 2.10
constexpr int kNcclFloat32 = 7;
constexpr int kNcclAvg = 4;

struct NcclApi {
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
    return 0;
}

int comm_destroy(dppo_ctx* ctx) {
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
