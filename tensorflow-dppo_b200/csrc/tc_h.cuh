// Helpers shared by the persistent fp16-split tensor-core kernels (tc_gemm_h.cu: forward / dX, tc_dw_h.cu: dW).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace dppo {
namespace th {

using namespace tc;

constexpr int kMoverWarps = 8;
constexpr int kMoverThreads = kMoverWarps * 32;
constexpr int kMmaWarp = kMoverWarps;
constexpr int kProdWarp = kMoverWarps + 1;
constexpr int kEpiWarp0 = kMoverWarps + 2;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;      // 576
constexpr int kKB = 32;                                      // reduction elements per k-block
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D = f32, A = B = fp16 (format 0) or bf16 (format 1), both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_f16(int n, int m = BM, bool bf16 = false) {
    return (1u << 4) | (bf16 ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// shared-memory accesses by 32-bit shared-space address (a pointer derived from the aligned dynamic buffer is a GENERIC
// pointer to the compiler, which then emits generic LD/ST)
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst_smem), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// one barrier arrival per WARP: 256 per-thread arrivals on one mbarrier serialise in the LSU (measured: ~40 % of a k-block);
// the barriers the mover / converter warps signal are initialised with kMoverWarps
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// two fp32 -> packed fp16 hi pair and packed fp16 lo pair (lo pre-scaled by 2^11)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((a - hf.x) * kLoScale, (b - hf.y) * kLoScale);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster issue ONE M = 256 MMA; each holds its own 128 rows of A and half
// of B's rows, the accumulator rows of a CTA live in its own TMEM.  Only the leader (cluster rank 0) issues. ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {      // pairs with a remote release.cluster arrive
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
// TMA load issued by either CTA of a pair whose completion bytes are counted on an mbarrier of the LEADER (cluster address):
// the leader's MMA thread then waits on one barrier for both CTAs' operands, with no software relay in between
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst_smem), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once every MMA issued so far has completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// ONE-TERM mode (DPPO_GEMM_BF16, config 4): every operand is a single bf16 (round to nearest), one MMA per product
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// ---- TMA stores (shared -> global, bulk async-group completion) and split TMEM loads ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every committed store of this thread has finished READING shared memory (the global writes may still be in flight)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// two 32-column TMEM loads in flight, ONE wait, in one asm statement (so that no use of the registers can be scheduled
// between the loads and the wait)
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr_a, uint32_t (&a)[32], uint32_t taddr_b, uint32_t (&b)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=&r"(a[0]), "=&r"(a[1]), "=&r"(a[2]), "=&r"(a[3]), "=&r"(a[4]), "=&r"(a[5]), "=&r"(a[6]), "=&r"(a[7]), "=&r"(a[8]), "=&r"(a[9]), "=&r"(a[10]), "=&r"(a[11]), "=&r"(a[12]), "=&r"(a[13]), "=&r"(a[14]), "=&r"(a[15]), "=&r"(a[16]), "=&r"(a[17]), "=&r"(a[18]), "=&r"(a[19]), "=&r"(a[20]), "=&r"(a[21]), "=&r"(a[22]), "=&r"(a[23]), "=&r"(a[24]), "=&r"(a[25]), "=&r"(a[26]), "=&r"(a[27]), "=&r"(a[28]), "=&r"(a[29]), "=&r"(a[30]), "=&r"(a[31]),
          "=&r"(b[0]), "=&r"(b[1]), "=&r"(b[2]), "=&r"(b[3]), "=&r"(b[4]), "=&r"(b[5]), "=&r"(b[6]), "=&r"(b[7]), "=&r"(b[8]), "=&r"(b[9]), "=&r"(b[10]), "=&r"(b[11]), "=&r"(b[12]), "=&r"(b[13]), "=&r"(b[14]), "=&r"(b[15]), "=&r"(b[16]), "=&r"(b[17]), "=&r"(b[18]), "=&r"(b[19]), "=&r"(b[20]), "=&r"(b[21]), "=&r"(b[22]), "=&r"(b[23]), "=&r"(b[24]), "=&r"(b[25]), "=&r"(b[26]), "=&r"(b[27]), "=&r"(b[28]), "=&r"(b[29]), "=&r"(b[30]), "=&r"(b[31])
        : "r"(taddr_a), "r"(taddr_b) : "memory");
}

typedef CUresult (*HEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline HEncodeFn h_encode() {
    static HEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<HEncodeFn>(p);
    }
    return fn;
}
inline CUtensorMapL2promotion h_l2_promotion() {          // DPPO_H_L2PROMO = 0 | 128 | 256 (default 256: measured 1.5-2.5 % over 128 on the activation stream)
    static int v = -1;
    if (v < 0) { const char* e = getenv("DPPO_H_L2PROMO"); v = e ? atoi(e) : 256; }
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}
// row-major [rows, cols] fp32 matrix, box {32 cols (128 B), box_rows}, SWIZZLE_128B; out-of-range reads give zeros
inline bool h_make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows) {
    HEncodeFn enc = h_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, h_l2_promotion(),
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// row-major [rows, cols] fp16 matrix (one plane of an activation stored as hi | lo planes), box {box_cols, box_rows}
inline bool h_make_map16(CUtensorMap* map, const void* base, int rows, int cols, long long ld, int box_cols, int box_rows,
                         CUtensorMapSwizzle swizzle) {
    HEncodeFn enc = h_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2u};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, h_l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline int aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace th
}  // namespace dppo
