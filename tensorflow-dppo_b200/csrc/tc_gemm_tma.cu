// dW product on tensor cores with TMA-fed operands:  C[M,N] (+)= X[rows,M]^T . G[rows,N]   (split over rows)
//
// Replaces the weight-gradient part of TF's auto-diff (/root/reference/PPO.py:46) for the dense layers of
// /root/reference/Model.py:12-14.  Both operands are activations whose reduction dimension (rollout rows) is the
// slow axis, i.e. MN-major UMMA operands.  The register-staged movers of tc_gemm.cu cannot keep enough bytes in
// flight for this product (two activation operands, 48 KB per 32-row block), so here the TMA engine streams the
// fp32 tiles straight into the canonical SWIZZLE_128B_ATOM_32B layout, 4 stages deep:
//
//   warp 9   (1 thread)  : cp.async.bulk.tensor.2d  {32 features x 16 rows} boxes -> stage s   -> tma_full[s]
//   warps 0-7            : split IN PLACE: raw -> round-to-nearest tf32 `hi` (overwrites raw), `lo` to the
//                          companion buffer; column sums of G (the bias gradient) fall out of the same pass;
//                          fence.proxy.async -> conv_full[s]
//   warp 8   (1 thread)  : 6 tcgen05.mma kind::tf32 per stage (2 k-steps x {lo.hi, hi.lo -> correction
//                          accumulator; hi.hi -> main accumulator}), tcgen05.commit -> empty[s]
//   warps 0-7            : epilogue: TMEM -> smem staging -> coalesced partial-slice store (+= when accumulating)
//
// Rows per CTA are bounded by the split-K plan (<= 512) because TMEM accumulation truncates.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>

namespace dppo {
namespace tc {

constexpr int TBK = 16;                       // rollout rows per stage
constexpr int kConvWarps = 8;
constexpr int kConvThreads = kConvWarps * 32;
constexpr int kTThreads = kConvThreads + 64;  // + MMA warp + TMA warp

struct TmaArgs {
    float* C; float* bias_out;
    int M, N, K;                  // C is [M,N]; K = rollout rows
    int ldc;
    int k_per_slice;
    long long slice_stride;
    int accumulate;
    int vecC;
    int raw_hi;                   // leave the raw fp32 tile in place as the hi operand (a tf32 MMA ignores the low 13 bits)
};

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// BN = 256: one CTA per SM (TMEM full), 4 stages.  BN <= 128: sized so that TWO CTAs share an SM (256 TMEM columns and
// ~100 KB of shared memory each): while one CTA drains its accumulator the other keeps the tensor pipe busy.
template <int BN>
struct TmaPlan {
    static constexpr int kTStages = BN > 128 ? 4 : 3;
    static constexpr int kMinBlocks = BN > 128 ? 1 : 2;
    static constexpr int kARaw = TBK * BM * 4;            // 8 KB: 4 feature atoms x 16 rows x 128 B
    static constexpr int kBRaw = TBK * BN * 4;
    static constexpr int kRaw = kARaw + kBRaw;            // hi (in place) region of a stage
    static constexpr int kStage = 2 * kRaw;               // + lo region, same offsets
    static constexpr int kColsPerHalf = BN / 2 < 32 ? 32 : BN / 2;
    static constexpr int kStaging = kConvWarps * 32 * (kColsPerHalf + 4) * 4;
    static constexpr int kOperands = kTStages * kStage;
    static constexpr int kBytes = (kOperands > kStaging ? kOperands : kStaging) + 1024 + 512 + 4096;   // + align, barriers, colsum
};

__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// mapA3/mapB3 view the same matrices as [atom][row][32 features]: ONE box then brings all the full 32-feature atoms of a
// tile (the 2-D maps handle a ragged last atom, which needs the column bound).  use3 bit 0 / 1: 3-D map valid for A / B.
template <int BN>
__global__ void __launch_bounds__(kTThreads, TmaPlan<BN>::kMinBlocks)
tc_tn_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapA3, const __grid_constant__ CUtensorMap mapB3, const int use3,
                 const TmaArgs g) {
    using Plan = TmaPlan<BN>;
    constexpr int kTStages = Plan::kTStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kBarOff = (Plan::kOperands > Plan::kStaging ? Plan::kOperands : Plan::kStaging);
    uint64_t* tma_full = reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint64_t* conv_full = tma_full + kTStages;
    uint64_t* empty_bar = conv_full + kTStages;
    uint64_t* accum_bar = empty_bar + kTStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    float* colsum_smem = reinterpret_cast<float*>(smem + kBarOff + 512);    // [row owner][chunk] float4 partial column sums

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * BN;
    const int m0 = blockIdx.y * BM;
    const int slice = blockIdx.z;
    const int k_begin = slice * g.k_per_slice;
    const int k_end = min(g.K, k_begin + g.k_per_slice);
    const int nkb = max(0, (k_end - k_begin + TBK - 1) / TBK);

    if (tid == 0) {
        for (int s = 0; s < kTStages; ++s) {
            mbar_init(&tma_full[s], 1); mbar_init(&conv_full[s], kConvThreads); mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    constexpr uint32_t kTmemCols = 2 * (BN < 32 ? 32 : BN);
    if (warp == kConvWarps) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kConvWarps + 1) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kTStages;
                mbar_wait(&empty_bar[s], ((kb / kTStages) & 1) ^ 1);
                uint8_t* st = smem + s * Plan::kStage;
                mbar_arrive_expect_tx(&tma_full[s], Plan::kRaw);
                const int row = k_begin + kb * TBK;
                if ((use3 & 1) && m0 + BM <= (g.M / 32) * 32) tma_load_3d(st, &mapA3, 0, row, m0 / 32, &tma_full[s]);
                else {
#pragma unroll
                    for (int j = 0; j < BM / 32; ++j) tma_load_2d(st + j * (TBK * 128), &mapA, m0 + j * 32, row, &tma_full[s]);
                }
                if ((use3 & 2) && n0 + BN <= (g.N / 32) * 32) tma_load_3d(st + Plan::kARaw, &mapB3, 0, row, n0 / 32, &tma_full[s]);
                else {
#pragma unroll
                    for (int j = 0; j < BN / 32; ++j)
                        tma_load_2d(st + Plan::kARaw + j * (TBK * 128), &mapB, n0 + j * 32, row, &tma_full[s]);
                }
            }
        }
    } else if (warp == kConvWarps) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, true, true);
            constexpr uint32_t kLbo = TBK * 128, kSbo = 512;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kTStages;
                mbar_wait(&tma_full[s], (kb / kTStages) & 1);      // async-proxy writes of the raw tile
                mbar_wait(&conv_full[s], (kb / kTStages) & 1);     // hi rewritten + lo written by the converters
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * Plan::kStage);
                const uint32_t a_hi = st, b_hi = st + Plan::kARaw, a_lo = st + Plan::kRaw, b_lo = a_lo + Plan::kARaw;
#pragma unroll
                for (int k = 0; k < TBK / 8; ++k) {
                    const uint64_t da_hi = make_sdesc(a_hi + k * 1024, kLbo, kSbo, kLayoutSW128Base32);
                    const uint64_t da_lo = make_sdesc(a_lo + k * 1024, kLbo, kSbo, kLayoutSW128Base32);
                    const uint64_t db_hi = make_sdesc(b_hi + k * 1024, kLbo, kSbo, kLayoutSW128Base32);
                    const uint64_t db_lo = make_sdesc(b_lo + k * 1024, kLbo, kSbo, kLayoutSW128Base32);
                    umma_tf32(tmem_base + BN, da_lo, db_hi, idesc, (kb | k) != 0);
                    umma_tf32(tmem_base + BN, da_hi, db_lo, idesc, 1u);
                    umma_tf32(tmem_base, da_hi, db_hi, idesc, (kb | k) != 0);
                }
                umma_commit(&empty_bar[s]);
            }
            if (nkb > 0) umma_commit(accum_bar);
            tc_fence_before();
        }
    } else {
        // ================= converters (in-place hi/lo split + column sums of G), then epilogue =================
        // Chunk ownership is static: for operand X with Cx 16-B chunks per k-row, thread t always converts chunk
        // t % Cx (a fixed group of four features/columns) of rows t / Cx + i * (256 / Cx).
        constexpr int kChunksA = (BM / 32) * 8, kChunksB = (BN / 32) * 8;            // 32 and BN/4
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool want_bias = (g.bias_out != nullptr) && (blockIdx.y == 0);
        // compile-time trip counts: all LDS.128 of a stage issue back to back, then the cvt/STS stream
        auto convert = [&](uint8_t* base, auto cpr_tag, bool sum) {
            constexpr int CPR = decltype(cpr_tag)::value;
            constexpr int kStep = kConvThreads / CPR, kIt = TBK / kStep;
            const int cg = tid % CPR, atom = cg >> 3, lc = cg & 7, kr0 = tid / CPR;
            uint8_t* ab = base + atom * (TBK * 128);
            float4 v[kIt];
            uint32_t off[kIt];
#pragma unroll
            for (int i = 0; i < kIt; ++i) {
                const int kr = kr0 + i * kStep;
                off[i] = (uint32_t)kr * 128u + (uint32_t)((((lc >> 1) ^ (kr & 3)) << 5) | ((lc & 1) << 4));
                v[i] = *reinterpret_cast<const float4*>(ab + off[i]);
            }
#pragma unroll
            for (int i = 0; i < kIt; ++i) {
                uint4 h, l;
                if (g.raw_hi) {       // hi = the value with its low 13 mantissa bits cleared = what a tf32 MMA reads from raw fp32
                    h.x = __float_as_uint(v[i].x) & 0xFFFFE000u; h.y = __float_as_uint(v[i].y) & 0xFFFFE000u;
                    h.z = __float_as_uint(v[i].z) & 0xFFFFE000u; h.w = __float_as_uint(v[i].w) & 0xFFFFE000u;
                } else {
                    h.x = to_tf32(v[i].x); h.y = to_tf32(v[i].y); h.z = to_tf32(v[i].z); h.w = to_tf32(v[i].w);
                }
                l.x = to_tf32(v[i].x - __uint_as_float(h.x)); l.y = to_tf32(v[i].y - __uint_as_float(h.y));
                l.z = to_tf32(v[i].z - __uint_as_float(h.z)); l.w = to_tf32(v[i].w - __uint_as_float(h.w));
                if (!g.raw_hi) *reinterpret_cast<uint4*>(ab + off[i]) = h;
                *reinterpret_cast<uint4*>(ab + Plan::kRaw + off[i]) = l;
                if (sum) { csum.x += v[i].x; csum.y += v[i].y; csum.z += v[i].z; csum.w += v[i].w; }
            }
        };
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kTStages;
            mbar_wait(&tma_full[s], (kb / kTStages) & 1);
            uint8_t* st = smem + s * Plan::kStage;
            convert(st, std::integral_constant<int, kChunksA>{}, false);
            convert(st + Plan::kARaw, std::integral_constant<int, kChunksB>{}, want_bias);
            fence_proxy_async();
            mbar_arrive(&conv_full[s]);
        }
        constexpr int kRowStep = kConvThreads / kChunksB;
        // bias gradient: combine the kRowStep threads that own the same G columns (fixed order), one slice entry each
        if (want_bias) reinterpret_cast<float4*>(colsum_smem)[tid] = csum;    // slot = (row owner) * kChunksB + chunk
        asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");              // converters only
        if (want_bias && tid < kChunksB) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < kRowStep; ++r) {
                const float4 p = reinterpret_cast<const float4*>(colsum_smem)[r * kChunksB + tid];
                t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
            }
            const int n = n0 + (tid >> 3) * 32 + (tid & 7) * 4;
            float* bo = g.bias_out + (size_t)slice * g.slice_stride + n;
            const float vals[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (n + e < g.N) bo[e] = g.accumulate ? bo[e] + vals[e] : vals[e];
        }

        // ---------------- epilogue ----------------
        float* __restrict__ C = g.C + (size_t)slice * g.slice_stride;
        const int q = warp & 3, half = warp >> 2;
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
        constexpr int kColsPerHalf = Plan::kColsPerHalf;
        constexpr int kStgLd = kColsPerHalf + 4;
        float* stg = reinterpret_cast<float*>(smem) + (size_t)warp * (32 * kStgLd);
        const int col0 = half * kColsPerHalf;
#pragma unroll 1
        for (int cb = 0; cb < kColsPerHalf && col0 + cb < BN; cb += 32) {
            uint32_t v[32];
            if (nkb > 0) {
                uint32_t w2[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + cb), v);
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + col0 + cb), w2);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w2[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<uint4*>(stg + lane * kStgLd + cb + j4 * 4) = make_uint4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        }
        __syncwarp();
        if (col0 < BN) {
            constexpr int kLanesPerRow = kColsPerHalf / 4;
            constexpr int kRowsPerIter = 32 / kLanesPerRow;
            constexpr int kIters = 32 / kRowsPerIter;
            constexpr int kBatch = kIters < 8 ? kIters : 8;
            const int cl = (lane % kLanesPerRow) * 4;
            const int n = n0 + col0 + cl;
            const int nv = min(4, g.N - n);
#pragma unroll 1
            for (int it0 = 0; it0 < kIters; it0 += kBatch) {
                float4 aux[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int r = (it0 + u) * kRowsPerIter + lane / kLanesPerRow;
                    const int m = m0 + q * 32 + r;
                    aux[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g.accumulate && m < g.M && nv > 0) {
                        const float* sp = C + (size_t)m * g.ldc + n;
                        if (g.vecC && nv == 4) aux[u] = *reinterpret_cast<const float4*>(sp);
                        else {
                            aux[u].x = sp[0];
                            if (nv > 1) aux[u].y = sp[1];
                            if (nv > 2) aux[u].z = sp[2];
                            if (nv > 3) aux[u].w = sp[3];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int r = (it0 + u) * kRowsPerIter + lane / kLanesPerRow;
                    const int m = m0 + q * 32 + r;
                    if (m < g.M && nv > 0) {
                        const float4 t4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
                        const float x[4] = {t4.x + aux[u].x, t4.y + aux[u].y, t4.z + aux[u].z, t4.w + aux[u].w};
                        float* cp = C + (size_t)m * g.ldc + n;
                        if (g.vecC && nv == 4) *reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
                        else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (e < nv) cp[e] = x[e];
                        }
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == kConvWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- host side: tensor maps -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D map over a row-major [rows, cols] fp32 matrix, box = {32 cols, TBK rows}, SWIZZLE_128B_ATOM_32B
static int make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld) {
    EncodeTiledFn enc = get_encode();
    DPPO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)TBK};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DPPO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%d x %d] ld %d", (int)r, rows, cols, ld);
    return 0;
}

// 3-D view [atoms = cols/32][rows][32]: box {32, TBK, atoms_per_tile}; false when the driver rejects the strides
static bool make_map3(CUtensorMap* map, const float* base, int rows, int cols, int ld, int atoms_per_tile) {
    EncodeTiledFn enc = get_encode();
    if (!enc || cols / 32 < atoms_per_tile) return false;
    const cuuint64_t dims[3] = {32u, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), 128u};
    const cuuint32_t box[3] = {32u, (cuuint32_t)TBK, (cuuint32_t)atoms_per_tile};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
static int launch_tma(const CUtensorMap& mA, const CUtensorMap& mB, const float* A, int lda, const float* B, int ldb,
                      TmaArgs& g, int slices, cudaStream_t st) {
    using Plan = TmaPlan<BN>;
    static PerDeviceOnce configured;
    if (configured.need()) {
        DPPO_CHECK_CUDA(cudaFuncSetAttribute(tc_tn_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::kBytes));
    }
    static int want3 = -1;
    if (want3 < 0) { const char* e = getenv("DPPO_TC_TMA_2D"); want3 = (e && e[0] == '1') ? 0 : 1; }
    CUtensorMap mA3 = mA, mB3 = mB;
    int use3 = 0;
    if (want3) {
        if (make_map3(&mA3, A, g.K, g.M, lda, BM / 32)) use3 |= 1; else mA3 = mA;
        if (make_map3(&mB3, B, g.K, g.N, ldb, BN / 32)) use3 |= 2; else mB3 = mB;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, slices);
    ProfScope ps(KK_TC_GEMM, st);
    tc_tn_tma_kernel<BN><<<grid, kTThreads, Plan::kBytes, st>>>(mA, mB, mA3, mB3, use3, g);
    DPPO_LAUNCH_CHECK();
    return 0;
}

}  // namespace tc

// true when the TMA kernel can take this product (16-B aligned rows); otherwise use tc_gemm_tn_splitk
bool tc_tn_tma_supported(const float* A, int lda, const float* B, int ldb) {
    return (lda % 4 == 0) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
           ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
}

// C[s][M,N] (+)= A[rows,M]^T . B[rows,N] over the rows of slice s; bias_out[s][N] (+)= column sums of B (may be null)
int tc_gemm_tn_tma(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, float* bias_out,
                   int slices, long long slice_stride, int accumulate, cudaStream_t st) {
    CUtensorMap mA, mB;
    if (int r = tc::make_map(&mA, A, K, M, lda)) return r;
    if (int r = tc::make_map(&mB, B, K, N, ldb)) return r;
    tc::TmaArgs g{};
    g.C = C; g.bias_out = bias_out; g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.accumulate = accumulate;
    g.slice_stride = slice_stride;
    if (slices < 1) slices = 1;
    int kps = (K + slices - 1) / slices;
    kps = ((kps + tc::TBK - 1) / tc::TBK) * tc::TBK;
    if (kps == 0) kps = tc::TBK;
    g.k_per_slice = kps;
    g.vecC = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (slice_stride % 4 == 0);
    g.raw_hi = tc_raw_hi() ? 1 : 0;
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("DPPO_TC_TN_WIDE"); wide = (e && e[0] == '1') ? 1 : 0; }
    if (N > 128 && wide) return tc::launch_tma<256>(mA, mB, A, lda, B, ldb, g, slices, st);
    if (N > 64) return tc::launch_tma<128>(mA, mB, A, lda, B, ldb, g, slices, st);
    return tc::launch_tma<64>(mA, mB, A, lda, B, ldb, g, slices, st);
}

}  // namespace dppo
