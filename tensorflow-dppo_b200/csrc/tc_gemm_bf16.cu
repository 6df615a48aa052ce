// DPPO_GEMM_BF16: the dense layers on tcgen05 kind::f16 with bf16 operands and fp32 accumulation in TMEM.
//
// BASELINE.json configs[3] (Humanoid, 4x1024 MLP) names a bf16 tensor-core path with its own, looser tolerance:
// activations and weights stay fp32 in HBM (same buffers as the other paths) and are rounded to bf16 on their way
// into shared memory; one MMA per k-step (no hi/lo split), one accumulator.  Same products as tc_gemm.cu
// (/root/reference/Model.py:12-14 forward, auto-diff of PPO.py:46 backward):
//     K-major  (forward with transposed weights, dX):  A rows of 64 bf16 = 128 B, SWIZZLE_128B, UMMA_K = 16
//     MN-major (dW = X^T . dY):                       atoms of 64 features x 8 rows, SWIZZLE_128B, 32 rows per stage
// Warps 0-7 move operands (global fp32 -> registers -> cvt.rn.bf16x2 -> swizzled st.shared), warp 8 issues the MMAs
// and owns TMEM, warps 0-7 run the epilogue (TMEM -> smem staging -> coalesced global accesses).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda_bf16.h>

namespace dppo {
namespace bf {

using namespace dppo::tc;

constexpr int kLW = 8, kLT = kLW * 32, kThr = kLT + 32;
enum { EPI_NN = 0, EPI_NT = 1, EPI_TN = 2 };

struct Args {
    const float* A[2]; const float* B[2]; float* C[2]; const float* bias[2];
    const float* mask;
    int M, N, K, lda, ldb, ldc, tiles_n, nets, relu, accumulate, k_per_slice;
    long long slice_stride;
    int vecA, vecB, vecC;
};

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D = f32, A = B = bf16 (format 1), M = 128
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, bool mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((mn_major ? 1u : 0u) << 15) | ((mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);           // .x (low half) = lo
    return *reinterpret_cast<const uint32_t*>(&v);
}

// One operand tile per stage.  K-major: R rows x 64 bf16 (k-block = 64 elements).  MN-major: 32 k-rows x R features.
// Every loader thread owns 16-B smem chunks (8 bf16 = two float4 of fp32).
template <int R, bool MN>
struct Tile {
    static constexpr int kChunks = MN ? 32 * (R / 8) : R * 8;          // 16-B chunks per tile
    static constexpr int kPerThread = kChunks / kLT;
    static constexpr int kBytes = kChunks * 16;
    static constexpr int kKStep = MN ? 32 : 64;                        // reduction elements per stage
    static_assert(kChunks % kLT == 0, "tile must split evenly over the movers");
};

template <int R, bool MN>
__device__ __forceinline__ void load_tile(float4 (&reg)[Tile<R, MN>::kPerThread][2], const float* __restrict__ P, int ld, int r0,
                                          int rdim, int k0, int kend, int vec, int tid) {
#pragma unroll
    for (int i = 0; i < Tile<R, MN>::kPerThread; ++i) {
        const int q = tid + i * kLT;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* p; int left; bool ok;
        if (!MN) { const int r = r0 + (q >> 3), k = k0 + (q & 7) * 8; ok = r < rdim && k < kend; p = P + (size_t)r * ld + k; left = kend - k; }
        else { const int k = k0 + q / (R / 8), r = r0 + (q % (R / 8)) * 8; ok = k < kend && r < rdim; p = P + (size_t)k * ld + r; left = rdim - r; }
        if (ok) {
            if (vec && left >= 8) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (e < left) v[e] = __ldg(p + e);
            }
        }
        reg[i][0] = make_float4(v[0], v[1], v[2], v[3]); reg[i][1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

template <int R, bool MN>
__device__ __forceinline__ void store_tile(const float4 (&reg)[Tile<R, MN>::kPerThread][2], uint8_t* dst, int tid) {
#pragma unroll
    for (int i = 0; i < Tile<R, MN>::kPerThread; ++i) {
        const int q = tid + i * kLT;
        uint32_t off;
        if (!MN) { const int r = q >> 3, c = q & 7; off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }
        else {
            const int kr = q / (R / 8), cc = q % (R / 8), atom = cc >> 3, c = cc & 7;
            off = (uint32_t)atom * (32u * 128u) + (uint32_t)kr * 128u + (uint32_t)((c ^ (kr & 7)) << 4);
        }
        const float4 a = reg[i][0], b = reg[i][1];
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
    }
}

template <int BN, bool MN>
struct Plan {
    static constexpr int kA = Tile<BM, MN>::kBytes, kB = Tile<BN, MN>::kBytes, kStage = kA + kB;
    static constexpr int kStages = (192 * 1024) / kStage >= 4 ? 4 : (192 * 1024) / kStage;
    static constexpr int kColsPerHalf = BN / 2 < 32 ? 32 : BN / 2;
    static constexpr int kStaging = kLW * 32 * (kColsPerHalf + 4) * 4;
    static constexpr int kOperands = kStages * kStage;
    static constexpr int kBarOff = kOperands > kStaging ? kOperands : kStaging;
    static constexpr int kBytes = kBarOff + 1024 + 256;
};

template <int BN, bool MN, int EPI>
__global__ void __launch_bounds__(kThr, 1) bf16_gemm_kernel(const Args g) {
    using P = Plan<BN, MN>;
    constexpr int STAGES = P::kStages, KS = Tile<BM, MN>::kKStep;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P::kBarOff);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int net = blockIdx.x / g.tiles_n;
    const int n0 = (blockIdx.x % g.tiles_n) * BN, m0 = blockIdx.y * BM, slice = blockIdx.z;
    const int k_begin = slice * g.k_per_slice, k_end = min(g.K, k_begin + g.k_per_slice);
    const int nkb = max(0, (k_end - k_begin + KS - 1) / KS);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], kLT); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    constexpr uint32_t kCols = BN < 32 ? 32 : BN;
    if (warp == kLW) tmem_alloc(tmem_slot, kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kLW) {
        const float* __restrict__ A = g.A[net];
        const float* __restrict__ B = g.B[net];
        float4 ra[Tile<BM, MN>::kPerThread][2], rb[Tile<BN, MN>::kPerThread][2];
        if (nkb > 0) {
            load_tile<BM, MN>(ra, A, g.lda, m0, g.M, k_begin, k_end, g.vecA, tid);
            load_tile<BN, MN>(rb, B, g.ldb, n0, g.N, k_begin, k_end, g.vecB, tid);
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
            uint8_t* st = smem + s * P::kStage;
            store_tile<BM, MN>(ra, st, tid);
            store_tile<BN, MN>(rb, st + P::kA, tid);
            fence_proxy_async();
            mbar_arrive(&full_bar[s]);
            if (kb + 1 < nkb) {            // next block's loads are in flight while the tensor core works on this one
                const int k0 = k_begin + (kb + 1) * KS;
                load_tile<BM, MN>(ra, A, g.lda, m0, g.M, k0, k_end, g.vecA, tid);
                load_tile<BN, MN>(rb, B, g.ldb, n0, g.N, k0, k_end, g.vecB, tid);
            }
        }
        // ---------------- epilogue ----------------
        float* __restrict__ C = g.C[net];
        if (EPI == EPI_TN) C += (size_t)slice * g.slice_stride;
        const float* __restrict__ bias = g.bias[net];
        const int q = warp & 3, half = warp >> 2;
        if (nkb > 0) { mbar_wait(accum_bar, 0); tc_fence_after(); }
        constexpr int kColsPerHalf = P::kColsPerHalf, kStgLd = kColsPerHalf + 4;
        float* stg = reinterpret_cast<float*>(smem) + (size_t)warp * (32 * kStgLd);
        const int col0 = half * kColsPerHalf;
#pragma unroll 1
        for (int cb = 0; cb < kColsPerHalf && col0 + cb < BN; cb += 32) {
            uint32_t v[32];
            if (nkb > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + cb), v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<uint4*>(stg + lane * kStgLd + cb + j4 * 4) = make_uint4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        }
        __syncwarp();
        if (col0 < BN) {
            constexpr int kLanesPerRow = kColsPerHalf / 4, kRowsPerIter = 32 / kLanesPerRow, kIters = 32 / kRowsPerIter;
            constexpr int kBatch = kIters < 8 ? kIters : 8;
            const int cl = (lane % kLanesPerRow) * 4, n = n0 + col0 + cl, nv = min(4, g.N - n);
            float bv[4] = {0.f, 0.f, 0.f, 0.f};
            if (EPI == EPI_NN && bias) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (e < nv) bv[e] = __ldg(bias + n + e);
            }
#pragma unroll 1
            for (int it0 = 0; it0 < kIters; it0 += kBatch) {
                float4 aux[kBatch];
                if (EPI != EPI_NN) {
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int m = m0 + q * 32 + (it0 + u) * kRowsPerIter + lane / kLanesPerRow;
                        aux[u] = (EPI == EPI_NT) ? make_float4(1.f, 1.f, 1.f, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float* src = (EPI == EPI_NT) ? g.mask : (g.accumulate ? C : nullptr);
                        if (src && m < g.M && nv > 0) {
                            const float* sp = src + (size_t)m * g.ldc + n;
                            if (g.vecC && nv == 4) aux[u] = *reinterpret_cast<const float4*>(sp);
                            else { aux[u].x = sp[0]; if (nv > 1) aux[u].y = sp[1]; if (nv > 2) aux[u].z = sp[2]; if (nv > 3) aux[u].w = sp[3]; }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int r = (it0 + u) * kRowsPerIter + lane / kLanesPerRow, m = m0 + q * 32 + r;
                    if (m < g.M && nv > 0) {
                        const float4 t4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
                        float x[4] = {t4.x, t4.y, t4.z, t4.w};
                        float* cp = C + (size_t)m * g.ldc + n;
                        if (EPI == EPI_NN) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) { const float y = x[e] + bv[e]; x[e] = g.relu ? fmaxf(y, 0.f) : y; }
                        } else if (EPI == EPI_NT) {
                            x[0] = aux[u].x > 0.f ? x[0] : 0.f; x[1] = aux[u].y > 0.f ? x[1] : 0.f;
                            x[2] = aux[u].z > 0.f ? x[2] : 0.f; x[3] = aux[u].w > 0.f ? x[3] : 0.f;
                        } else { x[0] += aux[u].x; x[1] += aux[u].y; x[2] += aux[u].z; x[3] += aux[u].w; }
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
    } else if (lane == 0) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = make_idesc_bf16(BN, MN);
        constexpr uint32_t kLbo = MN ? 32u * 128u : 0u, kSbo = 1024u;
        constexpr uint32_t kStep = MN ? 16u * 128u : 32u;                 // one UMMA_K = 16 elements
        constexpr int kMmas = MN ? 2 : 4;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&full_bar[s], (kb / STAGES) & 1);
            tc_fence_after();
            const uint32_t a = smem_u32(smem + s * P::kStage), b = a + P::kA;
#pragma unroll
            for (int k = 0; k < kMmas; ++k)
                umma_bf16(tmem_base, make_sdesc(a + k * kStep, kLbo, kSbo, kLayoutSW128), make_sdesc(b + k * kStep, kLbo, kSbo, kLayoutSW128),
                          idesc, (kb | k) != 0);
            umma_commit(&empty_bar[s]);
        }
        if (nkb > 0) umma_commit(accum_bar);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == kLW) { tc_fence_after(); tmem_dealloc(tmem_base, kCols); }
}

static inline int aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static void vec_flags(Args& g) {
    g.vecA = (g.lda % 4 == 0); g.vecB = (g.ldb % 4 == 0); g.vecC = (g.ldc % 4 == 0);
    for (int i = 0; i < g.nets; ++i) { g.vecA &= aligned16(g.A[i]); g.vecB &= aligned16(g.B[i]); g.vecC &= aligned16(g.C[i]); }
}

template <int BN, bool MN, int EPI>
static int launch(Args& g, int slices, cudaStream_t st) {
    using P = Plan<BN, MN>;
    g.tiles_n = (g.N + BN - 1) / BN;
    const int tiles_m = (g.M + BM - 1) / BM;
    DPPO_REQUIRE(tiles_m <= 65535, "bf16 gemm: M=%d too large for one launch", g.M);
    static PerDeviceOnce configured;
    if (configured.need()) {
        DPPO_CHECK_CUDA(cudaFuncSetAttribute(bf16_gemm_kernel<BN, MN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kBytes));
    }
    dim3 grid(g.tiles_n * g.nets, tiles_m, slices);
    ProfScope ps(KK_TC_GEMM, st);
    bf16_gemm_kernel<BN, MN, EPI><<<grid, kThr, P::kBytes, st>>>(g);
    DPPO_LAUNCH_CHECK();
    return 0;
}
template <bool MN, int EPI>
static int dispatch(Args& g, int slices, cudaStream_t st) {
    if (g.N > 128) return launch<256, MN, EPI>(g, slices, st);
    if (g.N > 64) return launch<128, MN, EPI>(g, slices, st);
    return launch<64, MN, EPI>(g, slices, st);
}

}  // namespace bf

// C[M,N] = epi(A[M,K] . Bt[N,K]^T), bf16 operands.  nets = 1 or 2 products sharing shapes.
int bf16_gemm_kk(int epi_nt, int M, int N, int K, const float* A0, const float* A1, int lda, const float* B0, const float* B1,
                 int ldb, float* C0, float* C1, int ldc, const float* bias0, const float* bias1, int relu, const float* mask,
                 int nets, cudaStream_t st) {
    if (M == 0) return 0;
    bf::Args g{};
    g.A[0] = A0; g.A[1] = A1; g.B[0] = B0; g.B[1] = B1; g.C[0] = C0; g.C[1] = C1; g.bias[0] = bias0; g.bias[1] = bias1;
    g.mask = mask; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.nets = nets; g.relu = relu;
    g.k_per_slice = ((K + 63) / 64) * 64;
    bf::vec_flags(g);
    if (epi_nt) return bf::dispatch<false, bf::EPI_NT>(g, 1, st);
    return bf::dispatch<false, bf::EPI_NN>(g, 1, st);
}

// C[s][M,N] (+)= A[k,M]^T . B[k,N] over the rows of slice s, bf16 operands
int bf16_gemm_tn_splitk(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int slices,
                        long long slice_stride, int accumulate, cudaStream_t st) {
    bf::Args g{};
    g.A[0] = A; g.B[0] = B; g.C[0] = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.nets = 1;
    g.accumulate = accumulate; g.slice_stride = slice_stride;
    if (slices < 1) slices = 1;
    int kps = (K + slices - 1) / slices;
    kps = ((kps + 31) / 32) * 32;
    if (kps == 0) kps = 32;
    g.k_per_slice = kps;
    bf::vec_flags(g);
    if (g.slice_stride % 4 != 0) g.vecC = 0;
    return bf::dispatch<true, bf::EPI_TN>(g, slices, st);
}

}  // namespace dppo
