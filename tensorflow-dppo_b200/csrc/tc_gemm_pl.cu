// EXPERIMENT (not on the update path; dppo_dbg_gemm modes 80-83, tests/test_gpu_parity.py::test_tc_planes_gemm):
// TMA-fed K-major products on pre-split operand PLANES (sm_100a, tcgen05 kind::tf32, 3-term fp32-accurate split).
//
// tc_gemm.cu moves the activation operand global -> registers -> RN hi/lo split -> swizzled st.shared with 8 SIMT warps
// per CTA.  The question this kernel answers: is that conversion what limits the tiles narrower than 256 columns?  Here
// the activation arrives as two fp32 planes (hi = RN-tf32(x), lo = RN-tf32(x - hi)) that a PRODUCING epilogue wrote, so
// one thread feeds the operand stages with cp.async.bulk.tensor (SWIZZLE_128B boxes {32 k, 128 rows}) next to the
// weight-image bulk copy and no thread touches the data between HBM and the MMA; the CTA is persistent (grid = #SMs)
// over 128x128 tiles, TMEM holds TWO accumulator sets (main + correction, 2 x 256 columns) and four epilogue warps
// drain set i while the MMA thread fills set i^1.
//   warp 0   : producer (one thread): per k-block 2 TMA boxes (A hi, A lo) + 1 bulk copy (B image hi|lo), 3 stages x 64 KB
//   warp 1   : TMEM owner + MMA issuer (one thread): 12 MMAs per k-block, tcgen05.commit frees the stage
//   warps 2-5: epilogue: tcgen05.ld -> smem staging (32-column slices) -> bias/ReLU | ReLU-mask -> coalesced stores of the
//              fp32 result and/or its hi/lo planes for the next consumer
// MEASURED ANSWER: no.  Long-K steady state 160 TFLOP/s (algorithmic) vs 244 for the 256-wide mover kernel; forward L2
// 111 us = 111 us.  A kind::tf32 MMA covers K = 8 per instruction with a fixed ~30-cycle cost that a 128x128x8
// instruction amortises far worse than a 128x256x8 one, so 256-wide tiles stay the product path.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace dppo {
namespace tc {

constexpr int kLBN = 128;                       // tile width
constexpr int kLBK = 32;                        // k-block: 32 fp32 = one 128-B swizzle row
constexpr int kLStages = 3;
constexpr int kLATile = BM * 128, kLBTile = kLBN * 128;
constexpr int kLStage = 2 * (kLATile + kLBTile);            // A hi, A lo, B hi, B lo: 64 KB
constexpr int kLStgLd = 36;                                  // staging row stride (floats): 32 columns + pad
constexpr int kLStaging = 4 * 32 * kLStgLd * 4;              // 4 epilogue warps x 32 rows
constexpr int kLThreads = 6 * 32;
constexpr int kLSmemBytes = kLStages * kLStage + kLStaging + 1024 + 256;
constexpr int kLImgReplicas = 8;

struct PlArgs {
    const float* Bimg[2];          // weight images (bn = 128, bk = 32), see tc_pack_images
    float* C[2];                   // fp32 result (nullable)
    float* Chi[2]; float* Clo[2];  // result planes (nullable)
    const float* bias[2];
    const float* mask;             // NT epilogue: C *= (mask > 0), [M,N] with row stride ldc
    int M, N, K, ldc;
    int tiles_n, nets, relu;
    long long img_replica_bytes;
};

__device__ __forceinline__ void pl_tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

template <int EPI>      // 0: C = relu?(acc + bias)   1: C = acc * (mask > 0)
__global__ void __launch_bounds__(kLThreads, 1)
tc_pl_kernel(const __grid_constant__ CUtensorMap hi0, const __grid_constant__ CUtensorMap lo0,
             const __grid_constant__ CUtensorMap hi1, const __grid_constant__ CUtensorMap lo1, const PlArgs g,
             int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* staging = reinterpret_cast<float*>(smem + kLStages * kLStage);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kLStages * kLStage + kLStaging);
    uint64_t* empty_bar = full_bar + kLStages;
    uint64_t* acc_full = empty_bar + kLStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = (g.K + kLBK - 1) / kLBK;
    const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per_cta;
    const int t_end = min(total_tiles, t_begin + per_cta);
    const int my_tiles = max(0, t_end - t_begin);
    const long long total_it = (long long)my_tiles * nkb;

    if (tid == 0) {
        for (int s = 0; s < kLStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile t -> (m_tile, net, n_tile): the n-tiles of one (rows, net) are consecutive, so the second re-reads A from L2
    auto decode = [&](int t, int& mt, int& net, int& nt) {
        nt = t % g.tiles_n; t /= g.tiles_n;
        net = t % g.nets; mt = t / g.nets;
    };

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            constexpr uint32_t kBBytes = 2 * kLBTile;
            constexpr uint32_t kTx = 2 * kLATile + kBBytes;
            for (long long it = 0; it < total_it; ++it) {
                int mt, net, nt;
                decode(t_begin + (int)(it / nkb), mt, net, nt);
                const int kb = (int)(it % nkb);
                const int s = (int)(it % kLStages);
                mbar_wait(&empty_bar[s], (uint32_t)(((it / kLStages) & 1) ^ 1));
                uint8_t* st = smem + s * kLStage;
                mbar_arrive_expect_tx(&full_bar[s], kTx);
                pl_tma_load_2d(st, net ? &hi1 : &hi0, kb * kLBK, mt * BM, &full_bar[s]);
                pl_tma_load_2d(st + kLATile, net ? &lo1 : &lo0, kb * kLBK, mt * BM, &full_bar[s]);
                const uint8_t* img = reinterpret_cast<const uint8_t*>(g.Bimg[net]) +
                                     (size_t)(blockIdx.x % kLImgReplicas) * g.img_replica_bytes +
                                     ((size_t)nt * nkb + (size_t)kb) * kBBytes;
                bulk_copy_g2s(st + 2 * kLATile, img, kBBytes, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(kLBN, false, false);
            long long it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t acc = tmem_base + (uint32_t)((i & 1) * 2 * kLBN);
                mbar_wait(&acc_empty[i & 1], (uint32_t)(((i >> 1) & 1) ^ 1));      // epilogue has drained this set
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = (int)(it % kLStages);
                    mbar_wait(&full_bar[s], (uint32_t)((it / kLStages) & 1));
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * kLStage);
                    const uint32_t a_hi = st, a_lo = st + kLATile, b_hi = st + 2 * kLATile, b_lo = b_hi + kLBTile;
#pragma unroll
                    for (int k = 0; k < kLBK / 8; ++k) {
                        const uint64_t da_hi = make_sdesc(a_hi + k * 32, 0, 1024, kLayoutSW128);
                        const uint64_t da_lo = make_sdesc(a_lo + k * 32, 0, 1024, kLayoutSW128);
                        const uint64_t db_hi = make_sdesc(b_hi + k * 32, 0, 1024, kLayoutSW128);
                        const uint64_t db_lo = make_sdesc(b_lo + k * 32, 0, 1024, kLayoutSW128);
                        umma_tf32(acc + kLBN, da_lo, db_hi, idesc, (kb | k) != 0);
                        umma_tf32(acc + kLBN, da_hi, db_lo, idesc, 1u);
                        umma_tf32(acc, da_hi, db_hi, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&acc_full[i & 1]);
            }
            tc_fence_before();
        }
    } else {
        // ---------------- epilogue warps ----------------
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        float* stg = staging + (size_t)(warp - 2) * (32 * kLStgLd);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;          // write-out: 8 lanes x float4 per row, 4 rows per pass
        for (int i = 0; i < my_tiles; ++i) {
            int mt, net, nt;
            decode(t_begin + i, mt, net, nt);
            const uint32_t acc = tmem_base + (uint32_t)((i & 1) * 2 * kLBN) + ((uint32_t)(q * 32) << 16);
            float* __restrict__ C = g.C[net];
            float* __restrict__ Chi = g.Chi[net];
            float* __restrict__ Clo = g.Clo[net];
            const float* __restrict__ bias = g.bias[net];
            mbar_wait(&acc_full[i & 1], (uint32_t)((i >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < kLBN; cb += 32) {
                {
                    uint32_t v[32], w2[32];
                    tmem_ld32(acc + (uint32_t)cb, v);
                    tmem_ld32(acc + (uint32_t)(kLBN + cb), w2);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float4 o;
                        o.x = __uint_as_float(v[j4 * 4 + 0]) + __uint_as_float(w2[j4 * 4 + 0]);
                        o.y = __uint_as_float(v[j4 * 4 + 1]) + __uint_as_float(w2[j4 * 4 + 1]);
                        o.z = __uint_as_float(v[j4 * 4 + 2]) + __uint_as_float(w2[j4 * 4 + 2]);
                        o.w = __uint_as_float(v[j4 * 4 + 3]) + __uint_as_float(w2[j4 * 4 + 3]);
                        *reinterpret_cast<float4*>(stg + lane * kLStgLd + j4 * 4) = o;
                    }
                }
                if (cb + 32 >= kLBN) {                             // whole accumulator set is in registers / smem now
                    tc_fence_before();
                    mbar_arrive(&acc_empty[i & 1]);
                }
                __syncwarp();
                const int n = nt * kLBN + cb + c4;
                const int nv = min(4, g.N - n);
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (EPI == 0 && bias) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (e < nv) bv[e] = __ldg(bias + n + e);
                }
                float4 mk[8];
                if (EPI == 1) {
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const int m = mt * BM + q * 32 + p * 4 + rsub;
                        mk[p] = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (g.mask && m < g.M && nv == 4) mk[p] = __ldg(reinterpret_cast<const float4*>(g.mask + (size_t)m * g.ldc + n));
                        else if (g.mask && m < g.M && nv > 0) {
                            const float* mp = g.mask + (size_t)m * g.ldc + n;
                            mk[p].x = __ldg(mp);
                            if (nv > 1) mk[p].y = __ldg(mp + 1);
                            if (nv > 2) mk[p].z = __ldg(mp + 2);
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int r = p * 4 + rsub;
                    const int m = mt * BM + q * 32 + r;
                    if (m < g.M && nv > 0) {
                        const float4 t4 = *reinterpret_cast<const float4*>(stg + r * kLStgLd + c4);
                        float x[4] = {t4.x, t4.y, t4.z, t4.w};
                        if (EPI == 0) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float y = x[e] + bv[e];
                                x[e] = g.relu ? fmaxf(y, 0.f) : y;
                            }
                        } else {
                            x[0] = mk[p].x > 0.f ? x[0] : 0.f; x[1] = mk[p].y > 0.f ? x[1] : 0.f;
                            x[2] = mk[p].z > 0.f ? x[2] : 0.f; x[3] = mk[p].w > 0.f ? x[3] : 0.f;
                        }
                        const size_t off = (size_t)m * g.ldc + n;
                        if (nv == 4) {
                            if (C) *reinterpret_cast<float4*>(C + off) = make_float4(x[0], x[1], x[2], x[3]);
                            if (Chi) {
                                float h[4], l[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    h[e] = __uint_as_float(to_tf32(x[e]));
                                    l[e] = __uint_as_float(to_tf32(x[e] - h[e]));
                                }
                                *reinterpret_cast<float4*>(Chi + off) = make_float4(h[0], h[1], h[2], h[3]);
                                *reinterpret_cast<float4*>(Clo + off) = make_float4(l[0], l[1], l[2], l[3]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (e < nv) {
                                    if (C) C[off + e] = x[e];
                                    if (Chi) {
                                        const float h = __uint_as_float(to_tf32(x[e]));
                                        Chi[off + e] = h; Clo[off + e] = __uint_as_float(to_tf32(x[e] - h));
                                    }
                                }
                        }
                    }
                }
                __syncwarp();                                      // staging is rewritten by the next slice
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---- host side --------------------------------------------------------------------------------------------
typedef CUresult (*PlEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PlEncodeFn pl_encode() {
    static PlEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PlEncodeFn>(p);
    }
    return fn;
}

// row-major [rows, cols] fp32 plane, box {32 cols, 128 rows}, SWIZZLE_128B -> the K-major UMMA tile; OOB reads give zeros
static int pl_make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld) {
    PlEncodeFn enc = pl_encode();
    DPPO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DPPO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for plane [%d x %d] ld %d", (int)r, rows, cols, ld);
    return 0;
}

__global__ void split_planes_kernel(const float* __restrict__ x, long long n4, float* __restrict__ hi, float* __restrict__ lo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        float4 h, l;
        h.x = __uint_as_float(to_tf32(v.x)); h.y = __uint_as_float(to_tf32(v.y));
        h.z = __uint_as_float(to_tf32(v.z)); h.w = __uint_as_float(to_tf32(v.w));
        l.x = __uint_as_float(to_tf32(v.x - h.x)); l.y = __uint_as_float(to_tf32(v.y - h.y));
        l.z = __uint_as_float(to_tf32(v.z - h.z)); l.w = __uint_as_float(to_tf32(v.w - h.w));
        reinterpret_cast<float4*>(hi)[i] = h;
        reinterpret_cast<float4*>(lo)[i] = l;
    }
}

}  // namespace tc

// x[n] -> hi[n], lo[n]   (n % 4 == 0, 16-B aligned)
int tc_split_planes(const float* x, long long n, float* hi, float* lo, cudaStream_t st) {
    DPPO_REQUIRE(n % 4 == 0, "tc_split_planes: n=%lld must be a multiple of 4", n);
    if (n == 0) return 0;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ProfScope ps(KK_MISC, st);
    tc::split_planes_kernel<<<(int)blocks, 256, 0, st>>>(x, n / 4, hi, lo);
    DPPO_LAUNCH_CHECK();
    return 0;
}

bool tc_pl_supported(const float* a_hi, const float* a_lo, int lda) {
    return tc::pl_encode() != nullptr && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(a_hi) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(a_lo) & 15) == 0;
}

long long tc_pl_image_replica_floats(int N, int K) {
    return (long long)((N + tc::kLBN - 1) / tc::kLBN) * ((K + tc::kLBK - 1) / tc::kLBK) * tc::kLBN * tc::kLBK * 2;
}
long long tc_pl_image_floats(int N, int K) { return tc::kLImgReplicas * tc_pl_image_replica_floats(N, K); }

// C[M,N] = epi(A[M,K] . B^T): A as hi/lo planes (row stride lda), B as weight images packed with bn = 128, bk = 32.
// C / Chi,Clo: fp32 result and/or its planes (row stride ldc).  nets = 1 or 2 products sharing shapes.
int tc_pl_gemm(int epi_nt, int M, int N, int K, const float* Ahi0, const float* Alo0, const float* Ahi1, const float* Alo1,
               int lda, const float* img0, const float* img1, float* C0, float* C1, float* Chi0, float* Clo0, float* Chi1,
               float* Clo1, int ldc, const float* bias0, const float* bias1, int relu, const float* mask, int nets,
               cudaStream_t st) {
    if (M == 0) return 0;
    DPPO_REQUIRE(ldc % 4 == 0, "tc_pl_gemm: ldc=%d must be a multiple of 4", ldc);
    tc::PlArgs g{};
    g.Bimg[0] = img0; g.Bimg[1] = img1; g.C[0] = C0; g.C[1] = C1; g.Chi[0] = Chi0; g.Clo[0] = Clo0; g.Chi[1] = Chi1; g.Clo[1] = Clo1;
    g.bias[0] = bias0; g.bias[1] = bias1; g.mask = mask; g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.nets = nets; g.relu = relu;
    g.tiles_n = (N + tc::kLBN - 1) / tc::kLBN;
    g.img_replica_bytes = tc_pl_image_replica_floats(N, K) * 4;
    CUtensorMap m[4];
    if (int r = tc::pl_make_map(&m[0], Ahi0, M, K, lda)) return r;
    if (int r = tc::pl_make_map(&m[1], Alo0, M, K, lda)) return r;
    if (nets > 1) {
        if (int r = tc::pl_make_map(&m[2], Ahi1, M, K, lda)) return r;
        if (int r = tc::pl_make_map(&m[3], Alo1, M, K, lda)) return r;
    } else { m[2] = m[0]; m[3] = m[1]; }
    const int tiles_m = (M + tc::BM - 1) / tc::BM;
    const long long total = (long long)tiles_m * g.tiles_n * nets;
    DPPO_REQUIRE(total < (1ll << 30), "tc_pl_gemm: too many tiles");
    static PerDeviceOnce configured[2];
    if (configured[epi_nt ? 1 : 0].need()) {
        if (epi_nt) DPPO_CHECK_CUDA(cudaFuncSetAttribute(tc::tc_pl_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kLSmemBytes));
        else DPPO_CHECK_CUDA(cudaFuncSetAttribute(tc::tc_pl_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kLSmemBytes));
    }
    const int grid = (int)(total < kPlanSMs ? total : kPlanSMs);
    ProfScope ps(KK_TC_GEMM, st);
    if (epi_nt) tc::tc_pl_kernel<1><<<grid, tc::kLThreads, tc::kLSmemBytes, st>>>(m[0], m[1], m[2], m[3], g, (int)total);
    else tc::tc_pl_kernel<0><<<grid, tc::kLThreads, tc::kLSmemBytes, st>>>(m[0], m[1], m[2], m[3], g, (int)total);
    DPPO_LAUNCH_CHECK();
    return 0;
}

}  // namespace dppo
