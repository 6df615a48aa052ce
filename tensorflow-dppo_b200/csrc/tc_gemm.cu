// Tensor-core dense layers: tcgen05.mma kind::tf32 with TMEM accumulators (sm_100a).
//
// Same three products as sgemm.cu (forward NN, dX NT, dW TN -- /root/reference/Model.py:12-14 and
// the auto-diff of PPO.py:46), but on the 5th-generation tensor cores.  fp32 parity (1e-5) is kept
// with the 3-term TF32 split  a = a_hi + a_lo (both round-to-nearest tf32):
//        A.B  ~=  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi        (fp32 accumulate in TMEM)
// whose dropped term a_lo.b_lo is ~2^-22 relative.  TERMS = 1 is plain TF32 (used by tests to
// validate layouts, not by the update path).
//
// CTA = 9 warps.  Warps 0-7 stream fp32 operand tiles global -> registers -> (split) -> shared
// memory in the canonical UMMA layouts (K-major SWIZZLE_128B, or MN-major SWIZZLE_128B_BASE32B for
// the dW product where the reduction dimension is the slow axis of both operands), signalling an
// mbarrier per stage; warp 8 owns TMEM and issues the MMAs from one thread, releasing stages with
// tcgen05.commit; after the K loop warps 0-7 read the accumulator tile back with tcgen05.ld
// (warp w -> TMEM lanes 32*(w%4).., columns split by w/4) and apply the fused epilogue.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace dppo {
namespace tc {

constexpr int BK = 32;           // 32 tf32 = 128 B = one swizzle row; 4 MMAs of K=8 per block
constexpr int kLoadWarps = 8;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kThreads = kLoadThreads + 64;   // + MMA warp + weight-image producer warp
constexpr int kImgReplicas = 8;                // copies of every weight image: CTAs spread their bulk reads over them (L2 hot lines)
constexpr int kPrefetch = 3;                  // A k-blocks in flight per loader thread when B arrives by bulk copy

enum { EPI_NN = 0, EPI_NT = 1, EPI_TN = 2 };

struct TcArgs {
    const float* A[2]; const float* B[2]; float* C[2]; const float* bias[2];
    const float* Bimg[2];      // B_IMAGE: weights pre-split and pre-swizzled as smem tile images (pack_images_kernel)
    const float* mask;
    // ReLU sign bitmap of an activation matrix (4 MB instead of the 134 MB fp32 tile at cfg3): the forward epilogue
    // writes it next to h, the dX epilogue of the same width reads it as its mask.  The layout is private to the two
    // epilogues, which share the thread -> (row, column) mapping: one word per (CTA tile, column half, lane quarter,
    // row batch, lane) = 8 rows x 4 columns, bit u*4 + e  ->  every warp access is one coalesced 128-B line.
    uint32_t* bits_out[2];
    const uint32_t* bits_in;
    int M, N, K;
    int lda, ldb, ldc;
    int tiles_n, nets;
    int relu, accumulate;
    int k_per_slice;
    long long slice_stride;
    int vecA, vecB, vecC;
    int raw_hi;                     // image mode: activation hi operand = the raw fp32 value (see tc_raw_hi)
    long long img_replica_bytes;    // B_IMAGE: byte distance between replicas of the weight image
};

// ---- tile movers -----------------------------------------------------------------------------------
// One operand tile is R (128 or BN) x KB fp32 = R*KB*4 B (KB = 32: 128-B rows, SWIZZLE_128B; KB = 16: 64-B rows,
// SWIZZLE_64B, K-major only).  Each loader thread owns R*KB/1024 16-B chunks.
template <int R, bool MN_MAJOR, int KB = BK>
__device__ __forceinline__ void tile_load(float4 (&reg)[R * KB / 1024], const float* __restrict__ P, int ld, int r0, int rdim,
                                          int k0, int kend, int vec, int tid) {
    static_assert(KB == 32 || (KB == 16 && !MN_MAJOR), "16-wide k-blocks exist for the K-major layout only");
    constexpr int kCpr = KB / 4;               // 16-B chunks per tile row
#pragma unroll
    for (int i = 0; i < R * KB / 1024; ++i) {
        const int q = tid + i * kLoadThreads;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!MN_MAJOR) {                       // element (r,k) at P[r*ld + k]
            const int r = r0 + q / kCpr, k = k0 + (q % kCpr) * 4;
            if (r < rdim && k < kend) {
                const float* p = P + (size_t)r * ld + k;
                if (vec && k + 3 < kend) v = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    v.x = __ldg(p);
                    if (k + 1 < kend) v.y = __ldg(p + 1);
                    if (k + 2 < kend) v.z = __ldg(p + 2);
                    if (k + 3 < kend) v.w = __ldg(p + 3);
                }
            }
        } else {                               // element (r,k) at P[k*ld + r]
            const int k = k0 + q / (R / 4), r = r0 + (q % (R / 4)) * 4;
            if (k < kend && r < rdim) {
                const float* p = P + (size_t)k * ld + r;
                if (vec && r + 3 < rdim) v = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    v.x = __ldg(p);
                    if (r + 1 < rdim) v.y = __ldg(p + 1);
                    if (r + 2 < rdim) v.z = __ldg(p + 2);
                    if (r + 3 < rdim) v.w = __ldg(p + 3);
                }
            }
        }
        reg[i] = v;
    }
}

template <int R, bool MN_MAJOR, int TERMS, int KB = BK>
__device__ __forceinline__ void tile_store(const float4 (&reg)[R * KB / 1024], uint8_t* hi, uint8_t* lo, int tid,
                                           bool raw_hi = false) {
    constexpr int kCpr = KB / 4;
#pragma unroll
    for (int i = 0; i < R * KB / 1024; ++i) {
        const int q = tid + i * kLoadThreads;
        uint32_t off;
        if (!MN_MAJOR) {
            const int r = q / kCpr, c = q % kCpr;
            off = (KB == 32) ? (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4)            // SWIZZLE_128B: chunk ^= row % 8
                             : (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4);     // SWIZZLE_64B:  chunk ^= (row / 2) % 4
        } else {
            const int kr = q / (R / 4), cc = q % (R / 4);
            const int atom = cc >> 3, c = cc & 7;
            off = (uint32_t)atom * (BK * 128u) + (uint32_t)kr * 128u + (uint32_t)((((c >> 1) ^ (kr & 3)) << 5) | ((c & 1) << 4));
        }
        const float4 v = reg[i];
        uint4 h;
        if (TERMS == 3 && raw_hi) {
            // the MMA ignores the low 13 mantissa bits: store x itself as the hi operand, split off x - trunc(x) as lo
            *reinterpret_cast<float4*>(hi + off) = v;
            h.x = __float_as_uint(v.x) & 0xFFFFE000u; h.y = __float_as_uint(v.y) & 0xFFFFE000u;
            h.z = __float_as_uint(v.z) & 0xFFFFE000u; h.w = __float_as_uint(v.w) & 0xFFFFE000u;
        } else {
            h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
            *reinterpret_cast<uint4*>(hi + off) = h;
        }
        if (TERMS == 3) {
            uint4 l;
            l.x = to_tf32(v.x - __uint_as_float(h.x)); l.y = to_tf32(v.y - __uint_as_float(h.y));
            l.z = to_tf32(v.z - __uint_as_float(h.z)); l.w = to_tf32(v.w - __uint_as_float(h.w));
            *reinterpret_cast<uint4*>(lo + off) = l;
        }
    }
}

template <int BN, int TERMS, int KB = BK>
struct SmemPlan {
    static constexpr int kATile = BM * KB * 4, kBTile = BN * KB * 4;
    static constexpr int kStage = TERMS == 3 ? 2 * (kATile + kBTile) : (kATile + kBTile);
    // Two CTAs per SM whenever they fit: one CTA's prologue / epilogue (TMEM drain, mask loads, C stores) then runs under
    // the other's MMA main loop instead of leaving the tensor pipe idle.
    //   BN = 64 (the 35-wide heads), 128-B rows: 2 stages x 48 KB;
    //   BN = 128 with 64-B rows (KB = 16):       3 stages x 32 KB, 2 x 256 TMEM columns.
    static constexpr int kMinBlocks = (BN <= 64 || KB == 16) ? 2 : 1;
    static constexpr int kStages = KB == 16 ? 3 : (BN <= 64 ? 2 : ((200 * 1024) / kStage >= 4 ? 4 : (200 * 1024) / kStage));
    static constexpr int kBytes = kStages * kStage + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(kStages >= 2, "need at least a double buffer");
    static_assert(KB == 32 || (BN <= 128 && kBytes <= 112 * 1024), "KB = 16 tiles are meant for two CTAs per SM");
};

template <int BN, bool MN_MAJOR, int TERMS, int EPI, bool B_IMAGE, int KB = BK>
__global__ void __launch_bounds__(kThreads, SmemPlan<BN, TERMS, KB>::kMinBlocks) tc_gemm_kernel(const TcArgs g) {
    static_assert(!(B_IMAGE && MN_MAJOR), "weight images exist for the K-major products only");
    static_assert(KB == 32 || (B_IMAGE && !MN_MAJOR), "KB = 16 is built for the image-fed K-major products");
    using Plan = SmemPlan<BN, TERMS, KB>;
    constexpr int kRegA = BM * KB / 1024;
    constexpr int STAGES = Plan::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Plan::kStage);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int net = blockIdx.x / g.tiles_n;
    const int n0 = (blockIdx.x % g.tiles_n) * BN;
    const int m0 = blockIdx.y * BM;
    const int slice = blockIdx.z;
    const int k_begin = slice * g.k_per_slice;
    const int k_end = min(g.K, k_begin + g.k_per_slice);
    const int nkb = max(0, (k_end - k_begin + KB - 1) / KB);

    // image mode: the first kPrefetch activation k-blocks are requested before the barrier/TMEM set-up, so their
    // global latency overlaps the CTA prologue
    float4 ra_img[B_IMAGE ? kPrefetch : 1][kRegA];
    if (B_IMAGE && warp < kLoadWarps) {
#pragma unroll
        for (int j = 0; j < kPrefetch; ++j)
            if (j < nkb) tile_load<BM, MN_MAJOR, KB>(ra_img[j], g.A[net], g.lda, m0, g.M, k_begin + j * KB, k_end, g.vecA, tid);
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], kLoadThreads + (B_IMAGE ? 1 : 0)); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    // TMEM: [0,BN) main accumulator (A_hi.B_hi); [BN,2BN) correction accumulator (A_lo.B_hi + A_hi.B_lo).
    // The tensor core truncates when it adds into an fp32 accumulator; keeping the 2^-11-scaled correction
    // terms out of the main accumulator cuts its number of truncating adds by 3 (precision).
    constexpr uint32_t kTmemCols = (TERMS == 3 ? 2 : 1) * (BN < 32 ? 32 : BN);
    if (warp == kLoadWarps) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kLoadWarps) {
        // ================= operand movers =================
        const float* __restrict__ A = g.A[net];
        const float* __restrict__ B = g.B[net];
        if (B_IMAGE) {
            // only the activation operand goes through registers: kPrefetch k-blocks of loads stay in flight
            float4 (&ra)[B_IMAGE ? kPrefetch : 1][kRegA] = ra_img;
            for (int kb0 = 0; kb0 < nkb; kb0 += kPrefetch) {
#pragma unroll
                for (int j = 0; j < kPrefetch; ++j) {
                    const int kb = kb0 + j;
                    if (kb < nkb) {
                        const int s = kb % STAGES;
                        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
                        uint8_t* st = smem + s * Plan::kStage;
                        tile_store<BM, MN_MAJOR, TERMS, KB>(ra[j], st, st + Plan::kATile, tid, g.raw_hi != 0);
                        fence_proxy_async();
                        mbar_arrive(&full_bar[s]);
                        if (kb + kPrefetch < nkb)
                            tile_load<BM, MN_MAJOR, KB>(ra[j], A, g.lda, m0, g.M, k_begin + (kb + kPrefetch) * KB, k_end, g.vecA, tid);
                    }
                }
            }
        } else if (KB == 32) {
        float4 ra[BM / 32], rb[BN / 32];
        if (nkb > 0) {
            tile_load<BM, MN_MAJOR>(ra, A, g.lda, m0, g.M, k_begin, k_end, g.vecA, tid);
            tile_load<BN, MN_MAJOR>(rb, B, g.ldb, n0, g.N, k_begin, k_end, g.vecB, tid);
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            float4 na[BM / 32], nb[BN / 32];
            if (kb + 1 < nkb) {
                const int k0 = k_begin + (kb + 1) * BK;
                tile_load<BM, MN_MAJOR>(na, A, g.lda, m0, g.M, k0, k_end, g.vecA, tid);
                tile_load<BN, MN_MAJOR>(nb, B, g.ldb, n0, g.N, k0, k_end, g.vecB, tid);
            }
            mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
            uint8_t* st = smem + s * Plan::kStage;
            uint8_t* a_hi = st;
            uint8_t* a_lo = st + Plan::kATile;
            uint8_t* b_hi = st + (TERMS == 3 ? 2 : 1) * Plan::kATile;
            uint8_t* b_lo = b_hi + Plan::kBTile;
            tile_store<BM, MN_MAJOR, TERMS>(ra, a_hi, a_lo, tid);
            tile_store<BN, MN_MAJOR, TERMS>(rb, b_hi, b_lo, tid);
            fence_proxy_async();               // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(&full_bar[s]);
            if (kb + 1 < nkb) {
#pragma unroll
                for (int i = 0; i < BM / 32; ++i) ra[i] = na[i];
#pragma unroll
                for (int i = 0; i < BN / 32; ++i) rb[i] = nb[i];
            }
        }
        }

        // ================= epilogue =================
        float* __restrict__ C = g.C[net];
        if (EPI == EPI_TN) C += (size_t)slice * g.slice_stride;
        const float* __restrict__ bias = g.bias[net];
        const int q = warp & 3, half = warp >> 2;
        // TMEM -> registers (thread = accumulator row) -> per-warp smem staging -> coalesced row-wise global
        // accesses (bias / ReLU-mask reads and the C store are full 128-B lines; a thread-per-row store would
        // scatter 16-B pieces over 32 rows per instruction).  The operand stages are free once accum_bar fired.
        constexpr int kColsPerHalf = BN / 2 < 32 ? 32 : BN / 2;
        constexpr int kStgLd = kColsPerHalf + 4;
        constexpr int kLanesPerRow = kColsPerHalf / 4;
        constexpr int kRowsPerIter = 32 / kLanesPerRow;
        constexpr int kIters = 32 / kRowsPerIter;
        constexpr int kBatch = kIters < 8 ? kIters : 8;       // rows whose mask / partial loads are in flight together
        constexpr int kNB = kIters / kBatch;
        // single-CTA-per-SM variants have registers to spare: the first batch of ReLU-mask (dX) / partial (dW +=) rows is
        // requested BEFORE waiting for the accumulator and batch b+1 while batch b is written, so that their global
        // latency hides under the MMA tail and the stores instead of standing between them
        constexpr bool kEarlyAux = (EPI != EPI_NN) && (Plan::kMinBlocks == 1);
        float* stg = reinterpret_cast<float*>(smem) + (size_t)warp * (32 * kStgLd);
        const int col0 = half * kColsPerHalf;
        const int cl = (lane % kLanesPerRow) * 4;
        const int n = n0 + col0 + cl;
        const int nv = min(4, g.N - n);
        static_assert(kBatch == 8, "sign-bitmap words hold 8 rows x 4 columns");
        const size_t bits_base = ((size_t)(blockIdx.y * g.tiles_n + (blockIdx.x % g.tiles_n)) * 8 + (size_t)(half * 4 + q)) * (kNB * 32) + lane;
        auto load_batch = [&](int it0, float4 (&aux)[kBatch]) {
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int r = (it0 + u) * kRowsPerIter + lane / kLanesPerRow;
                const int m = m0 + q * 32 + r;
                aux[u] = make_float4(1.f, 1.f, 1.f, 1.f);
                const float* src = (EPI == EPI_NT) ? g.mask : (g.accumulate ? C : nullptr);
                if (EPI == EPI_TN && !g.accumulate) aux[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (EPI == EPI_NT && g.bits_in) {            // one raw word per batch (decoded at use: no dependent op here)
                    if (u == 0) aux[0].x = __uint_as_float(__ldg(g.bits_in + bits_base + (size_t)(it0 / kBatch) * 32));
                } else if (src && m < g.M && nv > 0) {
                    const float* sp = src + (size_t)m * g.ldc + n;
                    if (g.vecC && nv == 4) aux[u] = __ldg(reinterpret_cast<const float4*>(sp));
                    else {
                        aux[u].x = sp[0];
                        if (nv > 1) aux[u].y = sp[1];
                        if (nv > 2) aux[u].z = sp[2];
                        if (nv > 3) aux[u].w = sp[3];
                    }
                }
            }
        };
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t* __restrict__ bits_out = (EPI == EPI_NN) ? g.bits_out[net] : nullptr;
        auto write_batch = [&](int it0, const float4 (&aux)[kBatch]) {
            uint32_t wbits = (EPI == EPI_NT) ? __float_as_uint(aux[0].x) : 0u;
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int r = (it0 + u) * kRowsPerIter + lane / kLanesPerRow;
                const int m = m0 + q * 32 + r;
                uint32_t nib = 0u;
                if (m < g.M && nv > 0) {
                    const float4 t4 = *reinterpret_cast<const float4*>(stg + r * kStgLd + cl);
                    float x[4] = {t4.x, t4.y, t4.z, t4.w};
                    float* cp = C + (size_t)m * g.ldc + n;
                    if (EPI == EPI_NN) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float y = x[e] + bv[e];
                            x[e] = g.relu ? fmaxf(y, 0.f) : y;
                        }
                        if (bits_out) {                      // warp-uniform: only the pi net's hidden layers keep a sign bitmap
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (e < nv && x[e] > 0.f) nib |= 1u << e;
                        }
                    } else if (EPI == EPI_NT) {
                        if (g.bits_in) {
                            const uint32_t b4 = wbits >> (4 * u);
                            x[0] = (b4 & 1u) ? x[0] : 0.f; x[1] = (b4 & 2u) ? x[1] : 0.f;
                            x[2] = (b4 & 4u) ? x[2] : 0.f; x[3] = (b4 & 8u) ? x[3] : 0.f;
                        } else {
                            x[0] = aux[u].x > 0.f ? x[0] : 0.f; x[1] = aux[u].y > 0.f ? x[1] : 0.f;
                            x[2] = aux[u].z > 0.f ? x[2] : 0.f; x[3] = aux[u].w > 0.f ? x[3] : 0.f;
                        }
                    } else {
                        x[0] += aux[u].x; x[1] += aux[u].y; x[2] += aux[u].z; x[3] += aux[u].w;
                    }
                    if (g.vecC && nv == 4) *reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
                    else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (e < nv) cp[e] = x[e];
                    }
                }
                if (EPI == EPI_NN) wbits |= nib << (4 * u);
            }
            if (EPI == EPI_NN && bits_out) bits_out[bits_base + (size_t)(it0 / kBatch) * 32] = wbits;
        };
        float4 aux_a[kBatch], aux_b[kBatch];
        if (kEarlyAux && col0 < BN) load_batch(0, aux_a);
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int cb = 0; cb < kColsPerHalf && col0 + cb < BN; cb += 32) {
            uint32_t v[32];
            if (nkb > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + cb), v);
                if (TERMS == 3) {
                    uint32_t w2[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + col0 + cb), w2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w2[j]));
                }
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
            if (EPI == EPI_NN && bias) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (e < nv) bv[e] = __ldg(bias + n + e);
            }
            if (EPI != EPI_NN && !kEarlyAux) load_batch(0, aux_a);
#pragma unroll 1
            for (int bidx = 0; bidx < kNB; bidx += 2) {
                if (EPI != EPI_NN && bidx + 1 < kNB) load_batch((bidx + 1) * kBatch, aux_b);
                write_batch(bidx * kBatch, aux_a);
                if (bidx + 1 < kNB) {
                    if (EPI != EPI_NN && bidx + 2 < kNB) load_batch((bidx + 2) * kBatch, aux_a);
                    write_batch((bidx + 1) * kBatch, aux_b);
                }
            }
        }
        tc_fence_before();
    } else if (warp == kLoadWarps + 1) {
        // ================= weight-image producer (one thread): bulk copies of ready-made B tiles =================
        if (B_IMAGE && lane == 0) {
            constexpr uint32_t kBytes = (TERMS == 3 ? 2 : 1) * Plan::kBTile;
            const int kbs_total = (g.K + KB - 1) / KB;
            const uint8_t* img = reinterpret_cast<const uint8_t*>(g.Bimg[net]) +
                                 (size_t)(blockIdx.y % kImgReplicas) * g.img_replica_bytes +
                                 (size_t)(blockIdx.x % g.tiles_n) * kbs_total * (2 * Plan::kBTile);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
                uint8_t* dst = smem + s * Plan::kStage + (TERMS == 3 ? 2 : 1) * Plan::kATile;
                mbar_arrive_expect_tx(&full_bar[s], kBytes);
                bulk_copy_g2s(dst, img + (size_t)(k_begin / KB + kb) * (2 * Plan::kBTile), kBytes, &full_bar[s]);
            }
        }
    } else if (lane == 0) {
        // ================= MMA issuer (one thread) =================
        constexpr uint32_t idesc = make_idesc(BN, MN_MAJOR, MN_MAJOR);
        constexpr uint32_t kLayout = MN_MAJOR ? kLayoutSW128Base32 : (KB == 32 ? kLayoutSW128 : kLayoutSW64);
        constexpr uint32_t kLbo = MN_MAJOR ? (uint32_t)(BK * 128) : 0u;         // MN-major: stride between 32-element MN atoms
        constexpr uint32_t kSbo = MN_MAJOR ? 512u : (KB == 32 ? 1024u : 512u);  // MN: between 4-row k groups; K: between 8-row groups
        constexpr uint32_t kStepBytes = MN_MAJOR ? 8u * 128u : 32u;             // advance of one UMMA_K = 8
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(&full_bar[s], (kb / STAGES) & 1);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + s * Plan::kStage);
            const uint32_t a_hi = st, a_lo = st + Plan::kATile;
            const uint32_t b_hi = st + (TERMS == 3 ? 2 : 1) * Plan::kATile, b_lo = b_hi + Plan::kBTile;
#pragma unroll
            for (int k = 0; k < KB / 8; ++k) {
                const uint64_t da_hi = make_sdesc(a_hi + k * kStepBytes, kLbo, kSbo, kLayout);
                const uint64_t db_hi = make_sdesc(b_hi + k * kStepBytes, kLbo, kSbo, kLayout);
                if (TERMS == 3) {
                    const uint64_t da_lo = make_sdesc(a_lo + k * kStepBytes, kLbo, kSbo, kLayout);
                    const uint64_t db_lo = make_sdesc(b_lo + k * kStepBytes, kLbo, kSbo, kLayout);
                    umma_tf32(tmem_base + BN, da_lo, db_hi, idesc, (kb | k) != 0);
                    umma_tf32(tmem_base + BN, da_hi, db_lo, idesc, 1u);
                    umma_tf32(tmem_base, da_hi, db_hi, idesc, (kb | k) != 0);
                } else {
                    umma_tf32(tmem_base, da_hi, db_hi, idesc, (kb | k) != 0);
                }
            }
            umma_commit(&empty_bar[s]);        // stage s may be overwritten once these MMAs retire
        }
        if (nkb > 0) umma_commit(accum_bar);   // accumulator complete
        tc_fence_before();
    }
    __syncthreads();
    if (warp == kLoadWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

static inline int aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN, bool MN_MAJOR, int TERMS, int EPI, bool B_IMAGE, int KB = BK>
static int launch(TcArgs& g, int slices, cudaStream_t st) {
    using Plan = SmemPlan<BN, TERMS, KB>;
    g.tiles_n = (g.N + BN - 1) / BN;
    const int tiles_m = (g.M + BM - 1) / BM;
    DPPO_REQUIRE(tiles_m <= 65535, "tc_gemm: M=%d too large for one launch", g.M);
    static PerDeviceOnce configured;
    if (configured.need()) {
        DPPO_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, MN_MAJOR, TERMS, EPI, B_IMAGE, KB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::kBytes));
    }
    dim3 grid(g.tiles_n * g.nets, tiles_m, slices);
    ProfScope ps(KK_TC_GEMM, st);
    tc_gemm_kernel<BN, MN_MAJOR, TERMS, EPI, B_IMAGE, KB><<<grid, kThreads, Plan::kBytes, st>>>(g);
    DPPO_LAUNCH_CHECK();
    return 0;
}

static void fill_vec_flags(TcArgs& g) {
    g.vecA = (g.lda % 4 == 0); g.vecB = (g.ldb % 4 == 0); g.vecC = (g.ldc % 4 == 0);
    for (int i = 0; i < g.nets; ++i) { g.vecA &= aligned16(g.A[i]); g.vecB &= aligned16(g.B[i]); g.vecC &= aligned16(g.C[i]); }
}

template <int TERMS, bool MN_MAJOR, int EPI, bool B_IMAGE = false>
static int dispatch_bn(TcArgs& g, int slices, cudaStream_t st) {
    if (g.N > 128) return launch<256, MN_MAJOR, TERMS, EPI, B_IMAGE>(g, slices, st);
    if (g.N > 64) return launch<128, MN_MAJOR, TERMS, EPI, B_IMAGE>(g, slices, st);
    return launch<64, MN_MAJOR, TERMS, EPI, B_IMAGE>(g, slices, st);
}

// ---- weight images: B tiles exactly as they must sit in shared memory (K-major SWIZZLE_128B, hi then lo), so a
// CTA fetches one (n-tile, k-block) with a single bulk copy instead of 8 warps of LDG/convert/STS ----
struct ImgEntry { const float* src; float* dst; int N, K, ld, transposed, bn, bk; };
struct ImgPlan { int n; ImgEntry e[3 * (kMaxLayers + 1)]; long long unit_end[3 * (kMaxLayers + 1)]; };

__global__ void pack_images_kernel(ImgPlan p, long long total_units) {
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total_units; u += (long long)gridDim.x * blockDim.x) {
        int i = 0;
        while (u >= p.unit_end[i]) ++i;
        const ImgEntry e = p.e[i];
        long long r_ = u - (i ? p.unit_end[i - 1] : 0);
        const int cpr = e.bk / 4;                              // 16-B chunks per tile row (128-B or 64-B rows)
        const int c = (int)(r_ % cpr); r_ /= cpr;
        const int r = (int)(r_ % e.bn); r_ /= e.bn;           // row within the tile
        const int kbs = (e.K + e.bk - 1) / e.bk;
        const int kb = (int)(r_ % kbs), nt = (int)(r_ / kbs);
        const int n = nt * e.bn + r, k0 = kb * e.bk + c * 4;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + q;
            v[q] = (n < e.N && k < e.K) ? (e.transposed ? __ldg(e.src + (size_t)k * e.ld + n) : __ldg(e.src + (size_t)n * e.ld + k)) : 0.f;
        }
        uint4 h, l;
        h.x = to_tf32(v[0]); h.y = to_tf32(v[1]); h.z = to_tf32(v[2]); h.w = to_tf32(v[3]);
        l.x = to_tf32(v[0] - __uint_as_float(h.x)); l.y = to_tf32(v[1] - __uint_as_float(h.y));
        l.z = to_tf32(v[2] - __uint_as_float(h.z)); l.w = to_tf32(v[3] - __uint_as_float(h.w));
        const size_t ntiles = (size_t)((e.N + e.bn - 1) / e.bn);
        const size_t plane = (size_t)e.bn * e.bk * 4;          // bytes of the hi (or lo) tile
        const size_t replica_bytes = ntiles * kbs * 2 * plane;
        const uint32_t off = (e.bk == 32) ? (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4)
                                          : (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4);
#pragma unroll
        for (int rep = 0; rep < kImgReplicas; ++rep) {
            uint8_t* tile = reinterpret_cast<uint8_t*>(e.dst) + rep * replica_bytes + ((size_t)nt * kbs + kb) * 2 * plane;
            *reinterpret_cast<uint4*>(tile + off) = h;
            *reinterpret_cast<uint4*>(tile + plane + off) = l;
        }
    }
}

}  // namespace tc

// ---- colsum(dY) for the bias gradients of the tensor-core path: slab s of rows -> out[s*stride + n] -------
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ G, int rows, int N, int ld, int rows_per_slab,
                                                     float* __restrict__ out, long long slab_stride, int accumulate) {
    __shared__ double red[4][64];
    const int col = threadIdx.x & 63, phase = threadIdx.x >> 6;
    const int n = blockIdx.x * 64 + col;
    const int r0 = blockIdx.y * rows_per_slab, r1 = min(rows, r0 + rows_per_slab);
    double acc = 0.0;
    if (n < N)
        for (int r = r0 + phase; r < r1; r += 4) acc += (double)__ldg(G + (size_t)r * ld + n);
    red[phase][col] = acc;
    __syncthreads();
    if (phase == 0 && n < N) {
        const float s = (float)(red[0][col] + red[1][col] + red[2][col] + red[3][col]);
        float* o = out + (size_t)blockIdx.y * slab_stride + n;
        *o = accumulate ? *o + s : s;
    }
}

int tc_colsum(const float* G, int rows, int N, int ld, int slabs, float* out, long long slab_stride, int accumulate,
              cudaStream_t st) {
    if (slabs < 1) slabs = 1;
    const int rps = (rows + slabs - 1) / slabs;
    dim3 grid((N + 63) / 64, slabs);
    ProfScope ps(KK_MISC, st);
    colsum_kernel<<<grid, 256, 0, st>>>(G, rows, N, ld, rps > 0 ? rps : 1, out, slab_stride, accumulate);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// ---- Wt[n][k] = W[k][n] for every dense layer of one net (K-major B operand of the forward product) -----
struct TransposePlan { int n_layers; long long off[kMaxLayers + 1]; int in[kMaxLayers + 1], out[kMaxLayers + 1]; };
__global__ void transpose_weights_kernel(TransposePlan p, const float* __restrict__ trunk0, const float* __restrict__ head0,
                                         float* __restrict__ wt0, const float* __restrict__ trunk1,
                                         const float* __restrict__ head1, float* __restrict__ wt1, long long total) {
    const float* trunk = blockIdx.y ? trunk1 : trunk0;
    const float* head = blockIdx.y ? head1 : head0;
    float* wt = blockIdx.y ? wt1 : wt0;
    if (trunk == nullptr) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int l = 0;
        while (l + 1 < p.n_layers && i >= p.off[l + 1]) ++l;
        const long long r = i - p.off[l];
        if (r >= (long long)p.in[l] * p.out[l]) continue;      // bias slots are not transposed
        const int n = (int)(r / p.in[l]), k = (int)(r % p.in[l]);
        const float* src = (l == p.n_layers - 1) ? head : trunk + p.off[l];
        wt[i] = src[(size_t)k * p.out[l] + n];
    }
}

int tc_transpose_weights(const NetShape& s, const float* params_pi, const float* head_pi, float* wt_pi,
                         const float* params_old, const float* head_old, float* wt_old, cudaStream_t st) {
    TransposePlan p;
    p.n_layers = s.L + 1;
    for (int l = 0; l <= s.L; ++l) { p.off[l] = s.w_off[l]; p.in[l] = s.in_dim(l); p.out[l] = s.out_dim(l); }
    dim3 grid((unsigned)((s.n_params + 255) / 256 > 592 ? 592 : (s.n_params + 255) / 256), 2);
    ProfScope ps(KK_MISC, st);
    transpose_weights_kernel<<<grid, 256, 0, st>>>(p, params_pi, head_pi, wt_pi, params_old, head_old, wt_old, s.n_params);
    DPPO_LAUNCH_CHECK();
    return 0;
}

static inline int tc_bk_for(int) { return tc::BK; }
static inline int tc_bn_for(int n) { return n > 128 ? 256 : (n > 64 ? 128 : 64); }

// floats of one weight image: B is [N,K]
static long long tc_image_replica_floats(int N, int K) {
    const int bn = tc_bn_for(N), bk = tc_bk_for(N);
    return (long long)((N + bn - 1) / bn) * ((K + bk - 1) / bk) * bn * bk * 2;          // hi + lo tile per (n-tile, k-block)
}
long long tc_image_floats(int N, int K) { return tc::kImgReplicas * tc_image_replica_floats(N, K); }

// Build every weight image of one update: forward images of pi (and oldpi) from W_l^T, dX images of pi from W_l.
// img_* are arrays of L+1 device pointers (nullptr = skip).
int tc_pack_images(const NetShape& s, const float* params_pi, const float* head_pi, const float* params_old,
                   const float* head_old, float* const* img_fwd_pi, float* const* img_fwd_old, float* const* img_dx_pi,
                   cudaStream_t st) {
    tc::ImgPlan p; p.n = 0;
    long long units = 0;
    auto add = [&](const float* src, float* dst, int N, int K, int ld, int transposed) {
        if (!src || !dst) return;
        tc::ImgEntry& e = p.e[p.n];
        e.src = src; e.dst = dst; e.N = N; e.K = K; e.ld = ld; e.transposed = transposed; e.bn = tc_bn_for(N);
        e.bk = tc_bk_for(N);
        units += (long long)((N + e.bn - 1) / e.bn) * ((K + e.bk - 1) / e.bk) * e.bn * (e.bk / 4);
        p.unit_end[p.n++] = units;
    };
    for (int l = 0; l <= s.L; ++l) {
        const int in = s.in_dim(l), out = s.out_dim(l);
        const float* w_pi = (l == s.L) ? head_pi : params_pi + s.w_off[l];
        const float* w_old = (l == s.L) ? head_old : (params_old ? params_old + s.w_off[l] : nullptr);
        add(w_pi, img_fwd_pi[l], out, in, out, 1);                 // forward: B[n=out][k=in] = W[k][n]
        add(params_old ? w_old : nullptr, img_fwd_old[l], out, in, out, 1);
        if (l > 0) add(w_pi, img_dx_pi[l], in, out, out, 0);      // dX:      B[n=in][k=out] = W[n][k]
    }
    if (p.n == 0) return 0;
    int grid = (int)((units + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope ps(KK_MISC, st);
    tc::pack_images_kernel<<<grid, 256, 0, st>>>(p, units);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// one image from a [N,K] (transposed = 0) or [K,N] (transposed = 1) matrix with row stride ld
int tc_pack_one(const float* src, float* dst, int N, int K, int ld, int transposed, cudaStream_t st) {
    tc::ImgPlan p; p.n = 1;
    tc::ImgEntry& e = p.e[0];
    e.src = src; e.dst = dst; e.N = N; e.K = K; e.ld = ld; e.transposed = transposed; e.bn = tc_bn_for(N);
    e.bk = tc_bk_for(N);
    const long long units = (long long)((N + e.bn - 1) / e.bn) * ((K + e.bk - 1) / e.bk) * e.bn * (e.bk / 4);
    p.unit_end[0] = units;
    int grid = (int)((units + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope ps(KK_MISC, st);
    tc::pack_images_kernel<<<grid, 256, 0, st>>>(p, units);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// C[M,N] = epi(A[M,K] . Bimg^T) with B given as weight images (see tc_pack_images)
bool tc_bits_supported() { return true; }
bool tc_raw_hi() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DPPO_TC_RAWHI"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
// 32-bit words of the sign bitmap of a [rows, N] activation matrix (whole 128 x BN tiles)
long long tc_bits_words(long long rows, int N) {
    const int bn = tc_bn_for(N);
    return ((rows + tc::BM - 1) / tc::BM) * ((N + bn - 1) / bn) * (long long)(tc::BM * bn / 32);
}

int tc_gemm_kk_img(int epi_nt, int M, int N, int K, const float* A0, const float* A1, int lda, const float* img0,
                   const float* img1, float* C0, float* C1, int ldc, const float* bias0, const float* bias1, int relu,
                   const float* mask, int nets, cudaStream_t st, uint32_t* bits_out0, const uint32_t* bits_in) {
    if (M == 0) return 0;
    tc::TcArgs g{};
    if (tc_bits_supported()) { g.bits_out[0] = bits_out0; g.bits_in = bits_in; }
    g.raw_hi = tc_raw_hi() ? 1 : 0;
    g.A[0] = A0; g.A[1] = A1; g.Bimg[0] = img0; g.Bimg[1] = img1; g.C[0] = C0; g.C[1] = C1; g.bias[0] = bias0; g.bias[1] = bias1;
    g.mask = mask; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = 4; g.ldc = ldc; g.nets = nets; g.relu = relu;
    g.k_per_slice = ((K + tc::BK - 1) / tc::BK) * tc::BK;
    g.B[0] = A0; g.B[1] = A1;          // unused in image mode; keeps fill_vec_flags well-defined
    g.img_replica_bytes = tc_image_replica_floats(N, K) * 4;
    tc::fill_vec_flags(g);
    if (epi_nt) return tc::dispatch_bn<3, false, tc::EPI_NT, true>(g, 1, st);
    return tc::dispatch_bn<3, false, tc::EPI_NN, true>(g, 1, st);
}

// C[M,N] = epi(A[M,K] . Bt[N,K]^T): both operands K-contiguous.  nets = 1 or 2 products sharing shapes.
int tc_gemm_kk(int terms, int epi_nt, int M, int N, int K, const float* A0, const float* A1, int lda, const float* B0,
               const float* B1, int ldb, float* C0, float* C1, int ldc, const float* bias0, const float* bias1,
               int relu, const float* mask, int nets, cudaStream_t st) {
    if (M == 0) return 0;
    tc::TcArgs g{};
    g.A[0] = A0; g.A[1] = A1; g.B[0] = B0; g.B[1] = B1; g.C[0] = C0; g.C[1] = C1; g.bias[0] = bias0; g.bias[1] = bias1;
    g.mask = mask; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.nets = nets; g.relu = relu;
    g.k_per_slice = ((K + tc::BK - 1) / tc::BK) * tc::BK;
    tc::fill_vec_flags(g);
    if (terms == 3) {
        if (epi_nt) return tc::dispatch_bn<3, false, tc::EPI_NT>(g, 1, st);
        return tc::dispatch_bn<3, false, tc::EPI_NN>(g, 1, st);
    }
    if (epi_nt) return tc::dispatch_bn<1, false, tc::EPI_NT>(g, 1, st);
    return tc::dispatch_bn<1, false, tc::EPI_NN>(g, 1, st);
}

// C[s][M,N] (+)= A[k,M]^T . B[k,N] over k in slice s: both operands MN-contiguous (reduction over rows)
int tc_gemm_tn_splitk(int terms, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                      int slices, long long slice_stride, int accumulate, cudaStream_t st) {
    tc::TcArgs g{};
    g.A[0] = A; g.B[0] = B; g.C[0] = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.nets = 1;
    g.accumulate = accumulate; g.slice_stride = slice_stride;
    if (slices < 1) slices = 1;
    int kps = (K + slices - 1) / slices;
    kps = ((kps + tc::BK - 1) / tc::BK) * tc::BK;
    if (kps == 0) kps = tc::BK;
    g.k_per_slice = kps;
    tc::fill_vec_flags(g);
    if (g.slice_stride % 4 != 0) g.vecC = 0;
    if (terms == 3) return tc::dispatch_bn<3, true, tc::EPI_TN>(g, slices, st);
    return tc::dispatch_bn<1, true, tc::EPI_TN>(g, slices, st);
}

}  // namespace dppo
