// Persistent tensor-core dense layers on 16-bit operands: tcgen05.mma kind::f16, fp32 accumulation in TMEM (sm_100a).
//
// DPPO_GEMM_FP16X2: the forward (NN) and dX (NT) products of /root/reference/Model.py:12-14 and of the auto-diff of
// PPO.py:46 at fp32 accuracy on the HALF-precision tensor pipe (twice the tf32 rate, half the shared-memory bytes):
//        x = hi + lo / 2^11,   hi = RN-fp16(x),  lo = RN-fp16((x - hi) * 2^11)          (11 + 11 significand bits,
//        A.B ~= A_hi.B_hi + (A_lo.B_hi + A_hi.B_lo) / 2^11                                the same 22 as the tf32 split)
// The 2^11 pre-scale keeps `lo` inside fp16's normal range, and the correction products live in their own TMEM
// accumulator anyway (the tensor core truncates on every fp32 add: DESIGN.md, precision), so the scale costs nothing.
// Operands must lie in fp16's range (|x| < 65504; far below 1e-4 only `lo` carries precision).  Observations, activations
// and weights of a working policy do; the 1/B-sized gradients of the backward products do not, so the loss kernel emits
// d_out pre-multiplied by a power of two, the whole backward chain carries that factor and finalize_kernel divides it out
// (exact in fp32) -- no kernel of the chain spends an instruction on scaling.
//
// What is new against tc_gemm.cu, whose one-tile-per-CTA kernels leave the tensor pipe idle during every prologue and
// epilogue (45 % of their time at cfg3):
//   * persistent CTAs (grid = #SMs) walk output tiles; the operand rings run straight across tile boundaries;
//   * the epilogue has its own 8 warps, thread = accumulator row: 64-column sub-phases leave TMEM, are finished in
//     registers (correction / 2^11, bias, ReLU + sign word | mask word) and staged as 128-B rows of SWIZZLE_128B boxes
//     that ONE thread hands to TMA (cp.async.bulk.tensor store, clipped at the matrix edge); two staging buffers, so the
//     store of sub-phase k runs under the drain of k+1 and TMEM is held for the drains only.  (C rows TMA cannot address
//     -- a stride that is not a multiple of 16 B -- keep a coalesced st.global write-out.)
//   * all 512 TMEM columns are one 128x256 tile's main + correction accumulators, so a 256-wide tile cannot overlap its
//     drain with the next tile's MMAs; tiles of <= 128 columns get TWO accumulator stages and do.  Products with a short
//     reduction (K < 65: the dX of the heads) therefore run as 128-wide tiles even when N = 256;
//   * the activation operand arrives by TMA: one thread issues cp.async.bulk.tensor.2d boxes {32 fp32 x 128 rows,
//     SWIZZLE_128B} into a raw ring (2 slots); 8 converter warps read their 16-B chunks back, split them into fp16 hi / lo
//     and write the K-major SWIZZLE_64B operand tiles (2 slots).  Measured on the first version of this kernel: register-
//     staged movers (LDG -> cvt -> STS, 48 KB in flight per SM) delivered only 7 B/clk/SM and left the tensor pipe at 26 %
//     -- the L1 miss path cannot hold enough loads in flight; the TMA queue can.  Rows whose stride or base is not 16-B
//     aligned (obs_dim 3, 17, ...) keep the LDG movers (template flag);
//   * the weight operand has its own ring (3 slots, ONE cp.async.bulk of a ready-made tile image per k-block);
//   * ReLU sign bitmaps are plain row-major bit matrices (1 bit / activation), written by the forward epilogue and read by
//     the dX epilogue of the same width.
// Warps: 0-7 movers | 8 TMEM owner + MMA issuer (one thread) | 9 weight-image producer (one thread) | 10-17 epilogue.
#include "tc_h.cuh"
#include <stdlib.h>
#include <string.h>

namespace dppo {
namespace th {

constexpr int kPrefetch = 3;                                 // activation k-blocks in flight per mover thread
constexpr int kImgReplicas = 8;                              // copies of every weight image (spreads the hot L2 lines)
constexpr int kPeerSlots = 8;                                // CTA pairs: ring of "the peer's slots are full" barriers in the leader

enum { EPI_NN = 0, EPI_NT = 1 };

struct HArgs {
    const float* A[2]; const uint8_t* Bimg[2]; float* C[2]; const float* bias[2];
    const float* mask;                 // NT: fp32 mask [M,N] (row stride ldc), C *= (mask > 0); used when bits_in == nullptr
    const uint32_t* bits_in;           // NT: ReLU sign bitmap [M, bits_wpr]
    uint32_t* bits_out[2];             // NN: sign bitmap of the result (nullable per net)
    int M, N, K, lda, ldc, ldmask;
    int tiles_n, nets, relu, vecA, vecC, bits_wpr;
    long long img_replica_bytes;
    long long* timeline;               // optional: clock64 stamps of CTA 0 (tools/h_microbench.py)
    int terms;                         // 2 = fp16 hi + lo operands (three MMAs per product), 1 = bf16 operands (one MMA)
    int dbg;                           // DPPO_H_DBG: timing-only ablations (WRONG results), tools/h_ablate.py
};
enum { DBG_NO_A = 1, DBG_NO_SPLIT = 2, DBG_ONE_MMA = 4, DBG_NO_W = 8, DBG_NO_STORE = 16, DBG_NO_FENCE = 32,
       DBG_EARLY_RELEASE = 512 };     // 512: release the raw slot right after ISSUING its loads (the race described in the converter loop)

// How the activation operand reaches shared memory / how the result leaves it
enum { A_LDG = 0,        // LDG movers: fp32 rows TMA cannot address (unaligned stride)
       A_RAW = 1,        // TMA boxes of fp32 into a raw ring, converter warps split them into fp16 hi / lo
       A_PLANES = 2 };   // the producer kernel already stored fp16 hi | lo planes: TMA writes the operand tiles directly
enum { C_STG = 0,        // st.global write-out (fp32)
       C_TMA = 1,        // fp32 boxes by TMA store
       C_PLANES = 2 };   // fp16 hi | lo planes by TMA store (the next product's A_PLANES operand, the dW kernel's plane operands)

// T = terms per operand: 2 = fp16 hi + lo (fp32-accurate, three MMAs per product), 1 = one bf16 (DPPO_GEMM_BF16, one MMA)
// PAIR: two CTAs of a cluster work on a 256-row super-tile with cta_group::2 MMAs (M = 256): each CTA streams its own 128
// activation rows and HALF of the weight tile's rows, so the weight bytes an SM pulls from L2 halve -- the single-CTA kernel
// is bound by exactly that (48 KB per k-block at ~46 B/clk against 768 MMA cycles).  The leader (cluster rank 0) issues every
// MMA once both CTAs' slots are full -- every TMA load of either CTA counts its bytes on the LEADER's full barrier
// (cp.async.bulk.tensor .cta_group::2; a software relay of the peer's barriers cost ~900 cycles per remote arrive and left
// the pair slower than two single CTAs) --, releases slots and publishes accumulators with multicast commits; each CTA drains
// its own TMEM rows.  Pairs take fp16-plane (TMA-written) activation operands only.
// k-block: 32 reduction elements (64-byte fp16 rows, SWIZZLE_64B); the one-term mode on plane operands takes 64 (128-byte bf16
// rows, SWIZZLE_128B): its two MMAs per 32 elements left the per-k-block costs (barrier round trip, commit, TMA issue) in charge
template <int AMODE, int T> struct HKb { static constexpr int kKb = (T == 1 && AMODE == A_PLANES) ? 64 : 32; };

template <int BN, int AMODE = A_RAW, int T = 2, bool PAIR = false>
struct HPlan {
    static constexpr int kKb = HKb<AMODE, T>::kKb;
    static constexpr int kRowBytes = kKb * 2;               // one operand row of a k-block
    static constexpr int kSz = (T == 1 && kKb == 32) ? 1 : 2;   // slot size class: half-size slots only for one-term 32-element k-blocks
    // (a pair runs ONE ring: the activation and weight slots of a k-block are released together by a single multicast commit)
    static constexpr int kPairSlots = AMODE == A_PLANES ? (kSz == 1 ? 8 : 5) : (kSz == 1 ? 4 : 3);
    static constexpr int kBSlots = PAIR ? kPairSlots : (kSz == 1 ? 4 : 3);
    // raw (TMA) ring depth: whatever shared memory is left -- narrow tiles have short k-blocks and need more boxes in flight.
    // With fp16 planes there is no raw ring and no converter hand-off: its bytes go to a deeper operand ring.
    static constexpr int kASlots2 = AMODE == A_PLANES ? (BN >= 256 ? 4 : (BN >= 128 ? 6 : 8))
                                                      : (BN >= 256 ? 2 : (BN >= 128 ? 3 : 4));   // operand ring: latency-bound when k-blocks are short
    static constexpr int kASlots = PAIR ? kPairSlots : (kSz == 1 ? (2 * kASlots2 > 8 ? 8 : 2 * kASlots2) : kASlots2);    // (half-size slots in one-term mode)
    static constexpr int kRSlots = AMODE == A_PLANES ? 0 : (PAIR ? kPairSlots : (BN >= 256 ? (T == 1 ? 3 : 2) : (BN >= 128 ? 4 : 6)));
    static constexpr int kAHalf = BM * kRowBytes;            // hi (or lo) activation tile: 128 rows x 64 B (128 B: 64-element k-blocks)
    static constexpr int kASlot = T * kAHalf;
    static constexpr int kBRows = PAIR ? BN / 2 : BN;        // a CTA of a pair holds half of the weight tile's rows
    static constexpr int kBHalf = kBRows * kRowBytes;
    static constexpr int kBSlot = T * kBHalf;
    static constexpr int kPhaseCols = BN < 128 ? BN : 128;   // columns staged per epilogue phase
    static constexpr int kPhases = BN / kPhaseCols;
    static constexpr int kStaging = BM * kPhaseCols * 4;
    static constexpr int kRawSlot = BM * 128;                // raw fp32 k-block: 128 rows x 32 floats
    static constexpr int kRawOff = kASlots * kASlot;
    static constexpr int kBOff = kRawOff + kRSlots * kRawSlot;
    static constexpr int kStgOff = kBOff + kBSlots * kBSlot;
    static constexpr int kBarOff = kStgOff + kStaging;
    static_assert((2 * kASlots + 2 * kBSlots + 2 * kRSlots + 4 + kPeerSlots) * 8 + 4 <= 512, "barrier block");
    static constexpr int kBiasOff = kBarOff + 512;            // the tile's bias vector (TMA write-out epilogue)
    static constexpr int kBytes = kBiasOff + BN * 4;
    static_assert(kStgOff % 1024 == 0, "staging boxes are SWIZZLE_128B");
    static_assert(kBytes <= 232448, "exceeds the 227 KB of shared memory one CTA may use");
};

// ---- activation movers: one k-block = 128 rows x 32 elements = 512 units of 8 elements, two per thread ----------------
struct MoverRegs { float4 v[2][2]; };

__device__ __forceinline__ void mover_load(MoverRegs& r, const float* __restrict__ A, int lda, int m0, int M, int k0, int K,
                                           int vec, int tid) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = (tid >> 2) + j * 64, c = tid & 3;
        const int m = m0 + row, k = k0 + c * 8;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (m < M && k < K) {
            const float* p = A + (size_t)m * lda + k;
            if (vec && k + 7 < K) {
                a = __ldg(reinterpret_cast<const float4*>(p));
                b = __ldg(reinterpret_cast<const float4*>(p) + 1);
            } else {
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = (k + e < K) ? __ldg(p + e) : 0.f;
                a = make_float4(t[0], t[1], t[2], t[3]); b = make_float4(t[4], t[5], t[6], t[7]);
            }
        }
        r.v[j][0] = a; r.v[j][1] = b;
    }
}

struct OperandRegs { uint4 h[2], l[2]; };
__device__ __forceinline__ void mover_split(const MoverRegs& r, OperandRegs& o) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float4 a = r.v[j][0], b = r.v[j][1];
        split2(a.x, a.y, o.h[j].x, o.l[j].x); split2(a.z, a.w, o.h[j].y, o.l[j].y);
        split2(b.x, b.y, o.h[j].z, o.l[j].z); split2(b.z, b.w, o.h[j].w, o.l[j].w);
    }
}
__device__ __forceinline__ void mover_store_split(const OperandRegs& o, uint32_t hi_tile, uint32_t lo_tile, int tid) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = (tid >> 2) + j * 64, c = tid & 3;
        const uint32_t off = (uint32_t)row * 64u + (uint32_t)((c ^ ((row >> 1) & 3)) << 4);     // SWIZZLE_64B
        sts128(hi_tile + off, o.h[j]);
        sts128(lo_tile + off, o.l[j]);
    }
}
__device__ __forceinline__ void mover_store(const MoverRegs& r, uint32_t hi_tile, uint32_t lo_tile, int tid) {
    OperandRegs o;
    mover_split(r, o);
    mover_store_split(o, hi_tile, lo_tile, tid);
}

// one-term mode: the two units as bf16 (round to nearest) into the single operand tile
__device__ __forceinline__ void mover_pack_bf16(const MoverRegs& r, OperandRegs& o) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float4 a = r.v[j][0], b = r.v[j][1];
        o.h[j] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    }
}
__device__ __forceinline__ void mover_store_packed(const OperandRegs& o, uint32_t tile, int tid) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = (tid >> 2) + j * 64, c = tid & 3;
        sts128(tile + (uint32_t)row * 64u + (uint32_t)((c ^ ((row >> 1) & 3)) << 4), o.h[j]);     // SWIZZLE_64B
    }
}

// timing ablation: the same stores without the fp32 -> fp16 split (bit patterns, wrong values)
__device__ __forceinline__ void mover_store_raw(const MoverRegs& r, uint32_t hi_tile, uint32_t lo_tile, int tid) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = (tid >> 2) + j * 64, c = tid & 3;
        const uint32_t off = (uint32_t)row * 64u + (uint32_t)((c ^ ((row >> 1) & 3)) << 4);
        const float4 a = r.v[j][0], b = r.v[j][1];
        sts128(hi_tile + off, make_uint4(__float_as_uint(a.x) & 0x3bff3bffu, __float_as_uint(a.y) & 0x3bff3bffu, __float_as_uint(a.z) & 0x3bff3bffu, __float_as_uint(a.w) & 0x3bff3bffu));
        sts128(lo_tile + off, make_uint4(__float_as_uint(b.x) & 0x3bff3bffu, __float_as_uint(b.y) & 0x3bff3bffu, __float_as_uint(b.z) & 0x3bff3bffu, __float_as_uint(b.w) & 0x3bff3bffu));
    }
}

// the same two units of a k-block read back from the raw tile TMA wrote (128-B rows, SWIZZLE_128B: chunk ^= row % 8)
__device__ __forceinline__ void raw_load(MoverRegs& r, uint32_t raw_tile, int tid) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = (tid >> 2) + j * 64, c = tid & 3;
        const uint32_t base = raw_tile + (uint32_t)row * 128u;
        r.v[j][0] = lds128f(base + (uint32_t)(((2 * c) ^ (row & 7)) << 4));
        r.v[j][1] = lds128f(base + (uint32_t)(((2 * c + 1) ^ (row & 7)) << 4));
    }
}

// The accumulators of a tile: main (hi.hi) and correction (lo.hi + hi.lo, scaled 2^11), BN columns each.  Tiles of up to 128
// columns leave room for TWO such stages in the 512 TMEM columns: the MMAs of tile i+1 run while tile i is drained.
template <int BN, int T> struct HAcc { static constexpr int kStages = (T == 2 && BN >= 256) ? 1 : 2; };

// the tensor maps of one launch: [net] x {fp32 matrix or hi plane, lo plane}
struct HMaps { CUtensorMap a[2], al[2], c[2], cl[2], b[2]; };      // b: the weight images as [rows of 64 B] (CTA pairs only)

template <int BN, int EPI, int AMODE, int CMODE, int T, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
h_gemm_kernel(const __grid_constant__ HMaps maps, const HArgs g, int total_tiles) {
    using P = HPlan<BN, AMODE, T, PAIR>;
    constexpr bool TMA_A = (AMODE == A_RAW);
    constexpr bool TMA_C = (CMODE != C_STG);
    constexpr int kBSlots = P::kBSlots, kASlots = P::kASlots, kRSlots = P::kRSlots, kAcc = HAcc<BN, T>::kStages;
    extern __shared__ __align__(1024) uint8_t smem[];         // SWIZZLE_128B boxes and UMMA tiles need 1024-B alignment
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t a_ring = smem_base, raw_ring = smem_base + P::kRawOff, b_ring = smem_base + P::kBOff;
    const uint32_t stg = smem_base + P::kStgOff;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + P::kBarOff);
    uint64_t* empty_a = full_a + kASlots;
    uint64_t* full_b = empty_a + kASlots;
    uint64_t* empty_b = full_b + kBSlots;
    uint64_t* raw_full = empty_b + kBSlots;
    uint64_t* raw_empty = raw_full + kRSlots;
    uint64_t* acc_full = raw_empty + kRSlots;
    uint64_t* acc_empty = acc_full + kAcc;
    uint64_t* peer_full = acc_empty + kAcc;         // PAIR, leader only: "the peer's slots of iteration it are full"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full + kPeerSlots);
    float* s_bias = reinterpret_cast<float*>(smem + P::kBiasOff);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int KB = P::kKb;                                // reduction elements per k-block
    const int nkb = (g.K + KB - 1) / KB;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int n_groups = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;         // a group = one CTA, or one CTA pair
    const int group = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
    const int per_cta = (total_tiles + n_groups - 1) / n_groups;
    const int t_begin = group * per_cta;
    const int t_end = min(total_tiles, t_begin + per_cta);
    const int my_tiles = max(0, t_end - t_begin);
    const long long total_it = (long long)my_tiles * nkb;

    if (tid == 0) {
        if (smem_base & 1023u) __trap();
        // (pair: the leader's two streams -- activations, weights -- each post the bytes of BOTH CTAs on full_a)
        for (int s = 0; s < kASlots; ++s) { mbar_init(&full_a[s], PAIR ? 2 : (AMODE == A_PLANES ? 1 : kMoverWarps)); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < kBSlots; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < kRSlots; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kMoverWarps); }
        for (int s = 0; s < kAcc; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], PAIR ? 2 * kEpiWarps : kEpiWarps); }
        for (int s = 0; s < kPeerSlots; ++s) mbar_init(&peer_full[s], 1);
        fence_barrier_init();
    }
    if (PAIR) cluster_sync_all();                   // both CTAs' barriers exist before anything arrives on them remotely
    constexpr uint32_t kStageCols = T * BN;         // [0,BN) main, [BN,2BN) correction (two-term mode)
    constexpr uint32_t kTmemCols = kAcc * kStageCols;
    if (warp == kMmaWarp) { if (PAIR) tmem_alloc_pair(tmem_slot, kTmemCols); else tmem_alloc(tmem_slot, kTmemCols); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile t -> (m-tile, net, n-tile): the two nets of one row block are neighbours, the second finds the rows in L2
    auto decode = [&](int t, int& mt, int& net, int& nt) {
        nt = t % g.tiles_n; t /= g.tiles_n;
        net = t % g.nets; mt = t / g.nets;
        if (PAIR) mt = mt * 2 + (int)rank;          // the pair's 256-row super-tile: rank r takes rows [r*128, r*128+128)
    };
    // the epilogue's "this tile has left TMEM": in a pair both CTAs report to the leader, whose MMA thread waits for it
    auto release_acc = [&](int as) {
        if (PAIR && rank != 0) mbar_arrive_remote(mapa_u32(smem_u32(&acc_empty[as]), 0));
        else mbar_arrive(&acc_empty[as]);
    };

    if (warp < kMoverWarps) {
        if (AMODE == A_PLANES) {
            // nothing to convert: TMA writes the fp16 hi / lo operand tiles
        } else if (TMA_A) {
            // ================= converters: raw fp32 tile (TMA) -> fp16 hi / lo operand tiles =================
            for (long long it = 0; it < total_it; ++it) {
                constexpr int kRS = kRSlots > 0 ? kRSlots : 1;
                const int rs = (int)(it % kRS), sa = (int)(it % kASlots);
                MoverRegs r;
                OperandRegs o;
                const bool old_order = (g.dbg & DBG_EARLY_RELEASE) != 0;
                if (!(g.dbg & DBG_NO_A)) {
                    mbar_wait(&raw_full[rs], (uint32_t)((it / kRS) & 1));
                    raw_load(r, raw_ring + rs * P::kRawSlot, tid);
                    if (old_order) warp_arrive(&raw_empty[rs], lane);
                } else {
                    r.v[0][0] = r.v[0][1] = r.v[1][0] = r.v[1][1] = make_float4(1.f, 2.f, 3.f, 4.f);
                }
                // The raw slot is released only after its values have been USED (the split below cannot issue before the
                // shared-memory loads have returned) and behind a proxy fence: the refill is an async-proxy write, and a
                // barrier arrive alone does not keep it behind this thread's still outstanding generic-proxy reads.
                // Releasing straight after issuing the loads let TMA overwrite rows that had not been read yet (seen as
                // "k-block kb+2 instead of kb" in rows of a tile's first k-blocks whenever the epilogue kept the LSU busy).
                if (T == 2) mover_split(r, o); else mover_pack_bf16(r, o);
                if (!(g.dbg & DBG_NO_A) && !old_order) {
                    fence_proxy_async();
                    warp_arrive(&raw_empty[rs], lane);
                }
                mbar_wait(&empty_a[sa], (uint32_t)(((it / kASlots) & 1) ^ 1));
                const uint32_t slot = a_ring + sa * P::kASlot;
                if (T == 1) mover_store_packed(o, slot, tid);
                else if (g.dbg & DBG_NO_SPLIT) mover_store_raw(r, slot, slot + P::kAHalf, tid);
                else mover_store_split(o, slot, slot + P::kAHalf, tid);
                if (!(g.dbg & DBG_NO_FENCE)) fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
                warp_arrive(&full_a[sa], lane);
            }
        } else {
            // ================= movers: LDG fp32 -> fp16 hi / lo (rows that TMA cannot address: unaligned strides) ==========
            MoverRegs ra[kPrefetch];
            auto issue = [&](MoverRegs& r, long long it) {
                int mt, net, nt;
                decode(t_begin + (int)(it / nkb), mt, net, nt);
                mover_load(r, g.A[net], g.lda, mt * BM, g.M, (int)(it % nkb) * KB, g.K, g.vecA, tid);
            };
#pragma unroll
            for (int j = 0; j < kPrefetch; ++j)
                if (j < total_it) issue(ra[j], j);
            for (long long it0 = 0; it0 < total_it; it0 += kPrefetch) {
#pragma unroll
                for (int j = 0; j < kPrefetch; ++j) {
                    const long long it = it0 + j;
                    if (it < total_it) {
                        const int sa = (int)(it % kASlots);
                        mbar_wait(&empty_a[sa], (uint32_t)(((it / kASlots) & 1) ^ 1));
                        const uint32_t slot = a_ring + sa * P::kASlot;
                        if (T == 1) { OperandRegs o1; mover_pack_bf16(ra[j], o1); mover_store_packed(o1, slot, tid); }
                        else mover_store(ra[j], slot, slot + P::kAHalf, tid);
                        fence_proxy_async();
                        warp_arrive(&full_a[sa], lane);
                        if (it + kPrefetch < total_it) issue(ra[j], it + kPrefetch);
                    }
                }
            }
        }
    } else if (warp == kProdWarp) {
        // ================= producer (one thread): weight-image bulk copies and activation TMA boxes =================
        // two independent streams polled without blocking, so that a full ring of one never holds the other back
        // (stream positions are kept as (tile, k-block, slot, phase) counters: no division in this loop)
        if (lane == 0) {
            struct Pos { int i, kb, slot; uint32_t phase; int mt, net, nt; };
            auto start = [&](Pos& p) { p.i = 0; p.kb = 0; p.slot = 0; p.phase = 1u; if (my_tiles > 0) decode(t_begin, p.mt, p.net, p.nt); };
            auto advance = [&](Pos& p, int slots) {
                if (++p.slot == slots) { p.slot = 0; p.phase ^= 1u; }
                if (++p.kb == nkb) { p.kb = 0; if (++p.i < my_tiles) decode(t_begin + p.i, p.mt, p.net, p.nt); }
            };
            Pos pb, pa;
            start(pb); start(pa);
            if (g.dbg & DBG_NO_W) pb.i = my_tiles;
            if (AMODE == A_LDG || (g.dbg & DBG_NO_A)) pa.i = my_tiles;
            uint32_t full_leader[PAIR ? kASlots : 1];              // pair: the leader's full barriers (cluster addresses)
            if (PAIR) {
#pragma unroll
                for (int s2 = 0; s2 < kASlots; ++s2) full_leader[s2] = mapa_u32(smem_u32(&full_a[s2]), 0);
            }
            const uint8_t* img0 = g.Bimg[0] + (size_t)(group % kImgReplicas) * g.img_replica_bytes;
            const uint8_t* img1 = g.Bimg[1] + (size_t)(group % kImgReplicas) * g.img_replica_bytes;
            while (pb.i < my_tiles || pa.i < my_tiles) {
                if (pb.i < my_tiles && mbar_test(PAIR ? &empty_a[pb.slot] : &empty_b[pb.slot], pb.phase)) {
                    // image of (n-tile, k-block): [hi tile (BN rows x 64 B) | lo tile]; a CTA of a pair takes its half of the rows of each
                    const uint8_t* img = (pb.net ? img1 : img0) + ((size_t)pb.nt * nkb + (size_t)pb.kb) * (size_t)(T * BN * P::kRowBytes);
                    if (PAIR) {
                        // rows of 64 B in the image tensor: replica, then (n-tile, k-block) tiles of T x BN rows; this CTA's half
                        if (rank == 0) mbar_arrive_expect_tx(&full_a[pb.slot], 2u * (uint32_t)P::kBSlot);
                        const int row0 = (int)(((size_t)(group % kImgReplicas) * g.img_replica_bytes) / P::kRowBytes) +
                                         (pb.nt * nkb + pb.kb) * (T * BN) + (int)rank * (BN / 2);
#pragma unroll
                        for (int t = 0; t < T; ++t)
                            tma_load_2d_pair(b_ring + pb.slot * P::kBSlot + t * P::kBHalf, &maps.b[pb.net], 0, row0 + t * BN, full_leader[pb.slot]);
                    } else {
                        mbar_arrive_expect_tx(&full_b[pb.slot], (uint32_t)P::kBSlot);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(b_ring + pb.slot * P::kBSlot), "l"(img), "r"((uint32_t)P::kBSlot), "r"(smem_u32(&full_b[pb.slot]))
                                     : "memory");
                    }
                    advance(pb, kBSlots);
                }
                if (AMODE == A_PLANES) {
                    if (pa.i < my_tiles && mbar_test(&empty_a[pa.slot], pa.phase)) {
                        const uint32_t slot = a_ring + pa.slot * P::kASlot;
                        if (PAIR) {
                            if (rank == 0) mbar_arrive_expect_tx(&full_a[pa.slot], 2u * (uint32_t)P::kASlot);
                            tma_load_2d_pair(slot, &maps.a[pa.net], pa.kb * KB, pa.mt * BM, full_leader[pa.slot]);
                            if (T == 2) tma_load_2d_pair(slot + P::kAHalf, &maps.al[pa.net], pa.kb * KB, pa.mt * BM, full_leader[pa.slot]);
                        } else {
                            mbar_arrive_expect_tx(&full_a[pa.slot], (uint32_t)P::kASlot);
                            tma_load_2d(slot, &maps.a[pa.net], pa.kb * KB, pa.mt * BM, &full_a[pa.slot]);
                            if (T == 2) tma_load_2d(slot + P::kAHalf, &maps.al[pa.net], pa.kb * KB, pa.mt * BM, &full_a[pa.slot]);
                        }
                        advance(pa, kASlots);
                    }
                } else if (AMODE == A_RAW) {
                    if (pa.i < my_tiles && mbar_test(&raw_empty[pa.slot], pa.phase)) {
                        mbar_arrive_expect_tx(&raw_full[pa.slot], (uint32_t)P::kRawSlot);
                        tma_load_2d(raw_ring + pa.slot * P::kRawSlot, &maps.a[pa.net], pa.kb * KB, pa.mt * BM, &raw_full[pa.slot]);
                        advance(pa, kRSlots > 0 ? kRSlots : 1);
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (one thread; in a pair: the leader issues, the peer relays its full barriers) ==========
        if (PAIR && rank != 0) {
            // (the peer issues nothing: its loads report to the leader's barriers by themselves)
        } else if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BN, PAIR ? 2 * BM : BM, T == 1);
            static_assert(!PAIR || (kASlots == kBSlots && AMODE == A_PLANES), "a pair runs one ring, fed by TMA alone");
            int sa = 0, sb = 0;
            uint32_t pha = 0u, phb = 0u;
            for (int i = 0; i < my_tiles; ++i) {
                const int as = i % kAcc;
                const uint32_t acc = tmem_base + (uint32_t)as * kStageCols;
                // the epilogue (of both CTAs of a pair) has read tile i - kAcc out of this stage
                if (PAIR) mbar_wait_cluster(&acc_empty[as], (uint32_t)(((i / kAcc) & 1) ^ 1));
                else mbar_wait(&acc_empty[as], (uint32_t)(((i / kAcc) & 1) ^ 1));
                tc_fence_after();
                if (g.timeline && blockIdx.x == 0 && i < 16) g.timeline[i * 8 + 0] = clock64();
                const bool tl = g.timeline && blockIdx.x == 0 && i < 16;      // cycles this tile's MMA thread spent waiting, per barrier kind
                long long w_a = 0, w_b = 0, w_p = 0, t0 = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (tl) t0 = clock64();
                    if (PAIR) {                                       // one barrier: both CTAs' activation and weight bytes
                        mbar_wait_cluster(&full_a[sa], pha);
                        if (tl) w_p += clock64() - t0;
                    } else {
                        if (!(AMODE == A_PLANES && (g.dbg & DBG_NO_A))) mbar_wait(&full_a[sa], pha);
                        if (tl) { const long long t1 = clock64(); w_a += t1 - t0; t0 = t1; }
                        if (!(g.dbg & DBG_NO_W)) mbar_wait(&full_b[sb], phb);
                        if (tl) { const long long t1 = clock64(); w_b += t1 - t0; t0 = t1; }
                    }
                    tc_fence_after();
                    const uint32_t a_hi = a_ring + sa * P::kASlot, a_lo = a_hi + P::kAHalf;
                    const uint32_t b_hi = b_ring + sb * P::kBSlot, b_lo = b_hi + P::kBHalf;
                    // K-major tiles: 64-byte rows / SWIZZLE_64B / 512-byte 8-row groups, or 128-byte rows / SWIZZLE_128B / 1024
                    constexpr uint32_t kSbo = KB == 64 ? 1024u : 512u, kLay = KB == 64 ? kLayoutSW128 : kLayoutSW64;
#pragma unroll
                    for (int k = 0; k < KB / 16; ++k) {
                        const uint64_t da_hi = make_sdesc(a_hi + k * 32, 0, kSbo, kLay);
                        const uint64_t da_lo = make_sdesc(a_lo + k * 32, 0, kSbo, kLay);
                        const uint64_t db_hi = make_sdesc(b_hi + k * 32, 0, kSbo, kLay);
                        const uint64_t db_lo = make_sdesc(b_lo + k * 32, 0, kSbo, kLay);
                        if (PAIR) {
                            if (T == 2 && !(g.dbg & DBG_ONE_MMA)) {
                                umma_f16_pair(acc + BN, da_lo, db_hi, idesc, (kb | k) != 0);
                                umma_f16_pair(acc + BN, da_hi, db_lo, idesc, 1u);
                            }
                            umma_f16_pair(acc, da_hi, db_hi, idesc, (kb | k) != 0);
                        } else {
                            if (T == 2 && !(g.dbg & DBG_ONE_MMA)) {
                                umma_f16(acc + BN, da_lo, db_hi, idesc, (kb | k) != 0);
                                umma_f16(acc + BN, da_hi, db_lo, idesc, 1u);
                            }
                            umma_f16(acc, da_hi, db_hi, idesc, (kb | k) != 0);
                        }
                    }
                    if (PAIR) umma_commit_pair(&empty_a[sa]);      // the k-block's slots (one ring: sa == sb) are free in BOTH CTAs
                    else { umma_commit(&empty_a[sa]); umma_commit(&empty_b[sb]); }
                    if (++sa == kASlots) { sa = 0; pha ^= 1u; }
                    if (++sb == kBSlots) { sb = 0; phb ^= 1u; }
                }
                if (PAIR) umma_commit_pair(&acc_full[as]); else umma_commit(&acc_full[as]);     // accumulators of tile i complete
                if (tl) { g.timeline[i * 8 + 1] = clock64(); g.timeline[i * 8 + 5] = w_a; g.timeline[i * 8 + 6] = w_b; g.timeline[i * 8 + 7] = w_p; }
            }
            tc_fence_before();
        }
    } else if (TMA_C) {
        // ================= epilogue warps, TMA write-out =================
        // thread = accumulator row.  A tile leaves TMEM in 64-column sub-phases; each thread finishes its 32 columns of the
        // sub-phase in registers (correction, bias, ReLU + sign word | mask word) and stages them as one 128-B row of a
        // SWIZZLE_128B box; one thread then hands the two boxes of the sub-phase to TMA (cp.async.bulk.tensor store, clipped
        // at the matrix edge).  With two staging buffers the store of sub-phase k runs under the drain of k+1, so TMEM is
        // held for the drains only.
        constexpr int kSubs = BN / 64;
        constexpr int kBufs = P::kStaging / (BM * 64 * 4);         // 2 (BN >= 128) or 1
        static_assert(kBufs == 1 || kBufs == 2, "staging holds one or two 64-column sub-phases");
        const int ew = warp - kEpiWarp0;
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        const int ch = ew >> 2;                                   // which 32 columns of the sub-phase
        const int r = q * 32 + lane;
        const bool t0 = (ew == 0 && lane == 0);
        const uint32_t row_off = (uint32_t)r * 128u;
        uint32_t gsp = 0;                                         // sub-phases staged so far (buffer choice)
        for (int i = 0; i < my_tiles; ++i) {
            int mt, net, nt;
            decode(t_begin + i, mt, net, nt);
            const int m = mt * BM + r;
            const CUtensorMap* mapC = &maps.c[net];
            const CUtensorMap* mapCl = &maps.cl[net];
            uint32_t* __restrict__ bits_out = (EPI == EPI_NN) ? g.bits_out[net] : nullptr;
            uint32_t mw[kSubs];
            if (EPI == EPI_NN) {
                // the tile's bias vector -> shared memory (everybody finished reading the previous tile's at its last barrier)
                const int e = (warp - kEpiWarp0) * 32 + lane;
                const float* __restrict__ bias = g.bias[net];
                if (e < BN) s_bias[e] = (bias && nt * BN + e < g.N) ? __ldg(bias + nt * BN + e) : 0.f;
                epi_bar();
            } else if (g.bits_in) {
#pragma unroll
                for (int sp = 0; sp < kSubs; ++sp) {
                    const int wi = nt * (BN / 32) + sp * 2 + ch;
                    mw[sp] = (m < g.M && wi < g.bits_wpr) ? __ldg(g.bits_in + (size_t)m * g.bits_wpr + wi) : 0u;
                }
            }
            const int as = i % kAcc;
            mbar_wait(&acc_full[as], (uint32_t)((i / kAcc) & 1));
            tc_fence_after();
            if (g.timeline && blockIdx.x == 0 && i < 16 && t0) g.timeline[i * 8 + 2] = clock64();
            const uint32_t acc = tmem_base + (uint32_t)as * kStageCols + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int sp = 0; sp < kSubs; ++sp, ++gsp) {
                const uint32_t buf = stg + (gsp % kBufs) * (uint32_t)(BM * 64 * 4) + (uint32_t)ch * (uint32_t)(BM * 128);
                const int col = sp * 64 + ch * 32;               // first of this thread's 32 columns within the tile
                if (kBufs == 1) {                                 // one buffer: its previous store must have read it
                    if (t0) bulk_wait_read0();
                    epi_bar();
                }
                uint32_t v[32];
                if (T == 2) {
                    uint32_t w2[32];
                    tmem_ld32x2(acc + (uint32_t)col, v, acc + (uint32_t)(BN + col), w2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(w2[j]), kLoInv, __uint_as_float(v[j])));
                } else {
                    tmem_ld32(acc + (uint32_t)col, v);
                }
                if (sp == kSubs - 1) {                            // every column of this tile is out of TMEM
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(as);
                    if (g.timeline && blockIdx.x == 0 && i < 16 && t0) g.timeline[i * 8 + 3] = clock64();
                }
                // this thread's 32 finished columns
                float f[32];
                if (EPI == EPI_NN) {
                    uint32_t sign = 0u;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + col + j4 * 4);
                        float x0 = __uint_as_float(v[j4 * 4 + 0]) + b4.x, x1 = __uint_as_float(v[j4 * 4 + 1]) + b4.y;
                        float x2 = __uint_as_float(v[j4 * 4 + 2]) + b4.z, x3 = __uint_as_float(v[j4 * 4 + 3]) + b4.w;
                        if (g.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
                        sign |= ((x0 > 0.f ? 1u : 0u) | (x1 > 0.f ? 2u : 0u) | (x2 > 0.f ? 4u : 0u) | (x3 > 0.f ? 8u : 0u)) << (j4 * 4);
                        f[j4 * 4 + 0] = x0; f[j4 * 4 + 1] = x1; f[j4 * 4 + 2] = x2; f[j4 * 4 + 3] = x3;
                    }
                    const int wi = (nt * BN + col) >> 5;
                    if (bits_out && m < g.M && wi < g.bits_wpr) bits_out[(size_t)m * g.bits_wpr + wi] = sign;
                } else {
                    uint32_t keep = 0xffffffffu;
                    if (g.bits_in) keep = mw[sp];
                    else if (g.mask) {
                        keep = 0u;
                        const int n0 = nt * BN + col;
                        if (m < g.M && n0 + 32 <= g.N) {           // (a 16-B aligned fp32 mask row)
                            const float4* mp = reinterpret_cast<const float4*>(g.mask + (size_t)m * g.ldmask + n0);
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {
                                const float4 k = __ldg(mp + j4);
                                keep |= ((k.x > 0.f ? 1u : 0u) | (k.y > 0.f ? 2u : 0u) | (k.z > 0.f ? 4u : 0u) | (k.w > 0.f ? 8u : 0u)) << (j4 * 4);
                            }
                        } else if (m < g.M) {
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < g.N && __ldg(g.mask + (size_t)m * g.ldmask + n0 + j) > 0.f) keep |= 1u << j;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = ((keep >> j) & 1u) ? __uint_as_float(v[j]) : 0.f;
                }
                if (CMODE == C_TMA) {
                    // one 128-B row of this warp half's {32 fp32 x 128 rows} SWIZZLE_128B box
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        sts128f(buf + row_off + (uint32_t)(((j4 ^ r) & 7) << 4), make_float4(f[j4 * 4 + 0], f[j4 * 4 + 1], f[j4 * 4 + 2], f[j4 * 4 + 3]));
                } else if (T == 1) {
                    // one bf16 plane: 64 B of this row in the {64 bf16 x 128 rows} SWIZZLE_128B box
                    const uint32_t hb = stg + (gsp % kBufs) * (uint32_t)(BM * 64 * 4) + row_off;
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4)
                        sts128(hb + (uint32_t)((((ch * 4 + j4) ^ r) & 7) << 4),
                               make_uint4(pack_bf16(f[j4 * 8 + 0], f[j4 * 8 + 1]), pack_bf16(f[j4 * 8 + 2], f[j4 * 8 + 3]),
                                          pack_bf16(f[j4 * 8 + 4], f[j4 * 8 + 5]), pack_bf16(f[j4 * 8 + 6], f[j4 * 8 + 7])));
                } else {
                    // fp16 planes: 64 B of this row in the hi box and in the lo box ({64 fp16 x 128 rows}, SWIZZLE_128B); the split
                    // is the one the next product's converters would have applied to the fp32 value, so results do not change
                    const uint32_t hb = stg + (gsp % kBufs) * (uint32_t)(BM * 64 * 4) + row_off, lb = hb + (uint32_t)(BM * 128);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        uint4 h, l;
                        split2(f[j4 * 8 + 0], f[j4 * 8 + 1], h.x, l.x); split2(f[j4 * 8 + 2], f[j4 * 8 + 3], h.y, l.y);
                        split2(f[j4 * 8 + 4], f[j4 * 8 + 5], h.z, l.z); split2(f[j4 * 8 + 6], f[j4 * 8 + 7], h.w, l.w);
                        const uint32_t ck = (uint32_t)((((ch * 4 + j4) ^ r) & 7) << 4);
                        sts128(hb + ck, h);
                        sts128(lb + ck, l);
                    }
                }
                fence_proxy_async();                              // staged rows -> visible to TMA (async proxy)
                if (kBufs == 2 && t0) bulk_wait_read0();         // the other buffer's store (previous sub-phase) has read it
                epi_bar();                                        // sub-phase staged; the other buffer is free
                if (t0) {
                    const uint32_t b0 = stg + (gsp % kBufs) * (uint32_t)(BM * 64 * 4);
                    const int n0 = nt * BN + sp * 64;
                    if (CMODE == C_TMA) {
                        if (n0 < g.N && !(g.dbg & DBG_NO_STORE)) tma_store_2d(mapC, b0, n0, mt * BM);
                        if (n0 + 32 < g.N && !(g.dbg & DBG_NO_STORE)) tma_store_2d(mapC, b0 + (uint32_t)(BM * 128), n0 + 32, mt * BM);
                    } else if (n0 < g.N && !(g.dbg & DBG_NO_STORE)) {
                        tma_store_2d(mapC, b0, n0, mt * BM);
                        if (T == 2) tma_store_2d(mapCl, b0 + (uint32_t)(BM * 128), n0, mt * BM);
                    }
                    bulk_commit();
                }
            }
            if (g.timeline && blockIdx.x == 0 && i < 16 && t0) g.timeline[i * 8 + 4] = clock64();
        }
        if (t0) bulk_wait_read0();                                // shared memory stays valid until TMA has read it
    } else {
        // ================= epilogue warps, st.global write-out (C rows that TMA cannot address) =================
        constexpr int PC = P::kPhaseCols;
        constexpr int kRowBytes = PC * 4;
        constexpr int kDrainCols = PC / 2;                        // columns each of the two warps of a lane quarter drains
        constexpr int kLanesPerRow = PC / 4;                      // float4 lanes covering one staged row
        constexpr int kRowsPerIter = 32 / kLanesPerRow;
        constexpr int kIters = (BM / kEpiWarps) / kRowsPerIter;   // 16 rows per warp
        const int ew = warp - kEpiWarp0;
        const int q = warp & 3;                                   // TMEM lane quarter this warp may read
        const int ch = ew >> 2;
        const int lr = lane / kLanesPerRow, cl = lane % kLanesPerRow;
        for (int i = 0; i < my_tiles; ++i) {
            int mt, net, nt;
            decode(t_begin + i, mt, net, nt);
            float* __restrict__ C = g.C[net];
            const float* __restrict__ bias = g.bias[net];
            uint32_t* __restrict__ bits_out = (EPI == EPI_NN) ? g.bits_out[net] : nullptr;
            const int as = i % kAcc;
            mbar_wait(&acc_full[as], (uint32_t)((i / kAcc) & 1));
            tc_fence_after();
            if (g.timeline && blockIdx.x == 0 && i < 16 && ew == 0 && lane == 0) g.timeline[i * 8 + 2] = clock64();
            const uint32_t acc = tmem_base + (uint32_t)as * kStageCols + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int p = 0; p < P::kPhases; ++p) {
                epi_bar();                                         // the staging buffer has been written out
                // ---- drain: thread = accumulator row; main + correction / 2^11, scaled back, to the staging tile ----
                const int r = q * 32 + lane;
#pragma unroll 1
                for (int cb = 0; cb < kDrainCols; cb += 32) {
                    const int col = p * PC + ch * kDrainCols + cb;
                    uint32_t v[32];
                    tmem_ld32(acc + (uint32_t)col, v);
                    if (T == 2) {                        // main + correction / 2^11
                        uint32_t w2[32];
                        tmem_ld32(acc + (uint32_t)(BN + col), w2);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(w2[j]), kLoInv, __uint_as_float(v[j])));
                    }
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 o = make_float4(__uint_as_float(v[j4 * 4 + 0]), __uint_as_float(v[j4 * 4 + 1]),
                                                     __uint_as_float(v[j4 * 4 + 2]), __uint_as_float(v[j4 * 4 + 3]));
                        const int c = (ch * kDrainCols + cb) / 4 + j4;               // 16-B chunk of the staged row
                        const int pc = (c & ~7) | ((c ^ r) & 7);                      // XOR swizzle: conflict-free both ways
                        sts128f(stg + (uint32_t)(r * kRowBytes + pc * 16), o);
                    }
                }
                if (p == P::kPhases - 1) {                         // every column of this tile is out of TMEM
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(as);
                    if (g.timeline && blockIdx.x == 0 && i < 16 && ew == 0 && lane == 0) g.timeline[i * 8 + 3] = clock64();
                }
                // NT: the mask words of this warp's rows are requested before the barrier
                uint32_t mw[kIters];
                if (EPI == EPI_NT && g.bits_in) {
#pragma unroll
                    for (int it = 0; it < kIters; ++it) {
                        const int m = mt * BM + ew * 16 + it * kRowsPerIter + lr;
                        const int n = nt * BN + p * PC + cl * 4;
                        mw[it] = (m < g.M && n < g.N) ? __ldg(g.bits_in + (size_t)m * g.bits_wpr + (n >> 5)) : 0u;
                    }
                }
                // the bias of this lane's four columns, likewise ahead of the barrier
                const int n = nt * BN + p * PC + cl * 4;
                const int nv = min(4, g.N - n);
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (EPI == EPI_NN && bias) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (e < nv) bv[e] = __ldg(bias + n + e);
                }
                epi_bar();                                         // staging complete
                // ---- write-out: one staged row (or two) per warp instruction, coalesced ----
#pragma unroll
                for (int it = 0; it < kIters; ++it) {
                    const int rr = ew * 16 + it * kRowsPerIter + lr;
                    const int m = mt * BM + rr;
                    const int pc = (cl & ~7) | ((cl ^ rr) & 7);
                    const float4 t4 = lds128f(stg + (uint32_t)(rr * kRowBytes + pc * 16));
                    float x[4] = {t4.x, t4.y, t4.z, t4.w};
                    uint32_t nib = 0u;
                    const bool ok = (m < g.M) && (nv > 0);
                    if (EPI == EPI_NN) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float y = x[e] + bv[e];
                            x[e] = g.relu ? fmaxf(y, 0.f) : y;
                            if (ok && e < nv && x[e] > 0.f) nib |= 1u << e;
                        }
                        if (bits_out) {                            // warp-uniform
                            uint32_t w = nib << (4 * (lane & 7));
                            w |= __shfl_xor_sync(0xffffffffu, w, 1);
                            w |= __shfl_xor_sync(0xffffffffu, w, 2);
                            w |= __shfl_xor_sync(0xffffffffu, w, 4);
                            if ((lane & 7) == 0 && m < g.M && n < g.N) bits_out[(size_t)m * g.bits_wpr + (n >> 5)] = w;
                        }
                    } else {
                        if (g.bits_in) {
                            const uint32_t b4 = mw[it] >> (4 * (lane & 7));
                            x[0] = (b4 & 1u) ? x[0] : 0.f; x[1] = (b4 & 2u) ? x[1] : 0.f;
                            x[2] = (b4 & 4u) ? x[2] : 0.f; x[3] = (b4 & 8u) ? x[3] : 0.f;
                        } else if (g.mask && ok) {
                            const float* mp = g.mask + (size_t)m * g.ldmask + n;
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (e < nv) x[e] = (__ldg(mp + e) > 0.f) ? x[e] : 0.f;
                        }
                    }
                    if (ok) {
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
            if (g.timeline && blockIdx.x == 0 && i < 16 && ew == 0 && lane == 0) g.timeline[i * 8 + 4] = clock64();
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();       // the leader's MMAs read the peer's shared memory and signal its barriers until the very end
    else __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- weight images: per (n-tile, k-block) the fp16 hi tile then the lo tile, K-major SWIZZLE_64B, exactly as the MMA
// reads them from shared memory -> one cp.async.bulk per k-block ----
struct HImgEntry { const float* src; uint8_t* dst; int N, K, ld, transposed, bn, terms, kb; };      // kb: 32 or 64 elements per k-block
struct HImgPlan { int n; HImgEntry e[3 * (kMaxLayers + 1)]; long long unit_end[3 * (kMaxLayers + 1)]; };

__global__ void pack_images_h_kernel(HImgPlan p, long long total_units) {
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total_units; u += (long long)gridDim.x * blockDim.x) {
        int i = 0;
        while (u >= p.unit_end[i]) ++i;
        const HImgEntry e = p.e[i];
        long long r_ = u - (i ? p.unit_end[i - 1] : 0);
        const int cpr = e.kb / 8;                              // 16-B chunks (8 elements) per row: 4 (64-B rows) or 8 (128-B rows)
        const int c = (int)(r_ % cpr); r_ /= cpr;
        const int r = (int)(r_ % e.bn); r_ /= e.bn;           // row within the tile
        const int kbs = (e.K + e.kb - 1) / e.kb;
        const int kb = (int)(r_ % kbs), nt = (int)(r_ / kbs);
        const int n = nt * e.bn + r, k0 = kb * e.kb + c * 8;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = k0 + q;
            v[q] = (n < e.N && k < e.K) ? (e.transposed ? __ldg(e.src + (size_t)k * e.ld + n) : __ldg(e.src + (size_t)n * e.ld + k)) : 0.f;
        }
        uint4 h, l;
        if (e.terms == 1) {
            h = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            l = h;
        } else {
            split2(v[0], v[1], h.x, l.x); split2(v[2], v[3], h.y, l.y); split2(v[4], v[5], h.z, l.z); split2(v[6], v[7], h.w, l.w);
        }
        const size_t ntiles = (size_t)((e.N + e.bn - 1) / e.bn);
        const size_t half = (size_t)e.bn * (e.kb * 2);          // a tile image = [hi (bn rows x 64 | 128 B) | lo]  (one-term mode: hi only)
        const size_t tile_bytes = (size_t)e.terms * half;
        const size_t replica_bytes = ntiles * kbs * tile_bytes;
        const uint32_t off = e.kb == 64 ? (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4)               // SWIZZLE_128B
                                        : (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) << 4);       // SWIZZLE_64B
#pragma unroll
        for (int rep = 0; rep < kImgReplicas; ++rep) {
            uint8_t* tile = e.dst + rep * replica_bytes + ((size_t)nt * kbs + kb) * tile_bytes;
            *reinterpret_cast<uint4*>(tile + off) = h;
            if (e.terms == 2) *reinterpret_cast<uint4*>(tile + half + off) = l;
        }
    }
}

// DPPO_H_TMAC=0: keep the st.global write-out even where TMA could store C (A/B switch for measurements)
static bool tmac_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DPPO_H_TMAC"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// DPPO_H_PAIR=0: keep every product on single CTAs (A/B switch for measurements; default: pairs where they apply)
static bool pair_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DPPO_H_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <int BN, int EPI, int AMODE, int CMODE, int T, bool PAIR>
static int launch_h4(const HMaps& maps, HArgs& g, cudaStream_t st) {
    using P = HPlan<BN, AMODE, T, PAIR>;
    const int tiles_m = (g.M + BM - 1) / BM;
    const long long total = (long long)(PAIR ? (tiles_m + 1) / 2 : tiles_m) * g.tiles_n * g.nets;     // (super-)tiles
    DPPO_REQUIRE(total < (1ll << 30), "h_gemm: too many tiles");
    static PerDeviceOnce configured;
    if (configured.need())
        DPPO_CHECK_CUDA(cudaFuncSetAttribute(h_gemm_kernel<BN, EPI, AMODE, CMODE, T, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kBytes));
    int cap = PAIR ? kPlanSMs / 2 : kPlanSMs;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    if (PAIR) {
        cfg.gridDim = dim3(kPlanSMs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = P::kBytes; cfg.stream = st;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        // the persistent grid = the pairs that can be resident at once (a GPC with an odd number of usable SMs leaves one unpaired)
        static int max_clusters[16] = {0};
        int dev = 0;
        DPPO_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 16) {
            if (max_clusters[dev] == 0) {
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, h_gemm_kernel<BN, EPI, AMODE, CMODE, T, PAIR>, &cfg) != cudaSuccess || n < 1) {
                    (void)cudaGetLastError();
                    n = kPlanSMs / 2;
                }
                max_clusters[dev] = n;
                if (getenv("DPPO_H_VERBOSE")) fprintf(stderr, "h_gemm pair<%d,%d,%d,%d,%d>: cudaOccupancyMaxActiveClusters = %d\n", BN, EPI, AMODE, CMODE, T, n);
            }
            if (max_clusters[dev] < cap) cap = max_clusters[dev];
        }
    }
    const int groups = (int)(total < cap ? total : cap);
    ProfScope ps(KK_TC_GEMM, st);
    if (PAIR) {
        cfg.gridDim = dim3(2 * groups);
        DPPO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, h_gemm_kernel<BN, EPI, AMODE, CMODE, T, PAIR>, maps, g, (int)total));
    } else {
        h_gemm_kernel<BN, EPI, AMODE, CMODE, T, PAIR><<<groups, kThreads, P::kBytes, st>>>(maps, g, (int)total);
        DPPO_LAUNCH_CHECK();
    }
    return 0;
}
template <int BN, int EPI, int AMODE, int CMODE, int T>
static int launch_h3(const HMaps& maps, HArgs& g, cudaStream_t st) {
    // pairs: the 256-wide tiles whose operands arrive by TMA and leave by TMA (the L2-bound products); at least two row blocks
    if (BN == 256 && AMODE == A_PLANES && CMODE != C_STG && g.M > BM && pair_enabled()) {
        HMaps m2 = maps;
        for (int n = 0; n < g.nets; ++n) {
            constexpr int KB = HKb<AMODE, T>::kKb;
            const long long rows = (long long)kImgReplicas * g.img_replica_bytes / (KB * 2);
            DPPO_REQUIRE(rows < (1ll << 31) && h_make_map16(&m2.b[n], g.Bimg[n], (int)rows, KB, KB, KB, BN / 2, CU_TENSOR_MAP_SWIZZLE_NONE),
                         "cuTensorMapEncodeTiled failed for the weight image (%lld rows)", rows);
        }
        return launch_h4<BN, EPI, AMODE, CMODE, T, (BN == 256 && AMODE == A_PLANES && CMODE != C_STG)>(m2, g, st);
    }
    return launch_h4<BN, EPI, AMODE, CMODE, T, false>(maps, g, st);
}
template <int BN, int EPI, int AMODE, int CMODE>
static int launch_h2(const HMaps& maps, HArgs& g, cudaStream_t st) {
    if (g.terms == 1) return launch_h3<BN, EPI, AMODE, CMODE, 1>(maps, g, st);
    return launch_h3<BN, EPI, AMODE, CMODE, 2>(maps, g, st);
}

template <int BN, int EPI, int AMODE>
static int launch_h1(const HMaps& maps, int cmode, HArgs& g, cudaStream_t st) {
    if (cmode == C_PLANES) return launch_h2<BN, EPI, AMODE, C_PLANES>(maps, g, st);
    if (cmode == C_TMA) return launch_h2<BN, EPI, AMODE, C_TMA>(maps, g, st);
    return launch_h2<BN, EPI, AMODE, C_STG>(maps, g, st);
}

template <int BN, int EPI>
static int launch_h(HArgs& g, const ThPlanes& pl, cudaStream_t st) {
    g.tiles_n = (g.N + BN - 1) / BN;
    HMaps maps;
    memset(&maps, 0, sizeof(maps));
    int amode = A_LDG, cmode = C_STG;
    if (pl.a_planes) {
        // fp16 hi | lo planes, K-major operand tiles of {32 fp16 x 128 rows}, SWIZZLE_64B: exactly what the MMA reads
        for (int n = 0; n < g.nets; ++n) {
            const __half* hi = reinterpret_cast<const __half*>(g.A[n]);
            DPPO_REQUIRE(g.lda % 8 == 0 && aligned16(hi) && pl.a_lo_off % 8 == 0, "h_gemm: fp16 planes need 16-byte aligned rows");
            // (one-term mode: 64-element k-blocks = 128-byte rows, SWIZZLE_128B)
            DPPO_REQUIRE((g.terms == 1 ? h_make_map16(&maps.a[n], hi, g.M, g.K, g.lda, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B)
                                       : h_make_map16(&maps.a[n], hi, g.M, g.K, g.lda, 32, BM, CU_TENSOR_MAP_SWIZZLE_64B)) &&
                         (g.terms == 1 || h_make_map16(&maps.al[n], hi + pl.a_lo_off, g.M, g.K, g.lda, 32, BM, CU_TENSOR_MAP_SWIZZLE_64B)),
                         "cuTensorMapEncodeTiled failed for the A planes [%d x %d]", g.M, g.K);
        }
        amode = A_PLANES;
    } else {
        // TMA needs 16-byte aligned rows; anything else (obs_dim 3, 17, ...) takes the LDG movers
        bool tma = g.vecA && h_make_map(&maps.a[0], g.A[0], g.M, g.K, g.lda, BM);
        if (tma && g.nets > 1) tma = h_make_map(&maps.a[1], g.A[1], g.M, g.K, g.lda, BM);
        if (tma) amode = A_RAW;
    }
    if (pl.c_planes) {
        for (int n = 0; n < g.nets; ++n) {
            __half* hi = reinterpret_cast<__half*>(g.C[n]);
            DPPO_REQUIRE(g.ldc % 8 == 0 && aligned16(hi) && pl.c_lo_off % 8 == 0, "h_gemm: fp16 planes need 16-byte aligned rows");
            DPPO_REQUIRE(h_make_map16(&maps.c[n], hi, g.M, g.N, g.ldc, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
                         (g.terms == 1 || h_make_map16(&maps.cl[n], hi + pl.c_lo_off, g.M, g.N, g.ldc, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B)),
                         "cuTensorMapEncodeTiled failed for the C planes [%d x %d]", g.M, g.N);
        }
        cmode = C_PLANES;
    } else {
        bool tmc = g.vecC && tmac_enabled() && h_make_map(&maps.c[0], g.C[0], g.M, g.N, g.ldc, BM);
        if (tmc && g.nets > 1) tmc = h_make_map(&maps.c[1], g.C[1], g.M, g.N, g.ldc, BM);
        if (tmc) cmode = C_TMA;
    }
    if (amode == A_PLANES) return launch_h1<BN, EPI, A_PLANES>(maps, cmode, g, st);
    if (amode == A_RAW) return launch_h1<BN, EPI, A_RAW>(maps, cmode, g, st);
    return launch_h1<BN, EPI, A_LDG>(maps, cmode, g, st);
}

// ---- fp32 <-> fp16 hi | lo planes (the split the converters apply; tests and one-off conversions) ----
__global__ void split_planes_kernel(const float* __restrict__ src, long long n, __half* __restrict__ hi, __half* __restrict__ lo, int terms) {
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (long long)gridDim.x * blockDim.x * 2) {
        const float a = src[i], b = (i + 1 < n) ? src[i + 1] : 0.f;
        uint32_t h, l;
        if (terms == 1) {                                    // one bf16 plane
            h = pack_bf16(a, b);
            if (i + 1 < n) *reinterpret_cast<uint32_t*>(hi + i) = h;
            else hi[i] = __ushort_as_half((unsigned short)(h & 0xffffu));
            continue;
        }
        split2(a, b, h, l);
        if (i + 1 < n) { *reinterpret_cast<uint32_t*>(hi + i) = h; *reinterpret_cast<uint32_t*>(lo + i) = l; }
        else { hi[i] = __ushort_as_half((unsigned short)(h & 0xffffu)); lo[i] = __ushort_as_half((unsigned short)(l & 0xffffu)); }
    }
}
__global__ void join_planes_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, long long n, float* __restrict__ dst, int terms) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = terms == 1 ? __uint_as_float((uint32_t)__half_as_ushort(hi[i]) << 16) : fmaf(__half2float(lo[i]), kLoInv, __half2float(hi[i]));
}

}  // namespace th

// Tile width.  256 columns halve the activation re-reads but fill TMEM (no overlap of drain and MMAs); with a short
// reduction the drain dominates such a tile, so products with K < kWideMinK take two 128-wide tiles instead, whose two TMEM
// stages hide it (DPPO_H_WIDE_MINK overrides the threshold for measurements).
static int th_wide_min_k() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DPPO_H_WIDE_MINK"); v = e ? atoi(e) : 65; if (v < 0) v = 0; }
    return v;
}
// Few row blocks (minibatches of a few hundred rows, the strong-scaling slices): the product is latency-bound, not
// throughput-bound -- 64-wide tiles give four times the CTAs and the deepest operand rings (DPPO_H_SMALL_M overrides the bound).
static int th_small_m() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DPPO_H_SMALL_M"); v = e ? atoi(e) : 2048; if (v < 0) v = 0; }
    return v;
}
static inline int th_bn_for(int n, int k, int m = 1 << 30) {
    if (m <= th_small_m()) return 64;
    return n > 128 ? (k >= th_wide_min_k() ? 256 : 128) : (n > 64 ? 128 : 64);
}

// bytes of one replica of the weight image of B[N,K] (terms = 2: fp16 hi | lo tiles, 1: bf16 tiles)
static long long th_image_replica_bytes(int N, int K, int terms = 2, int M = 1 << 30, int kb = 32) {
    const int bn = th_bn_for(N, K, M);
    return (long long)((N + bn - 1) / bn) * ((K + kb - 1) / kb) * bn * (kb * 2) * terms;
}
// (allocation: the two-term image with 32-element k-blocks, or the one-term image with 64-element ones if that pads K further)
long long th_image_floats(int N, int K) {
    const long long a = th_image_replica_bytes(N, K, 2, 1 << 30, 32), b = th_image_replica_bytes(N, K, 1, 1 << 30, 64);
    return th::kImgReplicas * (a > b ? a : b) / 4;
}
// elements per k-block of a product: 64 for the one-term mode on plane operands (HKb), else 32
static inline int th_kb_for(int terms, int a_planes) { return (terms == 1 && a_planes) ? 64 : 32; }
long long th_bits_words(long long rows, int N) { return rows * ((N + 31) / 32); }

// Every weight image of one update: forward images of pi (and oldpi) from W_l^T, dX images of pi from W_l
// (img_* are arrays of L+1 device pointers; nullptr = skip).
int th_pack_images(const NetShape& s, const float* params_pi, const float* head_pi, const float* params_old,
                   const float* head_old, float* const* img_fwd_pi, float* const* img_fwd_old, float* const* img_dx_pi,
                   cudaStream_t st, int terms, int rows, int planes) {
    th::HImgPlan p; p.n = 0;
    long long units = 0;
    auto add = [&](const float* src, float* dst, int N, int K, int ld, int transposed, int a_planes) {
        if (!src || !dst) return;
        th::HImgEntry& e = p.e[p.n];
        e.src = src; e.dst = reinterpret_cast<uint8_t*>(dst); e.N = N; e.K = K; e.ld = ld; e.transposed = transposed;
        e.bn = th_bn_for(N, K, rows); e.terms = terms == 1 ? 1 : 2; e.kb = th_kb_for(e.terms, a_planes);
        units += (long long)((N + e.bn - 1) / e.bn) * ((K + e.kb - 1) / e.kb) * e.bn * (e.kb / 8);
        p.unit_end[p.n++] = units;
    };
    for (int l = 0; l <= s.L; ++l) {
        const int in = s.in_dim(l), out = s.out_dim(l);
        const float* w_pi = (l == s.L) ? head_pi : params_pi + s.w_off[l];
        const float* w_old = (l == s.L) ? head_old : (params_old ? params_old + s.w_off[l] : nullptr);
        // (a_planes: the activation operand of the product this image feeds is stored as planes -- forward of a layer above the
        //  first, dX of a hidden layer's gradient; it decides the k-block length of the one-term mode, th_kb_for)
        // (planes == 2: the observations are a plane too -- one-term mode, capi.cu: obs_pl)
        add(w_pi, img_fwd_pi[l], out, in, out, 1, planes && (l > 0 || planes == 2));                 // forward: B[n=out][k=in] = W[k][n]
        add(params_old ? w_old : nullptr, img_fwd_old[l], out, in, out, 1, planes && (l > 0 || planes == 2));
        if (l > 0) add(w_pi, img_dx_pi[l], in, out, out, 0, planes && l < s.L);     // dX:      B[n=in][k=out] = W[n][k]
    }
    if (p.n == 0) return 0;
    int grid = (int)((units + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope ps(KK_MISC, st);
    th::pack_images_h_kernel<<<grid, 256, 0, st>>>(p, units);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// one image from a [N,K] (transposed = 0) or [K,N] (transposed = 1) matrix with row stride ld
int th_pack_one(const float* src, float* dst, int N, int K, int ld, int transposed, cudaStream_t st, int terms, int rows, int a_planes) {
    th::HImgPlan p; p.n = 1;
    th::HImgEntry& e = p.e[0];
    e.src = src; e.dst = reinterpret_cast<uint8_t*>(dst); e.N = N; e.K = K; e.ld = ld; e.transposed = transposed; e.bn = th_bn_for(N, K, rows);
    e.terms = terms == 1 ? 1 : 2; e.kb = th_kb_for(e.terms, a_planes);
    const long long units = (long long)((N + e.bn - 1) / e.bn) * ((K + e.kb - 1) / e.kb) * e.bn * (e.kb / 8);
    p.unit_end[0] = units;
    int grid = (int)((units + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    ProfScope ps(KK_MISC, st);
    th::pack_images_h_kernel<<<grid, 256, 0, st>>>(p, units);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// C[M,N] = epi(A[M,K] . B^T) with B given as fp16 hi/lo weight images (th_pack_images).  nets = 1 or 2 products sharing
// shapes.  epi_nt = 0: C = relu?(acc + bias), optional sign bitmap of C to bits_out0 (net 0 only);
// epi_nt = 1: C = acc * mask, mask = bits_in (sign bitmap) or the fp32 matrix `mask`.
// Operands must lie in fp16's range: the backward products therefore run on gradients the loss kernel pre-scaled by a
// power of two (capi.cu: grad_prescale), and finalize_kernel divides it out again.
int th_gemm(int epi_nt, int M, int N, int K, const float* A0, const float* A1, int lda, const float* img0, const float* img1,
            float* C0, float* C1, int ldc, const float* bias0, const float* bias1, int relu, const float* mask, int nets,
            cudaStream_t st, uint32_t* bits_out0, const uint32_t* bits_in, long long* timeline, const ThPlanes* planes) {

    if (M == 0) return 0;
    th::HArgs g{};
    g.A[0] = A0; g.A[1] = A1; g.Bimg[0] = reinterpret_cast<const uint8_t*>(img0); g.Bimg[1] = reinterpret_cast<const uint8_t*>(img1);
    g.C[0] = C0; g.C[1] = C1; g.bias[0] = bias0; g.bias[1] = bias1; g.mask = mask; g.bits_in = bits_in;
    g.bits_out[0] = bits_out0; g.bits_out[1] = nullptr;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldc = ldc; g.ldmask = N; g.nets = nets; g.relu = relu;
    ThPlanes pl{};
    if (planes) pl = *planes;
    g.bits_wpr = (N + 31) / 32;
    g.terms = pl.terms == 1 ? 1 : 2;
    g.img_replica_bytes = th_image_replica_bytes(N, K, g.terms, M, th_kb_for(g.terms, pl.a_planes));
    g.timeline = timeline;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("DPPO_H_DBG"); dbg = e ? atoi(e) : 0; } g.dbg = dbg; }
    g.vecA = !pl.a_planes && (lda % 4 == 0) && th::aligned16(A0) && (nets < 2 || th::aligned16(A1));
    g.vecC = !pl.c_planes && (ldc % 4 == 0) && th::aligned16(C0) && (nets < 2 || th::aligned16(C1));
    const int bn = th_bn_for(N, K, M);
    if (epi_nt) {
        if (bn == 256) return th::launch_h<256, th::EPI_NT>(g, pl, st);
        if (bn == 128) return th::launch_h<128, th::EPI_NT>(g, pl, st);
        return th::launch_h<64, th::EPI_NT>(g, pl, st);
    }
    if (bn == 256) return th::launch_h<256, th::EPI_NN>(g, pl, st);
    if (bn == 128) return th::launch_h<128, th::EPI_NN>(g, pl, st);
    return th::launch_h<64, th::EPI_NN>(g, pl, st);
}

// x[n] fp32 -> hi[n] | lo[n] fp16 planes and back (x ~= hi + lo / 2^11)
int th_split_planes(const float* src, long long n, void* hi, void* lo, cudaStream_t st, int terms) {
    if (n <= 0) return 0;
    long long blocks = (n / 2 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ProfScope ps(KK_MISC, st);
    th::split_planes_kernel<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(src, n, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), terms);
    DPPO_LAUNCH_CHECK();
    return 0;
}
int th_join_planes(const void* hi, const void* lo, long long n, float* dst, cudaStream_t st, int terms) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ProfScope ps(KK_MISC, st);
    th::join_planes_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const __half*>(hi), reinterpret_cast<const __half*>(lo), n, dst, terms);
    DPPO_LAUNCH_CHECK();
    return 0;
}

}  // namespace dppo
