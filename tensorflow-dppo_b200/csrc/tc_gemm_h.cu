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
//   * the epilogue has its own 8 warps and drains TMEM -> registers -> a swizzled shared-memory staging buffer in two
//     128-column phases, releases TMEM as soon as the last column is read, and only then streams the second phase to
//     HBM: the next tile's MMAs overlap the write-out (all 512 TMEM columns are one 128x256 tile's main + correction
//     accumulators, so TMEM itself cannot be double-buffered);
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

constexpr int kPeerSlots = 4;
constexpr int kPrefetch = 3;                                 // activation k-blocks in flight per mover thread
constexpr int kImgReplicas = 8;                              // copies of every weight image (spreads the hot L2 lines)

enum { EPI_NN = 0, EPI_NT = 1 };

struct HArgs {
    const float* A[2]; const uint8_t* Bimg[2]; float* C[2]; const float* bias[2];
    const float* mask;                 // NT: fp32 mask [M,N] (row stride ldc), C *= (mask > 0); used when bits_in == nullptr
    const uint32_t* bits_in;           // NT: ReLU sign bitmap [M, bits_wpr]
    uint32_t* bits_out[2];             // NN: sign bitmap of the result (nullable per net)
    int M, N, K, lda, ldc;
    int tiles_n, nets, relu, vecA, vecC, bits_wpr;
    long long img_replica_bytes;
    long long* timeline;               // optional: clock64 stamps of CTA 0 (tools/h_microbench.py)
    int fold;                          // fold the correction accumulator into the main one with tensor-core MMAs (tc_h.cuh)
};

template <int BN, bool PAIR>
struct HPlan {
    static constexpr int kBSlots = PAIR ? 4 : 3;
    // raw (TMA) ring depth: whatever shared memory is left -- narrow tiles have short k-blocks and need more boxes in flight
    static constexpr int kASlots = PAIR ? 3 : (BN >= 256 ? 2 : (BN >= 128 ? 3 : 4));   // operand ring: latency-bound when k-blocks are short
    static constexpr int kRSlots = PAIR ? 3 : (BN >= 256 ? 2 : (BN >= 128 ? 4 : 6));
    static constexpr int kAHalf = BM * 64;                   // hi (or lo) activation tile: 128 rows x 64 B
    static constexpr int kASlot = 2 * kAHalf;
    static constexpr int kBRows = PAIR ? BN / 2 : BN;        // a CTA of a pair holds half of the weight tile's rows
    static constexpr int kBHalf = kBRows * 64;
    static constexpr int kBSlot = 2 * kBHalf;
    static constexpr int kPhaseCols = BN < 128 ? BN : 128;   // columns staged per epilogue phase
    static constexpr int kPhases = BN / kPhaseCols;
    static constexpr int kStaging = BM * kPhaseCols * 4;
    static constexpr int kRawSlot = BM * 128;                // raw fp32 k-block: 128 rows x 32 floats
    static constexpr int kRawOff = kASlots * kASlot;
    static constexpr int kBOff = kRawOff + kRSlots * kRawSlot;
    static constexpr int kStgOff = kBOff + kBSlots * kBSlot;
    static constexpr int kBarOff = kStgOff + kStaging;
    static constexpr int kFoldOff = kBarOff + 256;            // the two fold patterns (tc_h.cuh)
    static constexpr int kBytes = kBarOff + 256 + kFoldBytes + 1024;
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

__device__ __forceinline__ void mover_store(const MoverRegs& r, uint32_t hi_tile, uint32_t lo_tile, int tid) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int row = (tid >> 2) + j * 64, c = tid & 3;
        const uint32_t off = (uint32_t)row * 64u + (uint32_t)((c ^ ((row >> 1) & 3)) << 4);     // SWIZZLE_64B
        const float4 a = r.v[j][0], b = r.v[j][1];
        uint4 h, l;
        split2(a.x, a.y, h.x, l.x); split2(a.z, a.w, h.y, l.y);
        split2(b.x, b.y, h.z, l.z); split2(b.z, b.w, h.w, l.w);
        sts128(hi_tile + off, h);
        sts128(lo_tile + off, l);
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

// PAIR: two CTAs of a cluster work on a 256-row super-tile with cta_group::2 MMAs (M = 256): each CTA streams its own 128
// activation rows and HALF of the weight tile's rows, so the weight bytes a CTA pulls through its L2 port halve (the
// single-CTA kernel is bound by exactly that port).  The leader (rank 0) issues every MMA once both CTAs' slots are full
// (the peer's MMA warp relays its local full barriers to the leader) and releases slots / publishes accumulators with
// multicast commits; each CTA drains its own TMEM rows.
template <int BN, int EPI, bool TMA_A, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
h_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1, const HArgs g, int total_tiles) {
    using P = HPlan<BN, PAIR>;
    constexpr int kBSlots = P::kBSlots, kASlots = P::kASlots, kRSlots = P::kRSlots;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
    uint64_t* acc_empty = acc_full + 1;
    uint64_t* peer_full = acc_empty + 1;            // PAIR, leader only: "the peer's slots of iteration it are full"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full + kPeerSlots);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = (g.K + kKB - 1) / kKB;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int n_groups = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;         // a group = one CTA, or one CTA pair
    const int group = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
    const int per_cta = (total_tiles + n_groups - 1) / n_groups;
    const int t_begin = group * per_cta;
    const int t_end = min(total_tiles, t_begin + per_cta);
    const int my_tiles = max(0, t_end - t_begin);
    const long long total_it = (long long)my_tiles * nkb;

    if (tid == 0) {
        for (int s = 0; s < kASlots; ++s) { mbar_init(&full_a[s], kMoverThreads); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < kBSlots; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < kRSlots; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kMoverThreads); }
        mbar_init(acc_full, 1); mbar_init(acc_empty, PAIR ? 2 * kEpiWarps : kEpiWarps);
        for (int s = 0; s < kPeerSlots; ++s) mbar_init(&peer_full[s], 1);
        fence_barrier_init();
    }
    if (g.fold) { fold_init(smem + P::kFoldOff, tid); fence_proxy_async(); }
    if (PAIR) cluster_sync_all();                   // both CTAs' barriers exist before anything arrives on them remotely
    constexpr uint32_t kTmemCols = 2 * BN;          // [0,BN) main (hi.hi), [BN,2BN) correction (lo.hi + hi.lo, scaled 2^11)
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

    if (warp < kMoverWarps) {
        if (TMA_A) {
            // ================= converters: raw fp32 tile (TMA) -> fp16 hi / lo operand tiles =================
            for (long long it = 0; it < total_it; ++it) {
                const int rs = (int)(it % kRSlots), sa = (int)(it % kASlots);
                MoverRegs r;
                mbar_wait(&raw_full[rs], (uint32_t)((it / kRSlots) & 1));
                raw_load(r, raw_ring + rs * P::kRawSlot, tid);
                mbar_arrive(&raw_empty[rs]);               // release: the loads above are complete; TMA may refill the slot
                mbar_wait(&empty_a[sa], (uint32_t)(((it / kASlots) & 1) ^ 1));
                const uint32_t slot = a_ring + sa * P::kASlot;
                mover_store(r, slot, slot + P::kAHalf, tid);
                fence_proxy_async();                       // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(&full_a[sa]);
            }
        } else {
            // ================= movers: LDG fp32 -> fp16 hi / lo (rows that TMA cannot address: unaligned strides) ==========
            MoverRegs ra[kPrefetch];
            auto issue = [&](MoverRegs& r, long long it) {
                int mt, net, nt;
                decode(t_begin + (int)(it / nkb), mt, net, nt);
                mover_load(r, g.A[net], g.lda, mt * BM, g.M, (int)(it % nkb) * kKB, g.K, g.vecA, tid);
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
                        mover_store(ra[j], slot, slot + P::kAHalf, tid);
                        fence_proxy_async();
                        mbar_arrive(&full_a[sa]);
                        if (it + kPrefetch < total_it) issue(ra[j], it + kPrefetch);
                    }
                }
            }
        }
    } else if (warp == kProdWarp) {
        // ================= producer (one thread): weight-image bulk copies and activation TMA boxes =================
        // two independent streams polled without blocking, so that a full ring of one never holds the other back
        if (lane == 0) {
            long long ib = 0, ia = TMA_A ? 0 : total_it;
            while (ib < total_it || ia < total_it) {
                if (ib < total_it) {
                    const int sb = (int)(ib % kBSlots);
                    if (mbar_test(&empty_b[sb], (uint32_t)(((ib / kBSlots) & 1) ^ 1))) {
                        int mt, net, nt;
                        decode(t_begin + (int)(ib / nkb), mt, net, nt);
                        const int kb = (int)(ib % nkb);
                        const uint8_t* img = g.Bimg[net] + (size_t)(group % kImgReplicas) * g.img_replica_bytes +
                                             (((size_t)nt * nkb + (size_t)kb) * (PAIR ? 2 : 1) + (PAIR ? rank : 0u)) * P::kBSlot;
                        mbar_arrive_expect_tx(&full_b[sb], (uint32_t)P::kBSlot);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(b_ring + sb * P::kBSlot), "l"(img), "r"((uint32_t)P::kBSlot), "r"(smem_u32(&full_b[sb]))
                                     : "memory");
                        ++ib;
                    }
                }
                if (ia < total_it) {
                    const int rs = (int)(ia % kRSlots);
                    if (mbar_test(&raw_empty[rs], (uint32_t)(((ia / kRSlots) & 1) ^ 1))) {
                        int mt, net, nt;
                        decode(t_begin + (int)(ia / nkb), mt, net, nt);
                        mbar_arrive_expect_tx(&raw_full[rs], (uint32_t)P::kRawSlot);
                        tma_load_2d(raw_ring + rs * P::kRawSlot, net ? &mapA1 : &mapA0, (int)(ia % nkb) * kKB, mt * BM, &raw_full[rs]);
                        ++ia;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (one thread; in a pair: the leader issues, the peer relays its full barriers) ==========
        if (lane == 0 && PAIR && rank != 0) {
            uint32_t remote[kPeerSlots];
#pragma unroll
            for (int s = 0; s < kPeerSlots; ++s) remote[s] = mapa_u32(smem_u32(&peer_full[s]), 0);
            for (long long it = 0; it < total_it; ++it) {
                const int sa = (int)(it % kASlots), sb = (int)(it % kBSlots);
                mbar_wait(&full_a[sa], (uint32_t)((it / kASlots) & 1));
                mbar_wait(&full_b[sb], (uint32_t)((it / kBSlots) & 1));
                mbar_arrive_remote(remote[it % kPeerSlots]);
            }
        } else if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BN, PAIR ? 2 * BM : BM);
            const uint32_t a_base = a_ring, b_base = b_ring;
            long long it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                // the epilogue (of both CTAs) has read tile i-1 out of TMEM
                if (PAIR) mbar_wait_cluster(acc_empty, (uint32_t)((i & 1) ^ 1)); else mbar_wait(acc_empty, (uint32_t)((i & 1) ^ 1));
                tc_fence_after();
                if (g.timeline && blockIdx.x == 0 && i < 16) g.timeline[i * 8 + 0] = clock64();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int sa = (int)(it % kASlots), sb = (int)(it % kBSlots);
                    mbar_wait(&full_a[sa], (uint32_t)((it / kASlots) & 1));
                    mbar_wait(&full_b[sb], (uint32_t)((it / kBSlots) & 1));
                    if (PAIR) mbar_wait_cluster(&peer_full[it % kPeerSlots], (uint32_t)((it / kPeerSlots) & 1));
                    tc_fence_after();
                    const uint32_t a_hi = a_base + sa * P::kASlot, a_lo = a_hi + P::kAHalf;
                    const uint32_t b_hi = b_base + sb * P::kBSlot, b_lo = b_hi + P::kBHalf;
#pragma unroll
                    for (int k = 0; k < kKB / 16; ++k) {
                        const uint64_t da_hi = make_sdesc(a_hi + k * 32, 0, 512, kLayoutSW64);
                        const uint64_t da_lo = make_sdesc(a_lo + k * 32, 0, 512, kLayoutSW64);
                        const uint64_t db_hi = make_sdesc(b_hi + k * 32, 0, 512, kLayoutSW64);
                        const uint64_t db_lo = make_sdesc(b_lo + k * 32, 0, 512, kLayoutSW64);
                        if (PAIR) {
                            umma_f16_pair(tmem_base + BN, da_lo, db_hi, idesc, (kb | k) != 0);
                            umma_f16_pair(tmem_base + BN, da_hi, db_lo, idesc, 1u);
                            umma_f16_pair(tmem_base, da_hi, db_hi, idesc, (kb | k) != 0);
                        } else {
                            umma_f16(tmem_base + BN, da_lo, db_hi, idesc, (kb | k) != 0);
                            umma_f16(tmem_base + BN, da_hi, db_lo, idesc, 1u);
                            umma_f16(tmem_base, da_hi, db_hi, idesc, (kb | k) != 0);
                        }
                    }
                    if (PAIR) { umma_commit_pair(&empty_a[sa]); umma_commit_pair(&empty_b[sb]); }     // slots free in BOTH CTAs
                    else { umma_commit(&empty_a[sa]); umma_commit(&empty_b[sb]); }
                }
                if (!PAIR && g.fold) fold_issue<BN>(tmem_base, smem_base + P::kFoldOff);
                if (PAIR) umma_commit_pair(acc_full); else umma_commit(acc_full);     // accumulators of tile i complete
                if (g.timeline && blockIdx.x == 0 && i < 16) g.timeline[i * 8 + 1] = clock64();
            }
            tc_fence_before();
        }
    } else {
        // ================= epilogue warps =================
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
            mbar_wait(acc_full, (uint32_t)(i & 1));
            tc_fence_after();
            if (g.timeline && blockIdx.x == 0 && i < 16 && ew == 0 && lane == 0) g.timeline[i * 8 + 2] = clock64();
            const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16);
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
                    if (!g.fold) {                                  // warp-uniform: main + correction / 2^11 here, not on the tensor core
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
                    if (lane == 0) {
                        if (PAIR && rank != 0) mbar_arrive_remote(mapa_u32(smem_u32(acc_empty), 0));
                        else mbar_arrive(acc_empty);
                    }
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
                            const float* mp = g.mask + (size_t)m * g.ldc + n;
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
struct HImgEntry { const float* src; uint8_t* dst; int N, K, ld, transposed, bn, pair; };
struct HImgPlan { int n; HImgEntry e[3 * (kMaxLayers + 1)]; long long unit_end[3 * (kMaxLayers + 1)]; };

__global__ void pack_images_h_kernel(HImgPlan p, long long total_units) {
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total_units; u += (long long)gridDim.x * blockDim.x) {
        int i = 0;
        while (u >= p.unit_end[i]) ++i;
        const HImgEntry e = p.e[i];
        long long r_ = u - (i ? p.unit_end[i - 1] : 0);
        const int c = (int)(r_ % 4); r_ /= 4;                  // 16-B chunk (8 fp16) of the 64-B row
        const int r = (int)(r_ % e.bn); r_ /= e.bn;           // row within the tile
        const int kbs = (e.K + kKB - 1) / kKB;
        const int kb = (int)(r_ % kbs), nt = (int)(r_ / kbs);
        const int n = nt * e.bn + r, k0 = kb * kKB + c * 8;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = k0 + q;
            v[q] = (n < e.N && k < e.K) ? (e.transposed ? __ldg(e.src + (size_t)k * e.ld + n) : __ldg(e.src + (size_t)n * e.ld + k)) : 0.f;
        }
        uint4 h, l;
        split2(v[0], v[1], h.x, l.x); split2(v[2], v[3], h.y, l.y); split2(v[4], v[5], h.z, l.z); split2(v[6], v[7], h.w, l.w);
        const size_t ntiles = (size_t)((e.N + e.bn - 1) / e.bn);
        // single CTA: [hi (bn rows) | lo];  CTA pair: [rank 0: hi (bn/2 rows) | lo] [rank 1: hi | lo]
        const int rows_per = e.pair ? e.bn / 2 : e.bn;
        const int part = r / rows_per, rr = r % rows_per;
        const size_t half = (size_t)rows_per * 64;
        const size_t replica_bytes = ntiles * kbs * 2 * (size_t)e.bn * 64;
        const uint32_t off = (uint32_t)rr * 64u + (uint32_t)((c ^ ((rr >> 1) & 3)) << 4);
#pragma unroll
        for (int rep = 0; rep < kImgReplicas; ++rep) {
            uint8_t* tile = e.dst + rep * replica_bytes + ((size_t)nt * kbs + kb) * 2 * (size_t)e.bn * 64 + (size_t)part * 2 * half;
            *reinterpret_cast<uint4*>(tile + off) = h;
            *reinterpret_cast<uint4*>(tile + half + off) = l;
        }
    }
}

// DPPO_H_PAIR=1: every product on CTA pairs (cta_group::2); the weight images are then packed in the pair layout
static bool pair_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DPPO_H_PAIR"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <int BN, int EPI, bool TMA_A, bool PAIR>
static int launch_h2(const CUtensorMap& m0, const CUtensorMap& m1, HArgs& g, cudaStream_t st) {
    using P = HPlan<BN, PAIR>;
    const int tiles_m = (g.M + BM - 1) / BM;
    const long long total = (long long)(PAIR ? (tiles_m + 1) / 2 : tiles_m) * g.tiles_n * g.nets;     // (super-)tiles
    DPPO_REQUIRE(total < (1ll << 30), "h_gemm: too many tiles");
    static PerDeviceOnce configured;
    if (configured.need())
        DPPO_CHECK_CUDA(cudaFuncSetAttribute(h_gemm_kernel<BN, EPI, TMA_A, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kBytes));
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = P::kBytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // persistent grid = what is co-resident: pairs must sit in one GPC, so fewer than #SMs / 2 of them may fit
    static int max_groups[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (max_groups[dev] == 0) {
        int n = PAIR ? kPlanSMs / 2 : kPlanSMs;
        if (PAIR) {
            cfg.gridDim = dim3(kPlanSMs);
            int q = 0;
            if (cudaOccupancyMaxActiveClusters(&q, h_gemm_kernel<BN, EPI, TMA_A, PAIR>, &cfg) == cudaSuccess && q > 0 && q < n) n = q;
            else (void)cudaGetLastError();
        }
        max_groups[dev] = n;
    }
    const int groups = (int)(total < max_groups[dev] ? total : max_groups[dev]);
    cfg.gridDim = dim3(PAIR ? 2 * groups : groups);
    ProfScope ps(KK_TC_GEMM, st);
    DPPO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, h_gemm_kernel<BN, EPI, TMA_A, PAIR>, m0, m1, g, (int)total));
    return 0;
}

template <int BN, int EPI>
static int launch_h(HArgs& g, bool pair, cudaStream_t st) {
    g.tiles_n = (g.N + BN - 1) / BN;
    // TMA needs 16-byte aligned rows; anything else (obs_dim 3, 17, ...) takes the LDG movers
    CUtensorMap m0, m1;
    memset(&m0, 0, sizeof(m0)); memset(&m1, 0, sizeof(m1));
    bool tma = g.vecA && h_make_map(&m0, g.A[0], g.M, g.K, g.lda, BM);
    if (tma) { if (g.nets > 1) tma = h_make_map(&m1, g.A[1], g.M, g.K, g.lda, BM); else m1 = m0; }
    if (pair) {
        if (tma) return launch_h2<BN, EPI, true, true>(m0, m1, g, st);
        return launch_h2<BN, EPI, false, true>(m0, m1, g, st);
    }
    if (tma) return launch_h2<BN, EPI, true, false>(m0, m1, g, st);
    return launch_h2<BN, EPI, false, false>(m0, m1, g, st);
}

}  // namespace th

static inline int th_bn_for(int n) { return n > 128 ? 256 : (n > 64 ? 128 : 64); }

// bytes of one replica of the weight image of B[N,K]
static long long th_image_replica_bytes(int N, int K) {
    const int bn = th_bn_for(N);
    return (long long)((N + bn - 1) / bn) * ((K + th::kKB - 1) / th::kKB) * bn * 64 * 2;
}
long long th_image_floats(int N, int K) { return th::kImgReplicas * th_image_replica_bytes(N, K) / 4; }
long long th_bits_words(long long rows, int N) { return rows * ((N + 31) / 32); }

// Every weight image of one update: forward images of pi (and oldpi) from W_l^T, dX images of pi from W_l
// (img_* are arrays of L+1 device pointers; nullptr = skip).
int th_pack_images(const NetShape& s, const float* params_pi, const float* head_pi, const float* params_old,
                   const float* head_old, float* const* img_fwd_pi, float* const* img_fwd_old, float* const* img_dx_pi,
                   cudaStream_t st) {
    th::HImgPlan p; p.n = 0;
    long long units = 0;
    auto add = [&](const float* src, float* dst, int N, int K, int ld, int transposed) {
        if (!src || !dst) return;
        th::HImgEntry& e = p.e[p.n];
        e.src = src; e.dst = reinterpret_cast<uint8_t*>(dst); e.N = N; e.K = K; e.ld = ld; e.transposed = transposed;
        e.bn = th_bn_for(N); e.pair = th::pair_enabled() ? 1 : 0;
        units += (long long)((N + e.bn - 1) / e.bn) * ((K + th::kKB - 1) / th::kKB) * e.bn * 4;
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
    th::pack_images_h_kernel<<<grid, 256, 0, st>>>(p, units);
    DPPO_LAUNCH_CHECK();
    return 0;
}

// one image from a [N,K] (transposed = 0) or [K,N] (transposed = 1) matrix with row stride ld
int th_pack_one(const float* src, float* dst, int N, int K, int ld, int transposed, cudaStream_t st) {
    th::HImgPlan p; p.n = 1;
    th::HImgEntry& e = p.e[0];
    e.src = src; e.dst = reinterpret_cast<uint8_t*>(dst); e.N = N; e.K = K; e.ld = ld; e.transposed = transposed; e.bn = th_bn_for(N);
    e.pair = th::pair_enabled() ? 1 : 0;
    const long long units = (long long)((N + e.bn - 1) / e.bn) * ((K + th::kKB - 1) / th::kKB) * e.bn * 4;
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
            cudaStream_t st, uint32_t* bits_out0, const uint32_t* bits_in, long long* timeline) {

    if (M == 0) return 0;
    th::HArgs g{};
    g.A[0] = A0; g.A[1] = A1; g.Bimg[0] = reinterpret_cast<const uint8_t*>(img0); g.Bimg[1] = reinterpret_cast<const uint8_t*>(img1);
    g.C[0] = C0; g.C[1] = C1; g.bias[0] = bias0; g.bias[1] = bias1; g.mask = mask; g.bits_in = bits_in;
    g.bits_out[0] = bits_out0; g.bits_out[1] = nullptr;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldc = ldc; g.nets = nets; g.relu = relu;
    g.bits_wpr = (N + 31) / 32;
    g.img_replica_bytes = th_image_replica_bytes(N, K);
    g.timeline = timeline;
    g.fold = (th::fold_enabled() && !th::pair_enabled()) ? 1 : 0;
    g.vecA = (lda % 4 == 0) && th::aligned16(A0) && (nets < 2 || th::aligned16(A1));
    g.vecC = (ldc % 4 == 0) && th::aligned16(C0) && (nets < 2 || th::aligned16(C1));
    const int bn = th_bn_for(N);
    const bool pair = th::pair_enabled();         // the weight images were packed for the same choice (th_pack_images)
    if (epi_nt) {
        if (bn == 256) return th::launch_h<256, th::EPI_NT>(g, pair, st);
        if (bn == 128) return th::launch_h<128, th::EPI_NT>(g, pair, st);
        return th::launch_h<64, th::EPI_NT>(g, pair, st);
    }
    if (bn == 256) return th::launch_h<256, th::EPI_NN>(g, pair, st);
    if (bn == 128) return th::launch_h<128, th::EPI_NN>(g, pair, st);
    return th::launch_h<64, th::EPI_NN>(g, pair, st);
}

}  // namespace dppo
