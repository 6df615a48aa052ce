// Persistent fp16-split dW product:  C[s][M,N] (+)= X[rows_s, M]^T . G[rows_s, N]  over the rows of slice s  (sm_100a)
//
// The weight-gradient part of TF's auto-diff (/root/reference/PPO.py:46) for the dense layers of Model.py:12-14, on
// tcgen05 kind::f16 with the 2-term fp16 split of tc_gemm_h.cu (x = hi + lo / 2^11, three MMAs per product, main and
// correction accumulators in TMEM).  Both operands are activations whose reduction dimension (rollout rows) is the slow
// axis of a row-major matrix, i.e. MN-major UMMA operands (SWIZZLE_128B: 64 fp16 of M or N per 128-B row, 8 k-rows per
// atom, the layout tc_gemm_bf16.cu validated).  CTAs are persistent over (slice, m-tile, n-tile) work items:
//
//   warp 9 (one thread) : cp.async.bulk.tensor.2d boxes {32 fp32 features x 32 rows, SWIZZLE_128B} of X and of G into a
//                         raw ring (2 slots x 48 KB at BN = 256);
//   warps 0-7           : converters, lane = k-row, a fixed set of 8-feature column groups per warp: LDS raw -> fp16 hi/lo
//                         -> operand tiles (2 slots); the same pass adds up G's columns (the bias gradient) in registers --
//                         every column group belongs to ONE warp, so a warp shuffle finishes the sum, deterministically;
//   warp 8 (one thread) : 6 MMAs per 32-row k-block (2 k-steps x {lo.hi, hi.lo -> correction; hi.hi -> main});
//   warps 10-17         : epilogue, as in tc_gemm_h.cu: TMEM -> staging (64-column phases) -> coalesced partial-slice rows
//                         (+= when the update accumulates over row chunks); TMEM is released after the last phase's drain.
// G carries gradients the loss kernel pre-scaled by a power of two (capi.cu: grad_prescale); finalize_kernel divides it
// out after the split-K sum.  Rows per slice are bounded by the split-K plan (fp32 TMEM accumulation truncates: DESIGN.md).
#include "tc_h.cuh"
#include <stdlib.h>
#include <string.h>

namespace dppo {
namespace th {

constexpr int kDwPC = 64;                          // staged columns per epilogue phase

struct DwArgs {
    float* C; float* bias_out;
    int M, N, K;                  // C is [M,N] (ldc); K = rollout rows
    int ldc, k_per_slice, slices, accumulate, vecC;
    int tiles_m, tiles_n;
    long long slice_stride;
};

// XP / GP: the X / G operand is stored as fp16 hi | lo planes (written by the forward / dX epilogues of tc_gemm_h.cu): TMA
// boxes {64 fp16 features x 32 rows, SWIZZLE_128B} ARE the MN-major operand atoms, so that operand needs no raw slot and no
// converter pass; the bias gradient (column sums of G) is then read back from the operand tiles by the otherwise idle warps.
// rows per k-block: 32; the one-term mode on two plane operands takes 64 (its two MMAs per 32 rows left the per-k-block costs --
// barrier round trip, commit, twelve TMA boxes -- in charge)
template <bool XP, bool GP, int T> struct DwKb { static constexpr int kRows = (T == 1 && XP && GP) ? 64 : 32; };

template <int BN, bool XP, bool GP, int T = 2>
struct DwPlan {
    static constexpr int kRows = DwKb<XP, GP, T>::kRows;
    static constexpr int kARaw = XP ? 0 : 32 * BM * 4;   // 4 boxes of {32 features x 32 rows} fp32
    static constexpr int kBRaw = GP ? 0 : 32 * BN * 4;
    static constexpr int kRawSlot = kARaw + kBRaw;
    static constexpr int kAHalf = kRows * BM * 2;        // fp16 hi (or lo) operand tile: BM/64 atoms x kRows k-rows x 128 B
    static constexpr int kBHalf = kRows * BN * 2;
    static constexpr int kAtom = kRows * 128;            // bytes of one 64-feature atom column (the MN stride of the UMMA descriptor)
    static constexpr int kOpSlot = T * (kAHalf + kBHalf);           // [X hi | X lo | G hi | G lo] (one-term mode: [X | G], bf16)
    static constexpr int kPlaneBytes = (XP ? T * kAHalf : 0) + (GP ? T * kBHalf : 0);   // what TMA writes into an operand slot
    static constexpr int kPC = BN < kDwPC ? BN : kDwPC;
    static constexpr int kPhases = BN / kPC;
    static constexpr int kStaging = BM * kPC * 4;
    // ring depth (raw and operand rings alike): what shared memory allows, at most 4
    static constexpr int kRoom = 232448 - 256 - kStaging;
    static constexpr int kFit = kRoom / (kOpSlot + kRawSlot);
    static constexpr int kDepth = kFit > 4 ? 4 : kFit;
    static_assert(kDepth >= 2, "two slots per ring at least");
    static constexpr int kDwRaw = kRawSlot ? kDepth : 0;
    static constexpr int kDwOp = kDepth;
    static constexpr int kOpOff = kDepth * kRawSlot;
    static constexpr int kStgOff = kOpOff + kDwOp * kOpSlot;
    static constexpr int kBarOff = kStgOff + kStaging;
    static constexpr int kBytes = kBarOff + 256;
    static_assert(kBytes <= 232448, "exceeds the 227 KB of shared memory one CTA may use");
};

// D = f32, A = B = fp16 (or bf16), both MN-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int n, bool bf16 = false) {
    return (1u << 4) | (bf16 ? ((1u << 7) | (1u << 10)) : 0u) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

// one 8-feature unit (k-row `k`, box `box`, unit `j` of the box): raw fp32 -> hi / lo chunk of the MN-major operand tile
template <int T>
__device__ __forceinline__ void dw_convert_unit(uint32_t raw_op, uint32_t hi_tile, uint32_t lo_tile, int box, int j, int k,
                                                float (&csum)[8], bool sum) {
    const uint32_t rbase = raw_op + (uint32_t)box * 4096u + (uint32_t)k * 128u;
    const float4 a = lds128f(rbase + (uint32_t)(((2 * j) ^ (k & 7)) << 4));
    const float4 b = lds128f(rbase + (uint32_t)(((2 * j + 1) ^ (k & 7)) << 4));
    if (sum) {
        csum[0] += a.x; csum[1] += a.y; csum[2] += a.z; csum[3] += a.w;
        csum[4] += b.x; csum[5] += b.y; csum[6] += b.z; csum[7] += b.w;
    }
    const int c = (box & 1) * 4 + j;                                  // 16-B chunk of the 64-feature atom
    const uint32_t off = (uint32_t)(box >> 1) * 4096u + (uint32_t)k * 128u + (uint32_t)((c ^ (k & 7)) << 4);
    if (T == 1) {
        sts128(hi_tile + off, make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w)));
        return;
    }
    uint4 h, l;
    split2(a.x, a.y, h.x, l.x); split2(a.z, a.w, h.y, l.y);
    split2(b.x, b.y, h.z, l.z); split2(b.z, b.w, h.w, l.w);
    sts128(hi_tile + off, h);
    sts128(lo_tile + off, l);
}

struct DwMaps { CUtensorMap x, xl, g, gl; };          // fp32 matrix or hi plane, lo plane

template <int BN, bool XP, bool GP, int T>
__global__ void __launch_bounds__(kThreads, 1)
h_dw_kernel(const __grid_constant__ DwMaps maps, const DwArgs g, int total_tiles) {
    using P = DwPlan<BN, XP, GP, T>;
    constexpr int kDwRaw = P::kDwRaw > 0 ? P::kDwRaw : 1, kDwOp = P::kDwOp;
    constexpr bool kConv = !(XP && GP);                  // some operand still arrives as fp32 and is converted here
    constexpr int KR = P::kRows;                         // rows per k-block
    constexpr uint32_t kAtom = (uint32_t)P::kAtom;
    extern __shared__ __align__(1024) uint8_t smem[];         // SWIZZLE_128B boxes and UMMA tiles need 1024-B alignment
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t raw_ring = smem_base, op_ring = smem_base + P::kOpOff, stg = smem_base + P::kStgOff;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem + P::kBarOff);
    uint64_t* raw_empty = raw_full + kDwRaw;
    uint64_t* op_full = raw_empty + kDwRaw;
    static_assert((2 * kDwRaw + 2 * kDwOp + 2) * 8 + 4 <= 256, "barrier block");
    uint64_t* op_empty = op_full + kDwOp;
    uint64_t* acc_full = op_empty + kDwOp;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Work items are dealt out round-robin (CTA b takes b, b + grid, ...): the m-tiles of one slice are neighbours in the list,
    // so they run AT THE SAME TIME on neighbouring CTAs and the G rows they share are fetched from HBM once and found in L2 by
    // the others.  (Contiguous chunks per CTA, the first version, had one CTA walk a slice's m-tiles one after the other, a full
    // item apart: measured DRAM reads of 662 MB where the operands are 331 MB, and the kernel sat at HBM bandwidth.)
    const int t_begin = blockIdx.x;
    const int t_stride = gridDim.x;
    const int my_tiles = total_tiles > t_begin ? (total_tiles - t_begin + t_stride - 1) / t_stride : 0;
    const int tiles_mn = g.tiles_m * g.tiles_n;

    // work item t -> (slice, m-tile, n-tile); the tiles of one slice are neighbours (they share X / G rows through L2)
    auto decode = [&](int t, int& slice, int& mt, int& nt, int& nkb, int& k0) {
        slice = t / tiles_mn;
        const int r = t - slice * tiles_mn;
        mt = r / g.tiles_n; nt = r - mt * g.tiles_n;
        k0 = slice * g.k_per_slice;
        const int k1 = min(g.K, k0 + g.k_per_slice);
        nkb = max(0, (k1 - k0 + KR - 1) / KR);
    };

    if (tid == 0) {
        if (smem_base & 1023u) __trap();
        for (int s = 0; s < kDwRaw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kMoverWarps); }
        // an operand slot is full when the converters have stored their part AND the plane boxes have landed; it is empty when
        // its MMAs have completed AND (planes of G) the column-sum readers are done with it
        for (int s = 0; s < kDwOp; ++s) {
            mbar_init(&op_full[s], (kConv ? kMoverWarps : 0) + ((XP || GP) ? 1 : 0));
            mbar_init(&op_empty[s], 1 + (GP ? kMoverWarps : 0));
        }
        mbar_init(acc_full, 1); mbar_init(acc_empty, kEpiWarps);
        fence_barrier_init();
    }
    constexpr uint32_t kTmemCols = (T * BN) < 32 ? 32 : T * BN;
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kMoverWarps) {
        // ================= converters (+ column sums of G = the bias gradient) =================
        constexpr int kAPairs = (BM / 32) * 4 / kMoverWarps;          // (box, j) column groups of X per warp: 2
        constexpr int kBPairs = (BN / 32) * 4 / kMoverWarps;          // ... of G per warp: BN / 64
        static_assert(kBPairs >= 1, "BN >= 64");
        const int k = lane;
        long long it = 0;
        for (int i = 0; i < my_tiles; ++i) {
            int slice, mt, nt, nkb, k0;
            decode(t_begin + i * t_stride, slice, mt, nt, nkb, k0);
            const bool want_bias = (g.bias_out != nullptr) && (mt == 0);
            float csum[kBPairs][8];
#pragma unroll
            for (int u = 0; u < kBPairs; ++u)
#pragma unroll
                for (int e = 0; e < 8; ++e) csum[u][e] = 0.f;
            float dummy[8];
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int rs = (int)(it % kDwRaw), os = (int)(it % kDwOp);
                const uint32_t raw = raw_ring + rs * P::kRawSlot;
                const uint32_t op = op_ring + os * P::kOpSlot;
                const uint32_t a_hi = op, a_lo = op + P::kAHalf, b_hi = op + T * P::kAHalf, b_lo = b_hi + P::kBHalf;
                if (kConv) {
                    mbar_wait(&raw_full[rs], (uint32_t)((it / kDwRaw) & 1));
                    mbar_wait(&op_empty[os], (uint32_t)(((it / kDwOp) & 1) ^ 1));
                    if (!XP) {
#pragma unroll
                        for (int u = 0; u < kAPairs; ++u) {
                            const int p = warp + u * kMoverWarps;
                            dw_convert_unit<T>(raw, a_hi, a_lo, p >> 2, p & 3, k, dummy, false);
                        }
                    }
                    if (!GP) {
#pragma unroll
                        for (int u = 0; u < kBPairs; ++u) {
                            const int p = warp + u * kMoverWarps;
                            dw_convert_unit<T>(raw + P::kARaw, b_hi, b_lo, p >> 2, p & 3, k, csum[u], want_bias);
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&op_full[os]); mbar_arrive(&raw_empty[rs]); }
                }
                if (GP) {
                    // G arrived as planes: its column sums are read back from the operand tiles (lane = k-row, the same
                    // column groups as above).  Every warp passes through the slot's full barrier before it releases the slot,
                    // so that no warp can run a whole ring ahead of the barrier phases.
                    mbar_wait(&op_full[os], (uint32_t)((it / kDwOp) & 1));
                    if (want_bias) {
#pragma unroll
                        for (int hr = 0; hr < KR / 32; ++hr)
#pragma unroll
                        for (int u = 0; u < kBPairs; ++u) {
                            const int p = warp + u * kMoverWarps;          // columns [8p, 8p + 8): atom p / 8, chunk p % 8
                            const int kr = k + hr * 32;                    // k-row of this lane in this half of the k-block
                            const uint32_t off = (uint32_t)(p >> 3) * kAtom + (uint32_t)kr * 128u + (uint32_t)((((p & 7) ^ kr) & 7) << 4);
                            const uint4 h = lds128(b_hi + off);
                            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
                            if (T == 1) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    csum[u][2 * e] += __uint_as_float(hw[e] << 16);
                                    csum[u][2 * e + 1] += __uint_as_float(hw[e] & 0xffff0000u);
                                }
                            } else {
                                const uint4 l = lds128(b_lo + off);
                                const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                                    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                                    csum[u][2 * e] += fmaf(lf.x, kLoInv, hf.x);
                                    csum[u][2 * e + 1] += fmaf(lf.y, kLoInv, hf.y);
                                }
                            }
                        }
                        fence_proxy_async();       // the refill of this slot is an async-proxy write behind these generic reads
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&op_empty[os]);
                }
            }
            if (want_bias) {            // warp-uniform: fold the 32 k-rows (lanes), lane 0 writes the slice's bias partials
#pragma unroll
                for (int u = 0; u < kBPairs; ++u) {
                    const int p = warp + u * kMoverWarps;
                    const int n = nt * BN + (p >> 2) * 32 + (p & 3) * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float v = csum[u][e];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        if (lane == 0 && n + e < g.N) {
                            float* bo = g.bias_out + (size_t)slice * g.slice_stride + n + e;
                            *bo = g.accumulate ? *bo + v : v;
                        }
                    }
                }
            }
        }
    } else if (warp == kProdWarp) {
        // ================= TMA producer (one thread): raw fp32 boxes and / or plane boxes, polled without blocking =========
        if (lane == 0) {
            struct Cursor { int i, kb, nkb, slice, mt, nt, k0; long long it; bool done; };
            auto start = [&](Cursor& c) {
                c.i = 0; c.kb = 0; c.it = 0; c.done = (my_tiles == 0); c.nkb = 0;
                while (!c.done) {                                   // first work item that holds rows
                    decode(t_begin + c.i * t_stride, c.slice, c.mt, c.nt, c.nkb, c.k0);
                    if (c.nkb > 0) break;
                    if (++c.i >= my_tiles) c.done = true;
                }
            };
            auto advance = [&](Cursor& c) {
                ++c.it;
                if (++c.kb < c.nkb) return;
                c.kb = 0;
                for (;;) {
                    if (++c.i >= my_tiles) { c.done = true; return; }
                    decode(t_begin + c.i * t_stride, c.slice, c.mt, c.nt, c.nkb, c.k0);
                    if (c.nkb > 0) return;
                }
            };
            Cursor cr, cp;
            start(cr); start(cp);
            if (!kConv) cr.done = true;
            if (!(XP || GP)) cp.done = true;
            while (!cr.done || !cp.done) {
                if (!cr.done) {
                    const int rs = (int)(cr.it % kDwRaw);
                    if (mbar_test(&raw_empty[rs], (uint32_t)(((cr.it / kDwRaw) & 1) ^ 1))) {
                        const uint32_t raw = raw_ring + rs * P::kRawSlot;
                        mbar_arrive_expect_tx(&raw_full[rs], (uint32_t)P::kRawSlot);
                        const int row = cr.k0 + cr.kb * KR;
                        if (!XP) {
#pragma unroll
                            for (int b = 0; b < BM / 32; ++b) tma_load_2d(raw + b * 4096, &maps.x, cr.mt * BM + b * 32, row, &raw_full[rs]);
                        }
                        if (!GP) {
#pragma unroll
                            for (int b = 0; b < BN / 32; ++b)
                                tma_load_2d(raw + P::kARaw + b * 4096, &maps.g, cr.nt * BN + b * 32, row, &raw_full[rs]);
                        }
                        advance(cr);
                    }
                }
                if (!cp.done) {
                    const int os = (int)(cp.it % kDwOp);
                    if (mbar_test(&op_empty[os], (uint32_t)(((cp.it / kDwOp) & 1) ^ 1))) {
                        const uint32_t op = op_ring + os * P::kOpSlot;
                        const uint32_t a_hi = op, a_lo = op + P::kAHalf, b_hi = op + T * P::kAHalf, b_lo = b_hi + P::kBHalf;
                        mbar_arrive_expect_tx(&op_full[os], (uint32_t)P::kPlaneBytes);
                        const int row = cp.k0 + cp.kb * KR;
                        if (XP) {
#pragma unroll
                            for (int b = 0; b < BM / 64; ++b) {
                                tma_load_2d(a_hi + b * kAtom, &maps.x, cp.mt * BM + b * 64, row, &op_full[os]);
                                if (T == 2) tma_load_2d(a_lo + b * kAtom, &maps.xl, cp.mt * BM + b * 64, row, &op_full[os]);
                            }
                        }
                        if (GP) {
#pragma unroll
                            for (int b = 0; b < BN / 64; ++b) {
                                tma_load_2d(b_hi + b * kAtom, &maps.g, cp.nt * BN + b * 64, row, &op_full[os]);
                                if (T == 2) tma_load_2d(b_lo + b * kAtom, &maps.gl, cp.nt * BN + b * 64, row, &op_full[os]);
                            }
                        }
                        advance(cp);
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16_mn(BN, T == 1);
            constexpr uint32_t kLbo = kAtom, kSbo = 1024u, kStep = 2048u;     // MN atoms / 8-row k groups / one UMMA_K = 16 rows
            long long it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                int slice, mt, nt, nkb, k0;
                decode(t_begin + i * t_stride, slice, mt, nt, nkb, k0);
                mbar_wait(acc_empty, (uint32_t)((i & 1) ^ 1));
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int os = (int)(it % kDwOp);
                    mbar_wait(&op_full[os], (uint32_t)((it / kDwOp) & 1));
                    tc_fence_after();
                    const uint32_t op = op_ring + os * P::kOpSlot;
                    const uint32_t a_hi = op, a_lo = op + P::kAHalf, b_hi = op + T * P::kAHalf, b_lo = b_hi + P::kBHalf;
#pragma unroll
                    for (int k = 0; k < KR / 16; ++k) {
                        const uint64_t da_hi = make_sdesc(a_hi + k * kStep, kLbo, kSbo, kLayoutSW128);
                        const uint64_t da_lo = make_sdesc(a_lo + k * kStep, kLbo, kSbo, kLayoutSW128);
                        const uint64_t db_hi = make_sdesc(b_hi + k * kStep, kLbo, kSbo, kLayoutSW128);
                        const uint64_t db_lo = make_sdesc(b_lo + k * kStep, kLbo, kSbo, kLayoutSW128);
                        if (T == 2) {
                            umma_f16(tmem_base + BN, da_lo, db_hi, idesc, (kb | k) != 0);
                            umma_f16(tmem_base + BN, da_hi, db_lo, idesc, 1u);
                        }
                        umma_f16(tmem_base, da_hi, db_hi, idesc, (kb | k) != 0);
                    }
                    umma_commit(&op_empty[os]);
                }
                umma_commit(acc_full);                 // fires at once when the slice holds no rows (nkb == 0)
            }
            tc_fence_before();
        }
    } else {
        // ================= epilogue warps =================
        constexpr int PC = P::kPC;
        constexpr int kRowBytes = PC * 4;
        constexpr int kDrainCols = PC / 2;                        // 32 (or 16 when BN = 32: not built)
        constexpr int kLanesPerRow = PC / 4;                      // 16
        constexpr int kRowsPerIter = 32 / kLanesPerRow;           // 2
        constexpr int kIters = (BM / kEpiWarps) / kRowsPerIter;   // 8
        static_assert(kDrainCols == 32, "one 32-column TMEM load per warp and phase");
        const int ew = warp - kEpiWarp0;
        const int q = warp & 3;
        const int ch = ew >> 2;
        const int lr = lane / kLanesPerRow, cl = lane % kLanesPerRow;
        for (int i = 0; i < my_tiles; ++i) {
            int slice, mt, nt, nkb, k0;
            decode(t_begin + i * t_stride, slice, mt, nt, nkb, k0);
            float* __restrict__ C = g.C + (size_t)slice * g.slice_stride;
            mbar_wait(acc_full, (uint32_t)(i & 1));
            tc_fence_after();
            const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int p = 0; p < P::kPhases; ++p) {
                epi_bar();
                const int r = q * 32 + lane;
                {
                    const int col = p * PC + ch * kDrainCols;
                    uint32_t v[32];
                    if (nkb > 0) {
                        tmem_ld32(acc + (uint32_t)col, v);
                        if (T == 2) {
                            uint32_t w2[32];
                            tmem_ld32(acc + (uint32_t)(BN + col), w2);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(w2[j]), kLoInv, __uint_as_float(v[j])));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0u;
                    }
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 o = make_float4(__uint_as_float(v[j4 * 4 + 0]), __uint_as_float(v[j4 * 4 + 1]),
                                                     __uint_as_float(v[j4 * 4 + 2]), __uint_as_float(v[j4 * 4 + 3]));
                        const int c = (ch * kDrainCols) / 4 + j4;
                        const int pc = (c & ~7) | ((c ^ r) & 7);
                        sts128f(stg + (uint32_t)(r * kRowBytes + pc * 16), o);
                    }
                }
                if (p == P::kPhases - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty);
                }
                const int n = nt * BN + p * PC + cl * 4;
                const int nv = min(4, g.N - n);
                // accumulate over row chunks: the partial tile is requested ahead of the barrier
                float4 aux[kIters];
#pragma unroll
                for (int it = 0; it < kIters; ++it) {
                    const int m = mt * BM + ew * 16 + it * kRowsPerIter + lr;
                    aux[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g.accumulate && m < g.M && nv > 0) {
                        const float* sp = C + (size_t)m * g.ldc + n;
                        if (g.vecC && nv == 4) aux[it] = *reinterpret_cast<const float4*>(sp);
                        else {
                            aux[it].x = sp[0];
                            if (nv > 1) aux[it].y = sp[1];
                            if (nv > 2) aux[it].z = sp[2];
                            if (nv > 3) aux[it].w = sp[3];
                        }
                    }
                }
                epi_bar();
#pragma unroll
                for (int it = 0; it < kIters; ++it) {
                    const int rr = ew * 16 + it * kRowsPerIter + lr;
                    const int m = mt * BM + rr;
                    const int pc = (cl & ~7) | ((cl ^ rr) & 7);
                    const float4 t4 = lds128f(stg + (uint32_t)(rr * kRowBytes + pc * 16));
                    if (m < g.M && nv > 0) {
                        const float x[4] = {t4.x + aux[it].x, t4.y + aux[it].y, t4.z + aux[it].z, t4.w + aux[it].w};
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int BN, bool XP, bool GP, int T>
static int launch_dw(const DwMaps& maps, DwArgs& g, cudaStream_t st) {
    using P = DwPlan<BN, XP, GP, T>;
    g.tiles_m = (g.M + BM - 1) / BM;
    g.tiles_n = (g.N + BN - 1) / BN;
    const long long total = (long long)g.tiles_m * g.tiles_n * g.slices;
    DPPO_REQUIRE(total < (1ll << 30), "h_dw: too many work items");
    static PerDeviceOnce configured;
    if (configured.need())
        DPPO_CHECK_CUDA(cudaFuncSetAttribute(h_dw_kernel<BN, XP, GP, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kBytes));
    const int grid = (int)(total < kPlanSMs ? total : kPlanSMs);
    ProfScope ps(KK_TC_GEMM, st);
    h_dw_kernel<BN, XP, GP, T><<<grid, kThreads, P::kBytes, st>>>(maps, g, (int)total);
    DPPO_LAUNCH_CHECK();
    return 0;
}
template <int BN, int T>
static int launch_dw_t(const DwMaps& maps, int xp, int gp, DwArgs& g, cudaStream_t st) {
    if (xp && gp) return launch_dw<BN, true, true, T>(maps, g, st);
    if (xp) return launch_dw<BN, true, false, T>(maps, g, st);
    if (gp) return launch_dw<BN, false, true, T>(maps, g, st);
    return launch_dw<BN, false, false, T>(maps, g, st);
}
template <int BN>
static int launch_dw_bn(const DwMaps& maps, int xp, int gp, int terms, DwArgs& g, cudaStream_t st) {
    if (terms == 1) return launch_dw_t<BN, 1>(maps, xp, gp, g, st);
    return launch_dw_t<BN, 2>(maps, xp, gp, g, st);
}

}  // namespace th

// true when the TMA-fed dW kernel can take this product (16-byte aligned rows of both operands)
bool th_dw_supported(const float* X, int ldx, const float* G, int ldg) {
    return th::h_encode() != nullptr && (ldx % 4 == 0) && (ldg % 4 == 0) && th::aligned16(X) && th::aligned16(G);
}

// C[s][M,N] (+)= X[rows,M]^T . G[rows,N] over the rows of slice s (k_per_slice rows, a multiple of 32);
// bias_out[s][N] (+)= column sums of G (may be null).
int th_gemm_dw(int M, int N, int K, const float* X, int ldx, const float* G, int ldg, float* C, int ldc, float* bias_out,
               int slices, long long slice_stride, int accumulate, cudaStream_t st, const ThPlanes* planes) {
    th::DwMaps maps;
    memset(&maps, 0, sizeof(maps));
    const int xp = planes && planes->a_planes, gp = planes && planes->c_planes;
    const int terms = (planes && planes->terms == 1) ? 1 : 2;
    if (xp) {
        const __half* hi = reinterpret_cast<const __half*>(X);
        DPPO_REQUIRE(ldx % 8 == 0 && th::aligned16(hi) && planes->a_lo_off % 8 == 0, "h_dw: fp16 planes need 16-byte aligned rows");
        DPPO_REQUIRE(th::h_make_map16(&maps.x, hi, K, M, ldx, 64, (terms == 1 && gp) ? 64 : 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
                     (terms == 1 || th::h_make_map16(&maps.xl, hi + planes->a_lo_off, K, M, ldx, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)),
                     "cuTensorMapEncodeTiled failed for the X planes [%d x %d] ld %d", K, M, ldx);
    } else {
        DPPO_REQUIRE(th::h_make_map(&maps.x, X, K, M, ldx, 32), "cuTensorMapEncodeTiled failed for X [%d x %d] ld %d", K, M, ldx);
    }
    if (gp) {
        const __half* hi = reinterpret_cast<const __half*>(G);
        DPPO_REQUIRE(ldg % 8 == 0 && th::aligned16(hi) && planes->c_lo_off % 8 == 0, "h_dw: fp16 planes need 16-byte aligned rows");
        DPPO_REQUIRE(th::h_make_map16(&maps.g, hi, K, N, ldg, 64, (terms == 1 && xp) ? 64 : 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
                     (terms == 1 || th::h_make_map16(&maps.gl, hi + planes->c_lo_off, K, N, ldg, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)),
                     "cuTensorMapEncodeTiled failed for the G planes [%d x %d] ld %d", K, N, ldg);
    } else {
        DPPO_REQUIRE(th::h_make_map(&maps.g, G, K, N, ldg, 32), "cuTensorMapEncodeTiled failed for G [%d x %d] ld %d", K, N, ldg);
    }
    th::DwArgs g{};
    g.C = C; g.bias_out = bias_out; g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.accumulate = accumulate;
    g.slice_stride = slice_stride;
    if (slices < 1) slices = 1;
    int kps = (K + slices - 1) / slices;
    const int kr = (terms == 1 && xp && gp) ? 64 : 32;          // rows per k-block (th::DwKb)
    kps = ((kps + kr - 1) / kr) * kr;
    if (kps == 0) kps = kr;
    g.k_per_slice = kps; g.slices = slices;
    g.vecC = (ldc % 4 == 0) && th::aligned16(C) && (slice_stride % 4 == 0);
    // few rows (minibatches): latency-bound -- narrow tiles give more work items and the deepest rings
    static int small_k = -1;
    if (small_k < 0) { const char* e = getenv("DPPO_H_SMALL_M"); small_k = e ? atoi(e) : 2048; if (small_k < 0) small_k = 0; }
    if (N > 128 && K > small_k) return th::launch_dw_bn<256>(maps, xp, gp, terms, g, st);
    if (N > 64 && K > small_k) return th::launch_dw_bn<128>(maps, xp, gp, terms, g, st);
    return th::launch_dw_bn<64>(maps, xp, gp, terms, g, st);
}

}  // namespace dppo
