// K3/K5 (exact-fp32 path): dense layers of the actor-critic MLP as SIMT FFMA GEMMs.
//
// Replaces the tf.layers.dense MatMul/BiasAdd/Relu ops (/root/reference/Model.py:12-14) and
// their auto-diff (PPO.py:46): forward  H = relu(X.W + b);  backward  dX = (dY.W^T) * [H>0],
// dW = X^T.dY, db = colsum(dY).  fp32 FMA keeps the 1e-5 parity budget (TF32 tensor cores do
// not; the tensor-core paths are separate kernels).
//
// One kernel template, 256 threads, (16*TM)x(16*TN) output tile, BK = 16, register-staged
// double buffering through shared memory (global->reg prefetch of tile t+1 overlaps the FFMAs
// of tile t; one __syncthreads per tile).  Operands may be K-contiguous (transposed on the way
// into smem) or tile-dim-contiguous (copied as float4), which covers NN / NT / TN without
// materialising a transpose:
//     forward   NN : A[M,K] K-contig,  B=W[K,N] N-contig,  epilogue +bias, relu
//     backward  NT : A=dY[M,K] K-contig, B=W[N,K] K-contig, epilogue *[H>0]
//     backward  TN : A=X[Kr,M] M-contig, B=dY[Kr,N] N-contig, split-K over rows Kr into
//                    kSplitK slices (reduced deterministically by finalize_kernel), colsum(dY)
//                    accumulated on the fly by the first row-tile of CTAs.
#include "common.cuh"

namespace dppo {

constexpr int BK = 16;
constexpr int kThreads = 256;

enum { EPI_NN = 0, EPI_NT = 1, EPI_TN = 2 };

struct GemmArgs {
    const float* A[2]; const float* B[2]; float* C[2]; const float* bias[2];
    const float* mask;      // NT
    float* bias_out;        // TN
    int M, N, K;            // C is MxN, reduction length K
    int lda, ldb, ldc;
    int tiles_n, nets;
    int relu, accumulate;
    int k_per_slice;        // TN split-K (multiple of BK); == K rounded up otherwise
    long long slice_stride; // TN
    int vecA, vecB, vecC;   // 16-byte vector access legal for the operand
};

// ---- tile loaders: R x BK tile of an operand, staged in registers ---------------------------
// K-contiguous operand: element (r,k) at P[r*ld + k]
template <int R>
__device__ __forceinline__ void load_kc(float4 (&reg)[R / 64], const float* __restrict__ P, int ld,
                                        int r0, int rdim, int k0, int kend, int vec, int tid) {
#pragma unroll
    for (int i = 0; i < R / 64; ++i) {
        const int idx = tid + i * kThreads;
        const int r = r0 + idx / 4;
        const int k = k0 + (idx % 4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rdim) {
            const float* p = P + (size_t)r * ld + k;
            if (vec && k + 3 < kend) {
                v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                if (k + 0 < kend) v.x = __ldg(p + 0);
                if (k + 1 < kend) v.y = __ldg(p + 1);
                if (k + 2 < kend) v.z = __ldg(p + 2);
                if (k + 3 < kend) v.w = __ldg(p + 3);
            }
        }
        reg[i] = v;
    }
}
template <int R>
__device__ __forceinline__ void store_kc(const float4 (&reg)[R / 64], float (*S)[R], int tid) {
#pragma unroll
    for (int i = 0; i < R / 64; ++i) {
        const int idx = tid + i * kThreads;
        const int r = idx / 4;
        const int k = (idx % 4) * 4;
        S[k + 0][r] = reg[i].x; S[k + 1][r] = reg[i].y; S[k + 2][r] = reg[i].z; S[k + 3][r] = reg[i].w;
    }
}
// tile-dim-contiguous operand: element (r,k) at P[k*ld + r]
template <int R>
__device__ __forceinline__ void load_rc(float4 (&reg)[R / 64], const float* __restrict__ P, int ld,
                                        int r0, int rdim, int k0, int kend, int vec, int tid) {
#pragma unroll
    for (int i = 0; i < R / 64; ++i) {
        const int idx = tid + i * kThreads;
        const int k = k0 + idx / (R / 4);
        const int r = r0 + (idx % (R / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend) {
            const float* p = P + (size_t)k * ld + r;
            if (vec && r + 3 < rdim) {
                v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                if (r + 0 < rdim) v.x = __ldg(p + 0);
                if (r + 1 < rdim) v.y = __ldg(p + 1);
                if (r + 2 < rdim) v.z = __ldg(p + 2);
                if (r + 3 < rdim) v.w = __ldg(p + 3);
            }
        }
        reg[i] = v;
    }
}
template <int R>
__device__ __forceinline__ void store_rc(const float4 (&reg)[R / 64], float (*S)[R], int tid) {
#pragma unroll
    for (int i = 0; i < R / 64; ++i) {
        const int idx = tid + i * kThreads;
        const int k = idx / (R / 4);
        const int r = (idx % (R / 4)) * 4;
        *reinterpret_cast<float4*>(&S[k][r]) = reg[i];
    }
}

template <int TM, int TN, bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(kThreads, 2) sgemm_kernel(const GemmArgs g) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile is built from float4 groups");
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int net = blockIdx.x / g.tiles_n;
    const int n0 = (blockIdx.x % g.tiles_n) * BN;
    const int m0 = blockIdx.y * BM;
    const int slice = blockIdx.z;
    const int k_begin = slice * g.k_per_slice;
    const int k_end = min(g.K, k_begin + g.k_per_slice);

    const float* __restrict__ A = g.A[net];
    const float* __restrict__ B = g.B[net];

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    // colsum(dY): thread `tid` of the first row-tile's CTAs owns column n0+tid of the B tile and adds it
    // up straight from shared memory, Kahan-compensated (the sum cancels heavily over rows)
    float csum = 0.f, ccomp = 0.f;
    const bool do_bias = (EPI == EPI_TN) && (g.bias_out != nullptr) && (blockIdx.y == 0) && (tid < BN);

    float4 ra[BM / 64], rb[BN / 64];
    const int ntiles = (k_end - k_begin + BK - 1) / BK;

    auto gload = [&](int t) {
        const int k0 = k_begin + t * BK;
        if (A_KC) load_kc<BM>(ra, A, g.lda, m0, g.M, k0, k_end, g.vecA, tid);
        else      load_rc<BM>(ra, A, g.lda, m0, g.M, k0, k_end, g.vecA, tid);
        if (B_KC) load_kc<BN>(rb, B, g.ldb, n0, g.N, k0, k_end, g.vecB, tid);
        else      load_rc<BN>(rb, B, g.ldb, n0, g.N, k0, k_end, g.vecB, tid);
    };
    auto sstore = [&](int buf) {
        if (A_KC) store_kc<BM>(ra, As[buf], tid); else store_rc<BM>(ra, As[buf], tid);
        if (B_KC) store_kc<BN>(rb, Bs[buf], tid); else store_rc<BN>(rb, Bs[buf], tid);
    };

    if (ntiles > 0) {
        gload(0);
        sstore(0);
    }
    __syncthreads();

    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) gload(t + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int gi = 0; gi < TM / 4; ++gi) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][gi * 64 + ty * 4]);
                a[gi * 4 + 0] = v.x; a[gi * 4 + 1] = v.y; a[gi * 4 + 2] = v.z; a[gi * 4 + 3] = v.w;
            }
#pragma unroll
            for (int gj = 0; gj < TN / 4; ++gj) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][gj * 64 + tx * 4]);
                b[gj * 4 + 0] = v.x; b[gj * 4 + 1] = v.y; b[gj * 4 + 2] = v.z; b[gj * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (EPI == EPI_TN) {
            if (do_bias) {
#pragma unroll
                for (int k = 0; k < BK; ++k) {
                    const float y = Bs[buf][k][tid] - ccomp;
                    const float t2 = csum + y;
                    ccomp = (t2 - csum) - y;
                    csum = t2;
                }
            }
        }
        if (t + 1 < ntiles) sstore(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue ----
    float* __restrict__ C = g.C[net];
    if (EPI == EPI_TN) C += (size_t)slice * g.slice_stride;
    const float* __restrict__ bias = g.bias[net];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (m >= g.M) continue;
#pragma unroll
        for (int gj = 0; gj < TN / 4; ++gj) {
            const int n = n0 + gj * 64 + tx * 4;
            if (n >= g.N) continue;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = acc[i][gj * 4 + q];
            float* cp = C + (size_t)m * g.ldc + n;
            const int nv = min(4, g.N - n);
            if (EPI == EPI_NN) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < nv) {
                        float x = v[q] + (bias ? __ldg(bias + n + q) : 0.f);
                        v[q] = g.relu ? fmaxf(x, 0.f) : x;
                    }
                }
            } else if (EPI == EPI_NT) {
                if (g.mask) {
                    const float* mp = g.mask + (size_t)m * g.ldc + n;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q < nv) v[q] = (__ldg(mp + q) > 0.f) ? v[q] : 0.f;
                }
            } else {
                if (g.accumulate) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q < nv) v[q] += cp[q];
                }
            }
            if (g.vecC && nv == 4) {
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q < nv) cp[q] = v[q];
            }
        }
    }
    if (EPI == EPI_TN) {
        if (do_bias) {
            float* bo = g.bias_out + (size_t)slice * g.slice_stride;
            const int n = n0 + tid;
            if (n < g.N) bo[n] = g.accumulate ? bo[n] + csum : csum;
        }
    }
}

static inline int aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int TM, int TN, bool A_KC, bool B_KC, int EPI>
static int launch(GemmArgs& g, int slices, cudaStream_t st) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    g.tiles_n = (g.N + BN - 1) / BN;
    const int tiles_m = (g.M + BM - 1) / BM;
    DPPO_REQUIRE(tiles_m <= 65535, "gemm: M=%d too large for one launch", g.M);
    dim3 grid(g.tiles_n * g.nets, tiles_m, slices);
    ProfScope ps(EPI == EPI_NN ? KK_GEMM_NN : (EPI == EPI_NT ? KK_GEMM_NT : KK_GEMM_TN), st);
    sgemm_kernel<TM, TN, A_KC, B_KC, EPI><<<grid, kThreads, 0, st>>>(g);
    DPPO_LAUNCH_CHECK();
    return 0;
}

static void fill_vec_flags(GemmArgs& g) {
    g.vecA = (g.lda % 4 == 0);
    g.vecB = (g.ldb % 4 == 0);
    g.vecC = (g.ldc % 4 == 0);
    for (int i = 0; i < g.nets; ++i) {
        g.vecA &= aligned16(g.A[i]);
        g.vecB &= aligned16(g.B[i]);
        g.vecC &= aligned16(g.C[i]);
    }
}

int gemm_nn_dual(int M, int N, int K, const float* A0, const float* A1, int lda, const float* B0,
                 const float* B1, int ldb, float* C0, float* C1, int ldc, const float* bias0,
                 const float* bias1, int relu, int nets, cudaStream_t st) {
    if (M == 0) return 0;
    GemmArgs g{};
    g.A[0] = A0; g.A[1] = A1; g.B[0] = B0; g.B[1] = B1; g.C[0] = C0; g.C[1] = C1;
    g.bias[0] = bias0; g.bias[1] = bias1;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.nets = nets;
    g.relu = relu; g.k_per_slice = ((K + BK - 1) / BK) * BK;
    fill_vec_flags(g);
    if (N > 64) return launch<8, 8, true, false, EPI_NN>(g, 1, st);
    return launch<8, 4, true, false, EPI_NN>(g, 1, st);
}

int gemm_nn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const GemmEpi& e, cudaStream_t st) {
    return gemm_nn_dual(M, N, K, A, nullptr, lda, B, nullptr, ldb, C, nullptr, ldc, e.bias, nullptr,
                        e.relu, 1, st);
}

int gemm_nt(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const GemmEpi& e, cudaStream_t st) {
    if (M == 0) return 0;
    GemmArgs g{};
    g.A[0] = A; g.B[0] = B; g.C[0] = C; g.mask = e.mask;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.nets = 1;
    g.k_per_slice = ((K + BK - 1) / BK) * BK;
    fill_vec_flags(g);
    if (N > 64) return launch<8, 8, true, true, EPI_NT>(g, 1, st);
    return launch<8, 4, true, true, EPI_NT>(g, 1, st);
}

int gemm_tn_splitk(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                   int ldc, const GemmEpi& e, cudaStream_t st) {
    GemmArgs g{};
    g.A[0] = A; g.B[0] = B; g.C[0] = C; g.bias_out = e.bias_out;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.nets = 1;
    g.accumulate = e.accumulate; g.slice_stride = e.slice_stride;
    const int slices = e.slices < 1 ? 1 : e.slices;
    int kps = (K + slices - 1) / slices;
    kps = ((kps + BK - 1) / BK) * BK;
    if (kps == 0) kps = BK;
    g.k_per_slice = kps;       // slices past K see an empty range and write zeros (or keep +=)
    fill_vec_flags(g);
    // slice_stride must keep 16-B alignment of every slice for the vector epilogue
    if (g.slice_stride % 4 != 0) g.vecC = 0;
    if (N > 64) return launch<8, 8, false, false, EPI_TN>(g, slices, st);
    return launch<8, 4, false, false, EPI_TN>(g, slices, st);
}

}  // namespace dppo
