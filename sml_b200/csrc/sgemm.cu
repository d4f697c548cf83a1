// Grouped fp32 SIMT GEMM: the A/B reference of the tensor-core path (SML_GEMM=simt selects it for every fc1 / fc2 forward,
// data-gradient and weight-gradient product of the steps and of the full-table transfer, model/transfer.py:463-511,701-728,
// 884-902), and a column-sum kernel for the bias gradients of that path.  Exact fp32 FFMA accumulation (no TF32 rounding),
// so results are directly comparable with the reference's fp32 cuBLAS/CPU path.  The default path is umma_packed.cu /
// umma_gemm.cu (tcgen05, 3xTF32).
//
// C[M,N] = epi( sum_k opA(A)[m,k] * opB(B)[k,n] ), 64x64x16 tiles, 256 threads, 4x4 per thread.
#include "sml_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS_ = 68, GEMM_THREADS = 256;
constexpr int MAX_PROBS = 4;

struct GemmParams {
    SmlGemmProb p[MAX_PROBS];
};

// Loads a BK x 64 tile into smem laid out [k][x] (x = m for A, n for B).
// KCONTIG: global is [x][k] with k contiguous (ld = row pitch);  else global is [k][x], x contiguous.
template <bool KCONTIG, bool GELU>
__device__ __forceinline__ void load_tile(float (*s)[LDS_], const float *__restrict__ G, int ld, int x0, int X, int k0,
                                          int K) {
    const int tid = threadIdx.x;
    if (KCONTIG) {
        const int x = tid >> 2, kq = (tid & 3) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x0 + x < X) {
            const float *src = G + (size_t)(x0 + x) * ld + k0 + kq;
            if (k0 + kq + 3 < K) v = *reinterpret_cast<const float4 *>(src);
            else {
                if (k0 + kq + 0 < K) v.x = src[0];
                if (k0 + kq + 1 < K) v.y = src[1];
                if (k0 + kq + 2 < K) v.z = src[2];
            }
            if (GELU) { v.x = sml_gelu(v.x); v.y = sml_gelu(v.y); v.z = sml_gelu(v.z); v.w = sml_gelu(v.w); }
        }
        s[kq + 0][x] = v.x; s[kq + 1][x] = v.y; s[kq + 2][x] = v.z; s[kq + 3][x] = v.w;
    } else {
        const int k = tid >> 4, xq = (tid & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + k < K) {
            const float *src = G + (size_t)(k0 + k) * ld + x0 + xq;
            if (x0 + xq + 3 < X) v = *reinterpret_cast<const float4 *>(src);
            else {
                if (x0 + xq + 0 < X) v.x = src[0];
                if (x0 + xq + 1 < X) v.y = src[1];
                if (x0 + xq + 2 < X) v.z = src[2];
            }
            if (GELU) { v.x = sml_gelu(v.x); v.y = sml_gelu(v.y); v.z = sml_gelu(v.z); v.w = sml_gelu(v.w); }
        }
        *reinterpret_cast<float4 *>(&s[k][xq]) = v;
    }
}

template <int A_MODE, int B_MODE, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS)
k_sgemm(GemmParams P) {
    __shared__ __align__(16) float As[BK][LDS_];
    __shared__ __align__(16) float Bs[BK][LDS_];
    const SmlGemmProb p = P.p[blockIdx.z];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= p.M || n0 >= p.N) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += BK) {
        load_tile<A_MODE != SML_A_KM, A_MODE == SML_A_MK_GELU>(As, p.A, p.lda, m0, p.M, k0, p.K);
        load_tile<B_MODE == SML_B_NK, B_MODE == SML_B_KN_GELU>(Bs, p.B, p.ldb, n0, p.N, k0, p.K);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (EPI == SML_EPI_BIAS) v += p.bias[n];
            if (EPI == SML_EPI_MUL_GELU_GRAD) v *= sml_gelu_grad(p.aux[(size_t)m * p.ldc + n]);
            float *c = p.C + (size_t)m * p.ldc + n;
            if (EPI == SML_EPI_ACCUM) v += *c;
            *c = v;
        }
    }
}

// out[c] += sum_r X[r, c].  Grid = (column groups of 32, problems, row slices): each CTA reduces 8 row lanes of
// its slice in a fixed order and adds one partial per column with an fp32 atomic (bias gradients).
struct ColsumParams { SmlColsumProb p[MAX_PROBS]; };
constexpr int COLSUM_SLICES = 16;

__global__ void __launch_bounds__(256) k_colsum(ColsumParams P) {
    __shared__ float s[8][33];
    const SmlColsumProb p = P.p[blockIdx.y];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    if (blockIdx.x * 32 >= p.cols) return;
    const int per = (p.rows + COLSUM_SLICES - 1) / COLSUM_SLICES;
    const int r0 = blockIdx.z * per, r1 = min(p.rows, r0 + per);
    if (r0 >= r1) return;
    const int rl = threadIdx.x >> 5;
    float a = 0.f;
    if (c < p.cols)
        for (int r = r0 + rl; r < r1; r += 8) a += p.X[(size_t)r * p.ld + c];
    s[rl][threadIdx.x & 31] = a;
    __syncthreads();
    if (rl == 0 && c < p.cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x & 31];
        atomicAdd(p.out + c, t);
    }
}

template <int A_MODE, int B_MODE>
int launch_ab(const GemmParams &P, dim3 grid, int epi, cudaStream_t st) {
    switch (epi) {
        case SML_EPI_NONE: k_sgemm<A_MODE, B_MODE, SML_EPI_NONE><<<grid, GEMM_THREADS, 0, st>>>(P); break;
        case SML_EPI_BIAS: k_sgemm<A_MODE, B_MODE, SML_EPI_BIAS><<<grid, GEMM_THREADS, 0, st>>>(P); break;
        case SML_EPI_MUL_GELU_GRAD: k_sgemm<A_MODE, B_MODE, SML_EPI_MUL_GELU_GRAD><<<grid, GEMM_THREADS, 0, st>>>(P); break;
        case SML_EPI_ACCUM: k_sgemm<A_MODE, B_MODE, SML_EPI_ACCUM><<<grid, GEMM_THREADS, 0, st>>>(P); break;
        default: sml_set_error("sgemm: bad epilogue %d", epi); return SML_E_BADARG;
    }
    return SML_OK;
}

}  // namespace

int sml_launch_sgemm(const SmlGemmProb *probs, int n_probs, int a_mode, int b_mode, int epi, cudaStream_t st) {
    SML_REQUIRE(n_probs >= 1 && n_probs <= MAX_PROBS, SML_E_BADARG, "sgemm: bad problem count %d", n_probs);
    GemmParams P;
    int maxM = 0, maxN = 0;
    for (int i = 0; i < n_probs; ++i) {
        P.p[i] = probs[i];
        if (probs[i].M > maxM) maxM = probs[i].M;
        if (probs[i].N > maxN) maxN = probs[i].N;
        SML_REQUIRE((probs[i].lda % 4) == 0 && (probs[i].ldb % 4) == 0, SML_E_BADARG, "sgemm: lda/ldb must be multiples of 4");
    }
    if (maxM == 0 || maxN == 0) return SML_OK;
    dim3 grid((maxN + BN - 1) / BN, (maxM + BM - 1) / BM, n_probs);
    int rc = SML_E_BADARG;
#define SML_AB(A_, B_) if (a_mode == A_ && b_mode == B_) rc = launch_ab<A_, B_>(P, grid, epi, st)
    SML_AB(SML_A_MK, SML_B_NK);        // fc1 forward:        Z1 = A * W1^T
    SML_AB(SML_A_MK_GELU, SML_B_NK);   // fc2 forward:        Y  = g(Z1) * W2^T
    SML_AB(SML_A_MK, SML_B_KN);        // data gradients:     dF = dY * W2,  dA = dZ1 * W1
    SML_AB(SML_A_KM, SML_B_KN);        // weight gradient:    dW1 = dZ1^T * A
    SML_AB(SML_A_KM, SML_B_KN_GELU);   // weight gradient:    dW2 = dY^T * g(Z1)
#undef SML_AB
    if (rc == SML_E_BADARG) { sml_set_error("sgemm: unsupported mode combination (%d, %d)", a_mode, b_mode); return rc; }
    if (rc) return rc;
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_launch_colsum(const SmlColsumProb *probs, int n_probs, cudaStream_t st) {
    SML_REQUIRE(n_probs >= 1 && n_probs <= MAX_PROBS, SML_E_BADARG, "colsum: bad problem count %d", n_probs);
    ColsumParams P;
    int maxc = 0;
    for (int i = 0; i < n_probs; ++i) { P.p[i] = probs[i]; if (probs[i].cols > maxc) maxc = probs[i].cols; }
    if (maxc == 0) return SML_OK;
    dim3 grid((maxc + 31) / 32, n_probs, COLSUM_SLICES);
    k_colsum<<<grid, 256, 0, st>>>(P);
    SML_LAUNCH_OK();
    return SML_OK;
}
