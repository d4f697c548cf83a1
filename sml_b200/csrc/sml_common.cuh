// Shared device helpers and internal launcher prototypes for libsml_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "sml_b200.h"

#define SML_GELU_ALPHA 1.702f

void sml_set_error(const char *fmt, ...);
int sml_check_device();
void sml_note_launch();   // bumps the kernel-launch counter read by sml_launch_count()

#define SML_CUDA_OK(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            sml_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SML_E_CUDA;                                                                   \
        }                                                                                        \
    } while (0)

#define SML_LAUNCH_OK()                                                                          \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            sml_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SML_E_CUDA;                                                                   \
        }                                                                                        \
        sml_note_launch();                                                                       \
    } while (0)

#define SML_REQUIRE(cond, code, ...)                                                             \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            sml_set_error(__VA_ARGS__);                                                          \
            return (code);                                                                       \
        }                                                                                        \
    } while (0)

// x * sigmoid(1.702 x)  (model/conv_transfer.py:9-10).  expf (not __expf): fp32-accurate.
__device__ __forceinline__ float sml_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float sml_gelu(float x) { return x * sml_sigmoid(SML_GELU_ALPHA * x); }
__device__ __forceinline__ float sml_gelu_grad(float x) {
    float s = sml_sigmoid(SML_GELU_ALPHA * x);
    return s + x * SML_GELU_ALPHA * s * (1.0f - s);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline size_t sml_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- internal launchers (defined across the .cu files) ----------------------------------

// One group of rows that share a net and a pair of source tables.
struct SmlRowGroup {
    const float *x_t;    // table the x_t rows come from
    const float *x_hat;  // table the x_hat rows come from
    const int64_t *ids;  // [n] row ids, or null for identity
    const float *theta;  // this group's net
    int64_t n;           // rows in the group
    int64_t row0;        // first row of the group in the packed [N, .] workspace matrices
};

// conv prologue: A[N,320] (fc1 input) for every row of the groups; rowsq[n] = sum x_hat^2 (or null)
int sml_launch_conv_fwd(const SmlRowGroup *groups, int n_groups, int variant, float *A, float *rowsq, cudaStream_t st);

// conv backward.  dA [N,320].  mode 0: scatter (dx_hat + l2*x_hat) into g_tab rows (atomic);
// mode 1: write dx_hat to d_rows [N,64];  theta_grad != null: accumulate conv1/conv2 grads.
struct SmlConvBwdGroup {
    SmlRowGroup g;
    float *g_tab;      // dense gradient table for scatter (mode 0) or null
    float *g_theta;    // this group's net gradient block or null
};
int sml_launch_conv_bwd(const SmlConvBwdGroup *groups, int n_groups, int variant, const float *dA, float l2,
                        float *d_rows, cudaStream_t st);

// Generic SIMT fp32 GEMM (grouped over blockIdx.z):  C[M,N] = epi(opA(A)[M,K] * opB(B)[K,N])
enum { SML_A_MK = 0, SML_A_MK_GELU = 1, SML_A_KM = 2, SML_A_KM_GELU = 3 };   // A stored [m][k] / + gelu on load / [k][m] / + gelu
enum { SML_B_NK = 0, SML_B_KN = 1, SML_B_KN_GELU = 2 };         // B stored [n][k] / [k][n] / [k][n] + gelu on load
enum { SML_EPI_NONE = 0, SML_EPI_BIAS = 1, SML_EPI_MUL_GELU_GRAD = 2, SML_EPI_ACCUM = 3 };
struct SmlGemmProb {
    const float *A, *B, *bias, *aux;
    float *C;
    int M, N, K, lda, ldb, ldc;
};
int sml_launch_sgemm(const SmlGemmProb *probs, int n_probs, int a_mode, int b_mode, int epi, cudaStream_t st);
// tcgen05 3xTF32 GEMM, same contract (+ transposed store: C[n][m]); bn = 64 | 128
int sml_launch_umma_gemm(const SmlGemmProb *probs, int n_probs, int a_mode, int b_mode, int epi, int transpose_out, int bn,
                         cudaStream_t st);
// 1 = tensor-core GEMMs (default), 0 = SIMT fp32 GEMMs (SML_GEMM=simt in the environment, for A/B comparisons)
int sml_use_tensor_cores();

// column sums: out[c] (+)= sum_r X[r, c]   (bias gradients)
struct SmlColsumProb { const float *X; float *out; int rows, cols, ld; };
int sml_launch_colsum(const SmlColsumProb *probs, int n_probs, cudaStream_t st);

// loss + dY
int sml_launch_loss(const float *Y, const float *rowsq, int64_t B, int loss_kind, int normalize_user, float l2,
                    float *dY, float *scores, float *loss_out, float *partials, unsigned int *ticket,
                    cudaStream_t st);

int sml_launch_row_normalize(float *Y, int64_t n, cudaStream_t st);
