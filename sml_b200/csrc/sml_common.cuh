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

// sigmoid for the loss (accurate expf + IEEE division)
__device__ __forceinline__ float sml_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
// GELU(x) = x * sigmoid(1.702 x)  (model/conv_transfer.py:9-10) and its derivative.  The transfer network
// evaluates 15 of these per (row, latent dim) in the conv stage and 512 per row after fc1, so they use
// the two SFU approximations ex2.approx and rcp.approx (relative error ~2^-22 each) instead of expf + IEEE
// division: ~8 instructions instead of ~28.  Measured end-to-end effect on the forward: < 1e-6 relative
// (tests/test_gpu_parity.py tolerances are 1e-5).
__device__ __forceinline__ float sml_sig1702(float x) {
    // 1 / (1 + 2^(-1.702 * log2(e) * x))
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-2.4554669595930157f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
__device__ __forceinline__ float sml_gelu(float x) { return x * sml_sig1702(x); }
__device__ __forceinline__ float sml_gelu_grad(float x) {
    const float s = sml_sig1702(x);
    return s + x * SML_GELU_ALPHA * s * (1.0f - s);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline size_t sml_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A step is a chain of ~10 short dependent kernels; with PDL the next kernel's CTAs are scheduled while the previous
// kernel is still running and do their private prologue (barrier init, TMEM allocation, parameter loads) before
// blocking in griddepcontrol.wait until the previous grid has completed and flushed.  Every kernel of the chain calls
// sml_pdl_wait() before it touches global data and sml_pdl_trigger() as early as possible.  SML_PDL=0 disables it.
int sml_use_pdl();
#ifdef __CUDACC__
__device__ __forceinline__ void sml_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void sml_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t sml_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = sml_use_pdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif

// ---- Adam scalars (torch.optim.Adam, model/transfer.py:392-393) ------------------------------
// One tick: t += 1, step_size = lr / (1 - b1^t) and sqrt(1 - b2^t) in double, rounded to float; recorded in the
// history ring when the state carries one (see sml_adam_tick in the header).
#ifdef __CUDACC__
__device__ __forceinline__ void sml_adam_tick_body(int64_t *state, double lr, double beta1, double beta2) {
    const int64_t t = state[0] + 1;
    state[0] = t;
    const float ss = (float)(lr / (1.0 - pow(beta1, (double)t)));
    const float bs = (float)sqrt(1.0 - pow(beta2, (double)t));
    float *f = reinterpret_cast<float *>(state + 1);
    f[0] = ss;
    f[1] = bs;
    if (state[2] != 0) {
        float *h = reinterpret_cast<float *>(state + 4 + (t & (SML_ADAM_HISTORY - 1)));
        h[0] = ss;
        h[1] = bs;
    }
}

// One element of one Adam step; shared by the dense sweep and the row-lazy replay so both round identically.
// sqrt and the two divisions of torch's formula (denom = sqrt(v) / sqrt(bc2) + eps; p -= step_size * m / denom) use the SFU
// approximations (sqrt.approx / rcp.approx, <= 2 ulp each) instead of the IEEE sequences: the row-lazy replay executes this
// function once per element and MISSED step (catch-up before every batch, flush after every epoch: 28 % of the MF epochs with the
// IEEE forms, ~40 instructions per element-step against ~12).  Effect on an update: <= ~5e-7 relative, i.e. <= 5e-9 absolute per
// step at lr = 0.01 -- four orders of magnitude inside the 1e-4-after-a-period tolerance, the same trade as sml_gelu above.
// (.ftz: the reciprocals only ever see sqrt(bc2) in (0.03, 1] and denom >= eps = 1e-8; a subnormal v flushes to sqrt = 0 against eps.)
__device__ __forceinline__ float sml_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sml_sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void sml_adam1(float &p, float &m, float &v, float g, float b1c, float beta2, float b2c,
                                          float step_size, float bc2_sqrt, float eps, float wd) {
    if (wd != 0.f) g = fmaf(wd, p, g);                 // grad.add(param, alpha=weight_decay)
    m = fmaf(g - m, b1c, m);                           // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(b2c * g, g, beta2 * v);                   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = fmaf(sml_sqrt_approx(v), sml_rcp_approx(bc2_sqrt), eps);
    p = p - step_size * (m * sml_rcp_approx(denom));   // param.addcdiv_(exp_avg, denom, value=-step_size)
}
#endif

// ---- internal launchers (defined across the .cu files) ----------------------------------

// Row-lazy Adam over up to three id lists (user ids -> user table, positive / negative item ids -> item table).
struct SmlAdamRows {
    float *p, *m, *v, *g;
    int32_t *stamp;
    const int64_t *ids;
    int64_t n;
};
int sml_launch_adam_rows(const SmlAdamRows *rows, int n_groups, const int64_t *state, int apply, double beta1, double beta2,
                         double eps, cudaStream_t st);

// One group of rows that share a net and a pair of source tables.
struct SmlRowGroup {
    const float *x_t;    // table the x_t rows come from
    const float *x_hat;  // table the x_hat rows come from
    const int64_t *ids;  // [n] row ids, or null for identity
    const float *theta;  // this group's net
    int64_t n;           // rows in the group
    int64_t row0;        // first row of the group in the packed [N, .] workspace matrices
    int64_t pitch;       // floats between consecutive source rows (64 for tables, 128 for exchanged [last|hat] pairs)
};

// conv prologue: fc1 input for every row of the groups, as plain A[N,320] (or null) and/or as a packed
// tensor-core operand Apk (umma_pack.cuh, 128-row tiles, or null); rowsq[n] = sum x_hat^2 (or null)
// zero_y (optional): Y[N,64] rows of the groups are cleared (split-K fc2 accumulates into them)
// zero_dA (optional): dA[N,320] rows of the groups are cleared (split-K d1 accumulates into them)
int sml_launch_conv_fwd(const SmlRowGroup *groups, int n_groups, int variant, float *A, uint8_t *Apk, float *rowsq,
                        cudaStream_t st, float *zero_y = nullptr, float *zero_dA = nullptr);

// conv backward.  dA [N,320].  mode 0: scatter (dx_hat + l2*x_hat) into g_tab rows (atomic);
// mode 1: write dx_hat to d_rows [N,64];  theta_grad != null: accumulate conv1/conv2 grads.
struct SmlConvBwdGroup {
    SmlRowGroup g;
    float *g_tab;      // dense gradient table for scatter (mode 0) or null
    float *g_theta;    // this group's net gradient block or null
    int64_t d_base;    // mode 1: >= 0 -> the row gathered through id k writes d_rows[d_base + k]; < 0 -> d_rows[row0 + r]
};
// adaptive (scatter mode, first group = the user rows): adds 2 * adaptive * x_hat / ||x_hat|| per occurrence (--need_adaptive)
int sml_launch_conv_bwd(const SmlConvBwdGroup *groups, int n_groups, int variant, const float *dA, float l2,
                        float *d_rows, cudaStream_t st, float adaptive = 0.f);

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
                         cudaStream_t st, int ksplit = 1);
// 1 = tensor-core GEMMs (default), 0 = SIMT fp32 GEMMs (SML_GEMM=simt in the environment, for A/B comparisons)
int sml_use_tensor_cores();

// column sums: out[c] (+)= sum_r X[r, c]   (bias gradients)
struct SmlColsumProb { const float *X; float *out; int rows, cols, ld; };
int sml_launch_colsum(const SmlColsumProb *probs, int n_probs, cudaStream_t st);

// loss + dY.  Rows of Y / dY / rowsq: user b at b, positive item at row_pos + b, negative at row_neg + b.
// dYpk (optional): dY additionally as a packed tensor-core operand (K = 64).
// gb_user / gb_item (optional): += column sums of dY over the user rows / the item rows (fc2 bias gradients)
int sml_launch_loss(const float *Y, const float *rowsq, int64_t B, int64_t row_pos, int64_t row_neg, int loss_kind,
                    int normalize_user, float l2, float *dY, uint8_t *dYpk, float *scores, float *loss_out, float *partials,
                    unsigned int *ticket, cudaStream_t st, float *gb_user = nullptr, float *gb_item = nullptr,
                    float *zero_dA = nullptr,    // zero_dA (optional): dA[N,320] rows of the batch are cleared (split-K d1)
                    float adaptive = 0.f);       // --need_adaptive: + adaptive * sum_b ||x_hat user row b|| (needs rowsq)

// tcgen05 GEMM with pre-packed operands (umma_packed.cu)
enum { SML_PK_FC1 = 0, SML_PK_FC2 = 1, SML_PK_D2 = 2, SML_PK_D1 = 3, SML_PK_FC1_FC2 = 4 };
struct SmlPkProb {
    const uint8_t *A;     // packed A, 128-row tiles
    const uint8_t *B;     // packed B (weights), BN-row tiles
    int KC;               // K / 32
    int m_tiles;          // row tiles of this problem
    int a_tile0;          // first tile of this problem inside A
    int64_t row0;         // plain row of the first tile
    int M;                // valid rows (plain stores are guarded)
    int N;                // output columns
    const float *bias;    // [N] (fc1 / fc2)
    const float *aux;     // plain [rows][ldc] pre-activations for the GELU' epilogue (d2)
    float *C;             // plain output [rows][ldc] or null
    int ldc;
    uint8_t *Cpk;         // packed output = A operand of the next GEMM (K = N), or null
    int c_tile0;
    float *colsum;        // d2 only: += column sums of the valid output rows (fc1 bias gradient), or null
    // SML_PK_FC1_FC2 only: fc2 fused into the fc1 tiles
    const uint8_t *B2;    // packed W2 (P2 operand, 64-row blocks)
    const float *bias2;   // [64]
    float *Y;             // plain fc2 output [rows][ldy], zeroed by the caller (the tiles add their partial products)
    int ldy;
};
// SML_PK_D2 with the loss fused into its prologue (no k_loss launch in the step chain): every d2 CTA computes the dY tile it
// multiplies straight from Y (row dots, BCE / BPR derivative: the arithmetic of k_loss, conv.cu) and writes it into shared
// memory as its packed A operand; the CTAs of output-column tile 0 also emit what k_loss emitted (plain dY, scores, the scalar
// loss through partials + ticket in a fixed order, the fc2 bias gradients).  Rows as in sml_launch_loss.
struct SmlPkLoss {
    const float *Y;        // [rows][64] fc2 output
    const float *rowsq;    // [rows] sum x_hat^2 (MF step l2 term) or null
    int64_t B, row_pos, row_neg;
    int loss_kind, normalize_user;
    float l2, adaptive;
    float *dY;             // plain dL/dY [rows][64] (transfer step: operand of the fc2 weight gradient) or null
    float *scores;         // [2B] or null
    float *loss_out;       // [2]
    float *partials;       // [4 * user tiles]
    unsigned int *ticket;
    float *gb_user, *gb_item;   // += column sums of dY (fc2 bias gradients) or null
};
int sml_launch_umma_packed(const SmlPkProb *probs, int n_probs, int epi, cudaStream_t st, int ksplit = 1, const SmlPkLoss *loss = nullptr);
// packed weight operands of the nets: per net SML_PK_THETA_BYTES
constexpr size_t SML_PK_OFF_P1 = 0;            // W1   as B[512][320], 128-row blocks (fc1 forward)
constexpr size_t SML_PK_OFF_P2 = 1474560;      // W2   as B[64][512],   64-row blocks (fc2 forward)
constexpr size_t SML_PK_OFF_P3 = 1769472;      // W2^T as B[512][64],  128-row blocks (dZ1 = dY W2)
constexpr size_t SML_PK_OFF_P4 = 2064384;      // W1^T as B[320][512],  64-row blocks (dA = dZ1 W1)
constexpr size_t SML_PK_THETA_BYTES = 3538944;
// adam_state != null: also performs the step's Adam tick (betas 0.9 / 0.999) inside the same launch
int sml_launch_pack_theta(const float *theta, uint8_t *out, int n_nets, cudaStream_t st, int64_t *adam_state = nullptr, double lr = 0.0);

int sml_launch_row_normalize(float *Y, int64_t n, cudaStream_t st);
int sml_launch_sumsq(const float *g, int64_t n, float *sumsq, float *partials, unsigned int *ticket, cudaStream_t st);

// fused transfer forward (umma_fused_fwd.cu): one persistent kernel for conv -> fc1 -> GELU -> fc2 over a stream of rows.
// wpk = the net's W1 / W2 packed by sml_launch_pack_fused (sml_fused_fwd_packed_bytes() bytes).
size_t sml_fused_fwd_packed_bytes();
int sml_launch_pack_fused(const float *theta_net, uint8_t *out, cudaStream_t st);
int sml_launch_transfer_fused(const float *x_t, const float *x_hat, const int64_t *ids, int64_t n_rows, int64_t pitch, int variant,
                              const float *theta_net, const uint8_t *wpk, int normalize_out, float *out, cudaStream_t st);
// 1 = sml_transfer_fwd uses the fused kernel (default); SML_FUSED_FWD=0 or sml_debug_set_mask(2048) select the three-kernel path
int sml_use_fused_fwd();
// 1 = the step forward fuses fc2 into the fc1 tiles (default); SML_FUSE_FC2=0 or sml_debug_set_mask(4096) keep two launches
int sml_use_fused_fc2();
// 1 = the step computes loss + dL/dY inside the d2 GEMM (default); SML_FUSE_LOSS=0 or sml_debug_set_mask(8192) keep k_loss
int sml_use_fused_loss();
