// Row-local parts of the transfer network and of run_MF:
//   * conv prologue  : gather x_t / x_hat rows, build the stacked channels and run
//                      conv1 (R x 1, 1->10) -> GELU -> conv2 (1x1, 10->5) -> GELU, producing the
//                      [N, 320] fc1 input (channel-major flatten, model/conv_transfer.py:37-44,92-103)
//   * conv backward  : recomputes the conv internals from the two 256 B source rows (cheaper than
//                      saving 15 x 64 pre-activations per row), back-propagates dA to the x_hat
//                      channel only (x_com is detached, conv_transfer.py:93) and either scatters
//                      dx_hat + l2*x_hat into the dense gradient table (MF step,
//                      model/transfer.py:486-502) or reduces the conv1/conv2 parameter gradients
//                      (transfer step)
//   * loss           : row dots, BCE-mean / BPR-sum and dL/dY (conv_transfer.py:120-134), with a
//                      fixed-order reduction of the scalar loss
// One warp owns one row; lane l owns latent dims l and l+32, so every global access is a
// coalesced 128 B line and the row norm is one shuffle tree.
#include "sml_common.cuh"
#include "umma_pack.cuh"

namespace {

constexpr int CONV_THREADS = 128;
constexpr int CONV_WARPS = CONV_THREADS / 32;
constexpr int MAX_GROUPS = 3;

struct ConvParams {
    SmlRowGroup g[MAX_GROUPS];
    int n_groups;
};

// conv parameters of one net staged in shared memory
struct ConvW {
    float w1[10][3];
    float b1[10];
    float w2[5][10];
    float b2[5];
};

template <int R>
__device__ __forceinline__ void load_convw(ConvW &w, const float *__restrict__ theta) {
    for (int i = threadIdx.x; i < 10 * R; i += blockDim.x) w.w1[i / R][i % R] = theta[SML_OFF_C1W + i];
    for (int i = threadIdx.x; i < 10; i += blockDim.x) w.b1[i] = theta[SML_OFF_C1B + i];
    for (int i = threadIdx.x; i < 50; i += blockDim.x) w.w2[i / 10][i % 10] = theta[SML_OFF_C2W + i];
    for (int i = threadIdx.x; i < 5; i += blockDim.x) w.b2[i] = theta[SML_OFF_C2B + i];
}

template <int R>
__device__ __forceinline__ void conv_point(const ConvW &w, float x0, float x1, float x2, float (&z1)[10], float (&h1)[10],
                                           float (&z2)[5]) {
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        float z = w.b1[c];
        z = fmaf(w.w1[c][0], x0, z);
        z = fmaf(w.w1[c][1], x1, z);
        if (R == 3) z = fmaf(w.w1[c][2], x2, z);
        z1[c] = z;
        h1[c] = sml_gelu(z);
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        float z = w.b2[m];
#pragma unroll
        for (int c = 0; c < 10; ++c) z = fmaf(w.w2[m][c], h1[c], z);
        z2[m] = z;
    }
}

// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(CONV_THREADS)
k_conv_fwd(ConvParams P, float *__restrict__ A, uint8_t *__restrict__ Apk, float *__restrict__ rowsq, float *__restrict__ zero_y,
           float *__restrict__ zero_dA) {
    __shared__ ConvW sw;
    sml_pdl_wait();
    sml_pdl_trigger();
    const int gi = blockIdx.y;
    const SmlRowGroup g = P.g[gi];
    load_convw<R>(sw, g.theta);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * CONV_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * CONV_WARPS;
    // two warps per row: warp 2r+h converts latent dims [32h, 32h+32) (both load the whole row for the norm)
    for (int64_t wr = warp0; wr < 2 * g.n; wr += nwarps) {
        const int64_t r = wr >> 1;
        const int h0 = (int)(wr & 1);
        const int64_t id = g.ids ? __ldg(g.ids + r) : r;
        const float *xt = g.x_t + id * g.pitch;
        const float *xh = g.x_hat + id * g.pitch;
        float x0[2] = {__ldg(xt + lane), __ldg(xt + lane + 32)};
        float x1[2] = {__ldg(xh + lane), __ldg(xh + lane + 32)};
        float x2[2] = {0.f, 0.f};
        if (R == 3) {
            const float nrm = sqrtf(warp_sum(x0[0] * x0[0] + x0[1] * x0[1]));   // conv_transfer.py:94
            x2[0] = (x0[0] * x1[0]) / nrm;                                       // :93,98 (no eps: NaN on zero rows)
            x2[1] = (x0[1] * x1[1]) / nrm;
        }
        if (rowsq) {
            const float q = warp_sum(x1[0] * x1[0] + x1[1] * x1[1]);
            if (lane == 0 && h0 == 0) rowsq[g.row0 + r] = q;
        }
        if (zero_y) zero_y[(g.row0 + r) * SML_D + lane + 32 * h0] = 0.f;   // split-K fc2 accumulates into Y (no memset node in the chain)
        if (zero_dA) {                                                       // split-K d1 accumulates into dA
#pragma unroll
            for (int k = 0; k < SML_FC1_IN / 2; k += 32) zero_dA[(g.row0 + r) * SML_FC1_IN + h0 * (SML_FC1_IN / 2) + k + lane] = 0.f;
        }
        float *a = A ? A + (g.row0 + r) * SML_FC1_IN : nullptr;
#pragma unroll
        {
            const int h = h0;
            const float x0h = h0 ? x0[1] : x0[0], x1h = h0 ? x1[1] : x1[0], x2h = h0 ? x2[1] : x2[0];   // no dynamic register indexing
            float z1[10], h1[10], z2[5];
            conv_point<R>(sw, x0h, x1h, x2h, z1, h1, z2);
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const float v = sml_gelu(z2[m]);
                const int k = m * SML_D + lane + 32 * h;          // channel-major flatten (conv_transfer.py:43)
                if (a) a[k] = v;
                if (Apk) pk_store1(Apk, 128, SML_FC1_IN / PK_BK, g.row0 + r, k, v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct ConvBwdParams {
    SmlConvBwdGroup g[MAX_GROUPS];
    int n_groups;
};

// MODE 0: scatter into dense gradient table; MODE 1: write d_rows; THETA: accumulate conv grads
template <int R, int MODE, bool THETA>
__global__ void __launch_bounds__(CONV_THREADS)
k_conv_bwd(ConvBwdParams P, const float *__restrict__ dA, float l2, float *__restrict__ d_rows, float adaptive) {
    __shared__ ConvW sw;
    sml_pdl_wait();
    sml_pdl_trigger();
    __shared__ float s_acc[96];
    const int gi = blockIdx.y;
    const SmlConvBwdGroup bg = P.g[gi];
    const SmlRowGroup g = bg.g;
    load_convw<R>(sw, g.theta);
    if (THETA) for (int i = threadIdx.x; i < 96; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * CONV_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * CONV_WARPS;
    // per-thread partial parameter gradients, flat in the order of s_acc:
    //   [0,50) w2[m][c]   [50,55) b2[m]   [55, 55+10R) w1[c][r]   [85,95) b1[c]   (95: padding)
    float acc[96];
    if (THETA) {
#pragma unroll
        for (int i = 0; i < 96; ++i) acc[i] = 0.f;
    }
    // two warps per row: warp 2r+h converts latent dims [32h, 32h+32) (both load the whole row for the norm)
    for (int64_t wr = warp0; wr < 2 * g.n; wr += nwarps) {
        const int64_t r = wr >> 1;
        const int h0 = (int)(wr & 1);
        const int64_t id = g.ids ? __ldg(g.ids + r) : r;
        const float *xt = g.x_t + id * g.pitch;
        const float *xh = g.x_hat + id * g.pitch;
        float x0[2] = {__ldg(xt + lane), __ldg(xt + lane + 32)};
        float x1[2] = {__ldg(xh + lane), __ldg(xh + lane + 32)};
        float x2[2] = {0.f, 0.f};
        if (R == 3) {
            const float nrm = sqrtf(warp_sum(x0[0] * x0[0] + x0[1] * x0[1]));
            x2[0] = (x0[0] * x1[0]) / nrm;
            x2[1] = (x0[1] * x1[1]) / nrm;
        }
        const float *da = dA + (g.row0 + r) * SML_FC1_IN;
#pragma unroll
        {
            const int h = h0;
            const float x0h = h0 ? x0[1] : x0[0], x1h = h0 ? x1[1] : x1[0], x2h = h0 ? x2[1] : x2[0];   // no dynamic register indexing
            float z1[10], h1[10], z2[5];
            conv_point<R>(sw, x0h, x1h, x2h, z1, h1, z2);
            float dz2[5];
#pragma unroll
            for (int m = 0; m < 5; ++m) dz2[m] = __ldg(da + m * SML_D + lane + 32 * h) * sml_gelu_grad(z2[m]);
            float dx1 = 0.f;
#pragma unroll
            for (int c = 0; c < 10; ++c) {
                float dh = 0.f;
#pragma unroll
                for (int m = 0; m < 5; ++m) dh = fmaf(sw.w2[m][c], dz2[m], dh);
                const float dz1 = dh * sml_gelu_grad(z1[c]);
                dx1 = fmaf(sw.w1[c][1], dz1, dx1);
                if (THETA) {
                    acc[85 + c] += dz1;
                    acc[55 + c * R + 0] = fmaf(dz1, x0h, acc[55 + c * R + 0]);
                    acc[55 + c * R + 1] = fmaf(dz1, x1h, acc[55 + c * R + 1]);
                    if (R == 3) acc[55 + c * R + 2] = fmaf(dz1, x2h, acc[55 + c * R + 2]);
                }
            }
            if (THETA) {
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    acc[50 + m] += dz2[m];
#pragma unroll
                    for (int c = 0; c < 10; ++c) acc[m * 10 + c] = fmaf(dz2[m], h1[c], acc[m * 10 + c]);
                }
            }
            if (MODE == 0) {
                // dense-gradient scatter: grad of l2*0.5*sum(w^2) is l2*w per occurrence (transfer.py:486)
                float gsc = fmaf(l2, x1h, dx1);
                if (adaptive != 0.f && gi == 0) {
                    // --need_adaptive (transfer.py:490-499): beta * count_u / ||w_u||.detach() * ||w_u||^2 per distinct user
                    // = 2 * beta * w_u / ||w_u|| per occurrence
                    const float nh = sqrtf(warp_sum(x1[0] * x1[0] + x1[1] * x1[1]));
                    gsc = fmaf(2.0f * adaptive / nh, x1h, gsc);
                }
                atomicAdd(bg.g_tab + id * SML_D + lane + 32 * h, gsc);
            } else if (d_rows) {
                d_rows[(bg.d_base >= 0 ? bg.d_base + id : g.row0 + r) * SML_D + lane + 32 * h] = dx1;
            }
        }
    }
    if (THETA) {
        // transposed warp reduction: at every halving step a lane keeps one half of its values and hands the other half to
        // its partner (96 -> 48 -> 24 -> 12 -> 6 -> 3 values, 93 shuffles instead of 95 five-step butterflies); lane L ends
        // with the full 32-lane sums of 3 parameters, then shared -> one global atomic per parameter per CTA
#define SML_HALVE(N_, OFF_)                                                        \
        {                                                                          \
            const bool up = (lane & OFF_) != 0;                                    \
            _Pragma("unroll") for (int i = 0; i < N_; ++i) {                       \
                const float send = up ? acc[i] : acc[i + N_];                      \
                const float keep = up ? acc[i + N_] : acc[i];                      \
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF_);          \
            }                                                                      \
        }
        SML_HALVE(48, 16) SML_HALVE(24, 8) SML_HALVE(12, 4) SML_HALVE(6, 2) SML_HALVE(3, 1)
#undef SML_HALVE
        const int base = ((lane & 16) ? 48 : 0) + ((lane & 8) ? 24 : 0) + ((lane & 4) ? 12 : 0) + ((lane & 2) ? 6 : 0) + ((lane & 1) ? 3 : 0);
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (base + j < 95) atomicAdd(&s_acc[base + j], acc[j]);
        __syncthreads();
        float *gt = bg.g_theta;
        for (int i = threadIdx.x; i < 95; i += blockDim.x) {
            const float v = s_acc[i];
            if (i < 50) atomicAdd(gt + SML_OFF_C2W + i, v);
            else if (i < 55) atomicAdd(gt + SML_OFF_C2B + (i - 50), v);
            else if (i < 55 + 10 * R) atomicAdd(gt + SML_OFF_C1W + (i - 55), v);
            else if (i >= 85) atomicAdd(gt + SML_OFF_C1B + (i - 85), v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// loss + dL/dY.  One warp per triple; CTA-level and grid-level sums are done in a fixed order
// (partials[] + last-CTA ticket) so the scalar loss is run-to-run deterministic.
constexpr int LOSS_THREADS = 256;
constexpr int LOSS_WARPS = LOSS_THREADS / 32;

__global__ void __launch_bounds__(LOSS_THREADS)
k_loss(const float *__restrict__ Y, const float *__restrict__ rowsq, int64_t B, int64_t rowP, int64_t rowN, int loss_kind,
       int normalize_user, float l2, float *__restrict__ dY, uint8_t *__restrict__ dYpk, float *__restrict__ scores,
       float *__restrict__ loss_out, float *__restrict__ partials, unsigned int *__restrict__ ticket,
       float *__restrict__ gb_user, float *__restrict__ gb_item, float *__restrict__ zero_dA, float adaptive) {
    __shared__ float s_part[LOSS_WARPS][4];
    __shared__ float s_gb[2][SML_D];
    sml_pdl_wait();
    sml_pdl_trigger();
    float gbu[2] = {0.f, 0.f}, gbi[2] = {0.f, 0.f};          // fc2 bias gradients: column sums of dY (this thread's 2 columns)
    if (gb_user && threadIdx.x < 2 * SML_D) s_gb[threadIdx.x / SML_D][threadIdx.x % SML_D] = 0.f;
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float acc_pos = 0.f, acc_neg = 0.f, acc_sq = 0.f, acc_ad = 0.f;   // lane 0 only
    const float invB = 1.0f / (float)B;
    for (int64_t b = (int64_t)blockIdx.x * LOSS_WARPS + w; b < B; b += (int64_t)gridDim.x * LOSS_WARPS) {
        const float *yu = Y + b * SML_D, *yi = Y + (rowP + b) * SML_D, *yj = Y + (rowN + b) * SML_D;
        if (zero_dA) {                                       // split-K d1 accumulates into dA (no memset node in the chain)
#pragma unroll
            for (int k = 0; k < SML_FC1_IN; k += 32) {
                zero_dA[b * SML_FC1_IN + k + lane] = 0.f;
                zero_dA[(rowP + b) * SML_FC1_IN + k + lane] = 0.f;
                zero_dA[(rowN + b) * SML_FC1_IN + k + lane] = 0.f;
            }
        }
        float u[2] = {yu[lane], yu[lane + 32]};
        const float i_[2] = {yi[lane], yi[lane + 32]};
        const float j_[2] = {yj[lane], yj[lane + 32]};
        float inv_n = 1.0f;
        if (normalize_user) {   // ConvTransfer 'user': x / ||x||.detach()  (conv_transfer.py:62-63)
            const float nrm = sqrtf(warp_sum(u[0] * u[0] + u[1] * u[1]));
            u[0] = u[0] / nrm; u[1] = u[1] / nrm;
            inv_n = 1.0f / nrm;
        }
        const float sp = warp_sum(fmaf(u[1], i_[1], u[0] * i_[0]));
        const float sn = warp_sum(fmaf(u[1], j_[1], u[0] * j_[0]));
        float dsp, dsn;
        if (loss_kind == SML_LOSS_BCE) {
            const float gp = sml_sigmoid(sp), gn = sml_sigmoid(sn);
            const float ap = gp + 1e-15f, an = (1.0f - gn) + 1e-15f;     // conv_transfer.py:124-125
            if (lane == 0) { acc_pos += logf(ap); acc_neg += logf(an); }
            dsp = -(gp * (1.0f - gp)) / ap * invB;
            dsn = (gn * (1.0f - gn)) / an * invB;
        } else {
            const float x = sp - sn;                                      // :128
            // -logsigmoid(x) = softplus(-x)
            if (lane == 0) acc_pos += fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
            dsp = -sml_sigmoid(-x);
            dsn = -dsp;
        }
        if (rowsq && lane == 0) {
            acc_sq += rowsq[b] + rowsq[rowP + b] + rowsq[rowN + b];
            if (adaptive != 0.f) acc_ad += sqrtf(rowsq[b]);           // sum over occurrences of ||w_u|| (transfer.py:490-499)
        }
        if (scores && lane == 0) { scores[b] = sp; scores[B + b] = sn; }
        if (dY) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = lane + 32 * h;
                const float du = (dsp * i_[h] + dsn * j_[h]) * inv_n, di = dsp * u[h], dj = dsn * u[h];
                gbu[h] += du; gbi[h] += di + dj;
                dY[b * SML_D + k] = du;
                dY[(rowP + b) * SML_D + k] = di;
                dY[(rowN + b) * SML_D + k] = dj;
                if (dYpk) {
                    pk_store1(dYpk, 128, SML_D / PK_BK, b, k, du);
                    pk_store1(dYpk, 128, SML_D / PK_BK, rowP + b, k, di);
                    pk_store1(dYpk, 128, SML_D / PK_BK, rowN + b, k, dj);
                }
            }
        }
    }
    if (lane == 0) { s_part[w][0] = acc_pos; s_part[w][1] = acc_neg; s_part[w][2] = acc_sq; s_part[w][3] = acc_ad; }
    __syncthreads();
    if (gb_user) {
#pragma unroll
        for (int h = 0; h < 2; ++h) { atomicAdd(&s_gb[0][lane + 32 * h], gbu[h]); atomicAdd(&s_gb[1][lane + 32 * h], gbi[h]); }
        __syncthreads();
        if (threadIdx.x < SML_D) atomicAdd(gb_user + threadIdx.x, s_gb[0][threadIdx.x]);
        else if (threadIdx.x < 2 * SML_D) atomicAdd(gb_item + threadIdx.x - SML_D, s_gb[1][threadIdx.x - SML_D]);
    }
    if (threadIdx.x == 0) {
        float p = 0.f, n = 0.f, q = 0.f, ad = 0.f;
        for (int i = 0; i < LOSS_WARPS; ++i) { p += s_part[i][0]; n += s_part[i][1]; q += s_part[i][2]; ad += s_part[i][3]; }
        partials[4 * blockIdx.x + 0] = p; partials[4 * blockIdx.x + 1] = n; partials[4 * blockIdx.x + 2] = q; partials[4 * blockIdx.x + 3] = ad;
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        float p = 0.f, n = 0.f, q = 0.f, ad = 0.f;
        for (unsigned i = 0; i < gridDim.x; ++i) {
            p += __ldcg(partials + 4 * i); n += __ldcg(partials + 4 * i + 1); q += __ldcg(partials + 4 * i + 2); ad += __ldcg(partials + 4 * i + 3);
        }
        float loss;
        if (loss_kind == SML_LOSS_BCE) loss = (-(p * invB)) + (-(n * invB));   // -mean - mean
        else loss = p;                                                           // -sum(logsigmoid)
        loss = loss + l2 * (0.5f * q);                                           // transfer.py:486-488
        if (adaptive != 0.f) loss = loss + adaptive * ad;                        // :490-499
        loss_out[0] = loss;
        loss_out[1] += loss;
        *ticket = 0;   // re-arm for the next launch (graph replay)
    }
}

__global__ void __launch_bounds__(256) k_row_normalize(float *__restrict__ Y, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    float *y = Y + r * SML_D;
    const float a = y[lane], b = y[lane + 32];
    const float nrm = sqrtf(warp_sum(a * a + b * b));
    y[lane] = a / nrm;
    y[lane + 32] = b / nrm;
}

int grid_for_rows(int64_t max_n) {
    int64_t blocks = (2 * max_n + CONV_WARPS - 1) / CONV_WARPS;      // two warps per row
    const int64_t cap = (int64_t)sml_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

int sml_launch_conv_fwd(const SmlRowGroup *groups, int n_groups, int variant, float *A, uint8_t *Apk, float *rowsq,
                        cudaStream_t st, float *zero_y, float *zero_dA) {
    SML_REQUIRE(n_groups >= 1 && n_groups <= MAX_GROUPS, SML_E_BADARG, "conv_fwd: bad group count %d", n_groups);
    ConvParams P;
    P.n_groups = n_groups;
    int64_t max_n = 0;
    for (int i = 0; i < n_groups; ++i) { P.g[i] = groups[i]; if (groups[i].n > max_n) max_n = groups[i].n; }
    if (max_n == 0) return SML_OK;
    dim3 grid(grid_for_rows(max_n), n_groups);
    if (variant == SML_VARIANT_COM) SML_CUDA_OK(sml_launch(k_conv_fwd<3>, grid, dim3(CONV_THREADS), 0, st, P, A, Apk, rowsq, zero_y, zero_dA));
    else SML_CUDA_OK(sml_launch(k_conv_fwd<2>, grid, dim3(CONV_THREADS), 0, st, P, A, Apk, rowsq, zero_y, zero_dA));
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_launch_conv_bwd(const SmlConvBwdGroup *groups, int n_groups, int variant, const float *dA, float l2,
                        float *d_rows, cudaStream_t st, float adaptive) {
    SML_REQUIRE(n_groups >= 1 && n_groups <= MAX_GROUPS, SML_E_BADARG, "conv_bwd: bad group count %d", n_groups);
    ConvBwdParams P;
    P.n_groups = n_groups;
    int64_t max_n = 0;
    bool scatter = groups[0].g_tab != nullptr, theta = groups[0].g_theta != nullptr;
    for (int i = 0; i < n_groups; ++i) {
        P.g[i] = groups[i];
        if (groups[i].g.n > max_n) max_n = groups[i].g.n;
        SML_REQUIRE((groups[i].g_tab != nullptr) == scatter && (groups[i].g_theta != nullptr) == theta, SML_E_BADARG,
                    "conv_bwd: groups must agree on scatter/theta-grad mode");
    }
    if (max_n == 0) return SML_OK;
    // fewer, fatter CTAs when parameter gradients are reduced (one atomic flush per CTA)
    int gx = grid_for_rows(max_n);
    // each CTA ends with 95 global atomics onto the same 95 parameters: a few dozen CTAs per group keep that cheap
    // (96 per group measured 2-6 us slower per transfer step than 32, profiles/r01_tr_step_breakdown.md)
    // (at the 8 192-row groups of the sharded step 32 CTAs serialise 128 row halves per warp: 209 us; scale the cap with the rows)
    if (theta) { int cap = (int)(max_n / 32); if (cap < 32) cap = 32; if (cap > 2 * sml_sm_count()) cap = 2 * sml_sm_count(); if (gx > cap) gx = cap; }
    dim3 grid(gx, n_groups);
#define SML_CB(R_, MODE_, TH_) SML_CUDA_OK(sml_launch(k_conv_bwd<R_, MODE_, TH_>, grid, dim3(CONV_THREADS), 0, st, P, dA, l2, d_rows, adaptive))
    const bool com = variant == SML_VARIANT_COM;
    if (scatter) {
        if (theta) { if (com) SML_CB(3, 0, true); else SML_CB(2, 0, true); }
        else { if (com) SML_CB(3, 0, false); else SML_CB(2, 0, false); }
    } else {
        if (theta) { if (com) SML_CB(3, 1, true); else SML_CB(2, 1, true); }
        else { if (com) SML_CB(3, 1, false); else SML_CB(2, 1, false); }
    }
#undef SML_CB
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_launch_loss(const float *Y, const float *rowsq, int64_t B, int64_t row_pos, int64_t row_neg, int loss_kind,
                    int normalize_user, float l2, float *dY, uint8_t *dYpk, float *scores, float *loss_out, float *partials,
                    unsigned int *ticket, cudaStream_t st, float *gb_user, float *gb_item, float *zero_dA, float adaptive) {
    int64_t blocks = (B + LOSS_WARPS - 1) / LOSS_WARPS;
    if (blocks > 1024) blocks = 1024;   // partials[] holds 4 * 1024 floats
    SML_CUDA_OK(sml_launch(k_loss, dim3((unsigned)blocks), dim3(LOSS_THREADS), 0, st, Y, rowsq, B, row_pos, row_neg, loss_kind, normalize_user,
                           l2, dY, dYpk, scores, loss_out, partials, ticket, gb_user, gb_item, zero_dA, adaptive));
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_launch_row_normalize(float *Y, int64_t n, cudaStream_t st) {
    if (n <= 0) return SML_OK;
    const int64_t threads = n * 32;
    k_row_normalize<<<(int)((threads + 255) / 256), 256, 0, st>>>(Y, n);
    SML_LAUNCH_OK();
    return SML_OK;
}
