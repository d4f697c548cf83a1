// torch.optim.Adam semantics (model/transfer.py:392-393 of the reference: lr, betas (0.9, 0.999),
// eps 1e-8, amsgrad off, coupled L2 weight decay) as one streaming kernel.
//
// The MF optimizer in the reference is DENSE Adam over whole nn.Embedding tables
// (nn.Embedding(sparse=False), model/MF.py:21-24): rows touched at an earlier step keep moving
// through the momentum tail, so the usual lazy/sparse Adam is not equivalent.  Two exact forms:
//  * k_adam_dense: the sweep.  HBM-bound: per element read p, m, v, g and write p, m, v (+ the zeroed g
//    that replaces zero_grad()): 32 B per float, all 128-bit accesses, grid sized to the SM count.
//  * k_adam_rows / k_adam_flush: the same arithmetic, row-lazy.  A row that had no gradient for k steps
//    is brought up to date by replaying those k zero-gradient steps in registers (the per-step scalars come
//    from the history ring the tick kernel keeps), so only the 3B rows of a batch move per step and the
//    whole table moves once per flush: same bits as the sweep, 1/steps of its HBM traffic.
#include <math.h>

#include "sml_common.cuh"

namespace {

__global__ void k_adam_tick(int64_t *state, double lr, double beta1, double beta2) {
    sml_pdl_wait();
    sml_pdl_trigger();
    if (threadIdx.x == 0 && blockIdx.x == 0) sml_adam_tick_body(state, lr, beta1, beta2);
}

// sum of squares with a fixed summation order (per-CTA partials, the last CTA adds them up): ||g||^2 for clip_grad_norm_
__global__ void __launch_bounds__(256)
k_sumsq(const float4 *__restrict__ g, int64_t n4, float *__restrict__ out, float *__restrict__ partials, unsigned int *__restrict__ ticket) {
    sml_pdl_wait();
    sml_pdl_trigger();
    __shared__ float s_w[8];
    __shared__ bool s_last;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = g[i];
        acc = fmaf(x.x, x.x, acc); acc = fmaf(x.y, x.y, acc); acc = fmaf(x.z, x.z, acc); acc = fmaf(x.w, x.w, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_w[i];
        partials[blockIdx.x] = t;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        float t = 0.f;
        for (unsigned i = 0; i < gridDim.x; ++i) t += __ldcg(partials + i);
        *out = t;
        *ticket = 0;
    }
}

template <bool ZERO>
__global__ void __launch_bounds__(256)
k_adam_dense(float4 *__restrict__ p, float4 *__restrict__ m, float4 *__restrict__ v, float4 *__restrict__ g, int64_t n4,
             const int64_t *__restrict__ state, float b1c, float beta2, float b2c, float eps, float wd,
             const float *__restrict__ sumsq, float max_norm) {
    sml_pdl_wait();
    sml_pdl_trigger();
    // clip_grad_norm_: every gradient element times min(1, max_norm / (||g|| + 1e-6))
    float gs = 1.0f;
    if (sumsq) { const float c = max_norm / (sqrtf(__ldcg(sumsq)) + 1e-6f); gs = c < 1.0f ? c : 1.0f; }
    // b1c = (float)(1 - beta1), b2c = (float)(1 - beta2) are rounded from the double differences on the
    // host, as torch does (1.0f - 0.999f would be off by 5e-5 relative)
    const float *f = reinterpret_cast<const float *>(state + 1);
    const float step_size = f[0], bc2_sqrt = f[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = p[i], mm = m[i], vv = v[i];
        float4 gg = g[i];
        if (sumsq) { gg.x *= gs; gg.y *= gs; gg.z *= gs; gg.w *= gs; }
        sml_adam1(pp.x, mm.x, vv.x, gg.x, b1c, beta2, b2c, step_size, bc2_sqrt, eps, wd);
        sml_adam1(pp.y, mm.y, vv.y, gg.y, b1c, beta2, b2c, step_size, bc2_sqrt, eps, wd);
        sml_adam1(pp.z, mm.z, vv.z, gg.z, b1c, beta2, b2c, step_size, bc2_sqrt, eps, wd);
        sml_adam1(pp.w, mm.w, vv.w, gg.w, b1c, beta2, b2c, step_size, bc2_sqrt, eps, wd);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (ZERO) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---- row-lazy exact Adam ---------------------------------------------------------------------
struct AdamRowsGroup {
    float *p, *m, *v, *g;
    int32_t *stamp;
    const int64_t *ids;
    int64_t n;
};
struct AdamRowsParams { AdamRowsGroup g[3]; };

__device__ __forceinline__ float2 adam_hist(const int64_t *state, int s) {
    return *reinterpret_cast<const float2 *>(state + 4 + (s & (SML_ADAM_HISTORY - 1)));
}

// One warp per listed id (2 floats per lane).  The claim on stamp[id] makes exactly one warp the owner of a row that
// appears several times in the batch.
template <bool APPLY>
__global__ void __launch_bounds__(256)
k_adam_rows(AdamRowsParams P, const int64_t *__restrict__ state, float b1c, float beta2, float b2c, float eps) {
    sml_pdl_wait();
    sml_pdl_trigger();
    const int t = (int)state[0];
    const int target = APPLY ? t : t - 1;
    const int lane = threadIdx.x & 31;
    const int64_t n0 = P.g[0].n, n1 = n0 + P.g[1].n, total = n1 + P.g[2].n;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
        const int gi = w < n0 ? 0 : (w < n1 ? 1 : 2);
        const AdamRowsGroup &G = P.g[gi];
        const int64_t id = G.ids[w - (gi == 0 ? 0 : (gi == 1 ? n0 : n1))];
        int old = 0;
        if (lane == 0) old = atomicExch(G.stamp + id, target);
        old = __shfl_sync(0xffffffffu, old, 0);
        // SML_STAMP_IDLE: exp_avg = exp_avg_sq = +0 since the table was created -- nothing to replay, and this warp is the first
        // to touch the row at this step
        if (old != SML_STAMP_IDLE && old >= target) continue;
        if (old == SML_STAMP_IDLE && !APPLY) continue;
        const size_t e = (size_t)id * SML_D + 2 * lane;
        float2 pp = *reinterpret_cast<float2 *>(G.p + e), mm = *reinterpret_cast<float2 *>(G.m + e),
               vv = *reinterpret_cast<float2 *>(G.v + e);
        // exp_avg = exp_avg_sq = +0 (a row no gradient ever reached): a zero-gradient step changes nothing, bit for bit
        const bool idle = old == SML_STAMP_IDLE || (__float_as_uint(mm.x) | __float_as_uint(mm.y) | __float_as_uint(vv.x) | __float_as_uint(vv.y)) == 0u;
        for (int s = idle ? t : old + 1; s < t; ++s) {   // the zero-gradient steps this row missed
            const float2 c = adam_hist(state, s);
            sml_adam1(pp.x, mm.x, vv.x, 0.f, b1c, beta2, b2c, c.x, c.y, eps, 0.f);
            sml_adam1(pp.y, mm.y, vv.y, 0.f, b1c, beta2, b2c, c.x, c.y, eps, 0.f);
        }
        if (APPLY) {
            const float2 c = adam_hist(state, t);
            const float2 gg = *reinterpret_cast<float2 *>(G.g + e);
            sml_adam1(pp.x, mm.x, vv.x, gg.x, b1c, beta2, b2c, c.x, c.y, eps, 0.f);
            sml_adam1(pp.y, mm.y, vv.y, gg.y, b1c, beta2, b2c, c.x, c.y, eps, 0.f);
            *reinterpret_cast<float2 *>(G.g + e) = make_float2(0.f, 0.f);
        }
        *reinterpret_cast<float2 *>(G.p + e) = pp;
        *reinterpret_cast<float2 *>(G.m + e) = mm;
        *reinterpret_cast<float2 *>(G.v + e) = vv;
    }
}

// Every row up to step t: 16 threads (one float4 each) per row, the 16 sit in one warp.
__global__ void __launch_bounds__(256)
k_adam_flush(float4 *__restrict__ p, float4 *__restrict__ m, float4 *__restrict__ v, int32_t *__restrict__ stamp, int64_t n_rows,
             const int64_t *__restrict__ state, float b1c, float beta2, float b2c, float eps) {
    sml_pdl_wait();
    sml_pdl_trigger();
    const int t = (int)state[0];
    const int64_t n4 = n_rows * (SML_D / 4);
    const int64_t n4_up = (n4 + 31) & ~(int64_t)31;      // whole warps stay in the loop for the __syncwarp
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4_up; i += (int64_t)gridDim.x * blockDim.x) {
        const bool live = i < n4;
        const int64_t row = i >> 4;
        const int old = live ? stamp[row] : t;           // SML_STAMP_IDLE (INT32_MAX) never needs anything
        __syncwarp();
        const bool need = old < t;
        float4 pp, mm, vv;
        bool idle = true;
        if (need) {
            pp = p[i]; mm = m[i]; vv = v[i];
            idle = (__float_as_uint(mm.x) | __float_as_uint(mm.y) | __float_as_uint(mm.z) | __float_as_uint(mm.w) |
                    __float_as_uint(vv.x) | __float_as_uint(vv.y) | __float_as_uint(vv.z) | __float_as_uint(vv.w)) == 0u;
        }
        // a row whose moments are all +0 is marked idle: later flushes skip it on its stamp alone (4 B instead of 768 B per row --
        // on a 27 M-row shard of which a step touches 25 k rows this is the difference between 3.7 ms and 0.1 ms per flush)
        const unsigned idle_mask = __ballot_sync(0xffffffffu, idle);
        const bool row_idle = ((idle_mask >> (threadIdx.x & 16)) & 0xFFFFu) == 0xFFFFu;
        if (need) {
            for (int s = idle ? t + 1 : old + 1; s <= t; ++s) {
                const float2 c = adam_hist(state, s);
                sml_adam1(pp.x, mm.x, vv.x, 0.f, b1c, beta2, b2c, c.x, c.y, eps, 0.f);
                sml_adam1(pp.y, mm.y, vv.y, 0.f, b1c, beta2, b2c, c.x, c.y, eps, 0.f);
                sml_adam1(pp.z, mm.z, vv.z, 0.f, b1c, beta2, b2c, c.x, c.y, eps, 0.f);
                sml_adam1(pp.w, mm.w, vv.w, 0.f, b1c, beta2, b2c, c.x, c.y, eps, 0.f);
            }
            if (!idle) { p[i] = pp; m[i] = mm; v[i] = vv; }
            if ((i & 15) == 0) stamp[row] = row_idle ? SML_STAMP_IDLE : t;
        }
    }
}

}  // namespace

int sml_launch_adam_rows(const SmlAdamRows *rows, int n_groups, const int64_t *state, int apply, double beta1, double beta2,
                         double eps, cudaStream_t st) {
    SML_REQUIRE(n_groups >= 1 && n_groups <= 3, SML_E_BADARG, "adam_rows: bad group count %d", n_groups);
    AdamRowsParams P = {};
    int64_t total = 0;
    for (int i = 0; i < n_groups; ++i) {
        P.g[i] = AdamRowsGroup{rows[i].p, rows[i].m, rows[i].v, rows[i].g, rows[i].stamp, rows[i].ids, rows[i].n};
        total += rows[i].n;
    }
    if (total == 0) return SML_OK;
    int64_t blocks = (total + 7) / 8;                       // 8 warps per CTA, one warp per id
    const int64_t cap = (int64_t)sml_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const float b1c = (float)(1.0 - beta1), b2 = (float)beta2, b2c = (float)(1.0 - beta2), e = (float)eps;
    if (apply) SML_CUDA_OK(sml_launch(k_adam_rows<true>, dim3((unsigned)blocks), dim3(256), 0, st, P, state, b1c, b2, b2c, e));
    else SML_CUDA_OK(sml_launch(k_adam_rows<false>, dim3((unsigned)blocks), dim3(256), 0, st, P, state, b1c, b2, b2c, e));
    SML_LAUNCH_OK();
    return SML_OK;
}

// partials: 1024 floats of scratch; ticket: one zeroed unsigned int (re-armed by the kernel)
int sml_launch_sumsq(const float *g, int64_t n, float *sumsq, float *partials, unsigned int *ticket, cudaStream_t st) {
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    if (blocks < 1) blocks = 1;
    SML_CUDA_OK(sml_launch(k_sumsq, dim3((unsigned)blocks), dim3(256), 0, st, (const float4 *)g, n4, sumsq, partials, ticket));
    SML_LAUNCH_OK();
    return SML_OK;
}

extern "C" {

int sml_adam_tick(int64_t *state, double lr, double beta1, double beta2, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(state, SML_E_BADARG, "sml_adam_tick: null state");
    SML_CUDA_OK(sml_launch(k_adam_tick, dim3(1), dim3(32), 0, (cudaStream_t)stream, state, lr, beta1, beta2));
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_sumsq(const float *g, int64_t n, float *sumsq, void *scratch, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(g && sumsq && scratch && n >= 0 && (n % 4) == 0, SML_E_BADARG, "sml_sumsq: bad arguments");
    float *partials = (float *)scratch;
    return sml_launch_sumsq(g, n, sumsq, partials, (unsigned int *)(partials + 1024), (cudaStream_t)stream);
}

int sml_adam_dense(float *p, float *m, float *v, float *g, int64_t n, const int64_t *state, double beta1, double beta2,
                   double eps, double weight_decay, int zero_grad, void *stream) {
    return sml_adam_dense_clipped(p, m, v, g, n, state, beta1, beta2, eps, weight_decay, zero_grad, nullptr, 0.0, stream);
}

int sml_adam_dense_clipped(float *p, float *m, float *v, float *g, int64_t n, const int64_t *state, double beta1, double beta2,
                           double eps, double weight_decay, int zero_grad, const float *sumsq, double max_norm, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(p && m && v && g && state, SML_E_BADARG, "sml_adam_dense: null pointer");
    SML_REQUIRE(n >= 0 && (n % 4) == 0, SML_E_BADARG, "sml_adam_dense: n must be a non-negative multiple of 4");
    SML_REQUIRE((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) & 15) == 0, SML_E_BADARG,
                "sml_adam_dense: pointers must be 16-byte aligned");
    if (n == 0) return SML_OK;
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)sml_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (zero_grad)
        SML_CUDA_OK(sml_launch(k_adam_dense<true>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (float4 *)p, (float4 *)m, (float4 *)v, (float4 *)g, n4,
                               state, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, (float)weight_decay, sumsq, (float)max_norm));
    else
        SML_CUDA_OK(sml_launch(k_adam_dense<false>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (float4 *)p, (float4 *)m, (float4 *)v, (float4 *)g, n4,
                               state, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, (float)weight_decay, sumsq, (float)max_norm));
    SML_LAUNCH_OK();
    return SML_OK;
}


int sml_adam_rows(float *p, float *m, float *v, float *g, int32_t *stamp, const int64_t *ids, int64_t n_ids, int64_t n_rows,
                  const int64_t *state, int apply, double beta1, double beta2, double eps, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(n_ids >= 0 && n_rows >= 0, SML_E_BADARG, "sml_adam_rows: negative size");
    if (n_ids == 0) return SML_OK;
    SML_REQUIRE(p && m && v && stamp && ids && state && (g || !apply), SML_E_BADARG, "sml_adam_rows: null pointer");
    SmlAdamRows r = {p, m, v, g, stamp, ids, n_ids};
    return sml_launch_adam_rows(&r, 1, state, apply, beta1, beta2, eps, (cudaStream_t)stream);
}

int sml_adam_flush(float *p, float *m, float *v, int32_t *stamp, int64_t n_rows, const int64_t *state, double beta1,
                   double beta2, double eps, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(n_rows >= 0, SML_E_BADARG, "sml_adam_flush: negative n_rows");
    if (n_rows == 0) return SML_OK;
    SML_REQUIRE(p && m && v && stamp && state, SML_E_BADARG, "sml_adam_flush: null pointer");
    SML_REQUIRE((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0, SML_E_BADARG, "sml_adam_flush: pointers must be 16-byte aligned");
    const int64_t n4 = n_rows * (SML_D / 4);
    int64_t blocks = (n4 + 255) / 256;
    const int64_t cap = (int64_t)sml_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    SML_CUDA_OK(sml_launch(k_adam_flush, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (float4 *)p, (float4 *)m, (float4 *)v, stamp,
                           n_rows, state, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps));
    SML_LAUNCH_OK();
    return SML_OK;
}

}  // extern "C"
