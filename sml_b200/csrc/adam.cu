// torch.optim.Adam semantics (model/transfer.py:392-393 of the reference: lr, betas (0.9, 0.999),
// eps 1e-8, amsgrad off, coupled L2 weight decay) as one streaming kernel.
//
// The MF optimizer in the reference is DENSE Adam over whole nn.Embedding tables
// (nn.Embedding(sparse=False), model/MF.py:21-24): rows touched at an earlier step keep moving
// through the momentum tail, so a lazy/sparse Adam is not equivalent.  The dense sweep is
// HBM-bound: per element read p, m, v, g and write p, m, v (+ the zeroed g that replaces
// zero_grad()): 32 B per float, all 128-bit accesses, grid sized to the SM count.
#include <math.h>

#include "sml_common.cuh"

namespace {

__global__ void k_adam_tick(int64_t *state, double lr, double beta1, double beta2) {
    sml_pdl_wait();
    sml_pdl_trigger();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int64_t t = state[0] + 1;
        state[0] = t;
        const double bc1 = 1.0 - pow(beta1, (double)t);
        const double bc2 = 1.0 - pow(beta2, (double)t);
        float *f = reinterpret_cast<float *>(state + 1);
        f[0] = (float)(lr / bc1);          // step_size
        f[1] = (float)sqrt(bc2);           // bias_correction2_sqrt
    }
}

__device__ __forceinline__ void adam1(float &p, float &m, float &v, float g, float b1c, float beta2, float b2c,
                                      float step_size, float bc2_sqrt, float eps, float wd) {
    if (wd != 0.f) g = fmaf(wd, p, g);                 // grad.add(param, alpha=weight_decay)
    m = fmaf(g - m, b1c, m);                           // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(b2c * g, g, beta2 * v);                   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);                   // param.addcdiv_(exp_avg, denom, value=-step_size)
}

template <bool ZERO>
__global__ void __launch_bounds__(256)
k_adam_dense(float4 *__restrict__ p, float4 *__restrict__ m, float4 *__restrict__ v, float4 *__restrict__ g, int64_t n4,
             const int64_t *__restrict__ state, float b1c, float beta2, float b2c, float eps, float wd) {
    sml_pdl_wait();
    sml_pdl_trigger();
    // b1c = (float)(1 - beta1), b2c = (float)(1 - beta2) are rounded from the double differences on the
    // host, as torch does (1.0f - 0.999f would be off by 5e-5 relative)
    const float *f = reinterpret_cast<const float *>(state + 1);
    const float step_size = f[0], bc2_sqrt = f[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = p[i], mm = m[i], vv = v[i];
        const float4 gg = g[i];
        adam1(pp.x, mm.x, vv.x, gg.x, b1c, beta2, b2c, step_size, bc2_sqrt, eps, wd);
        adam1(pp.y, mm.y, vv.y, gg.y, b1c, beta2, b2c, step_size, bc2_sqrt, eps, wd);
        adam1(pp.z, mm.z, vv.z, gg.z, b1c, beta2, b2c, step_size, bc2_sqrt, eps, wd);
        adam1(pp.w, mm.w, vv.w, gg.w, b1c, beta2, b2c, step_size, bc2_sqrt, eps, wd);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (ZERO) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace

extern "C" {

int sml_adam_tick(int64_t *state, double lr, double beta1, double beta2, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(state, SML_E_BADARG, "sml_adam_tick: null state");
    SML_CUDA_OK(sml_launch(k_adam_tick, dim3(1), dim3(32), 0, (cudaStream_t)stream, state, lr, beta1, beta2));
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_adam_dense(float *p, float *m, float *v, float *g, int64_t n, const int64_t *state, double beta1, double beta2,
                   double eps, double weight_decay, int zero_grad, void *stream) {
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
                               state, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, (float)weight_decay));
    else
        SML_CUDA_OK(sml_launch(k_adam_dense<false>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (float4 *)p, (float4 *)m, (float4 *)v, (float4 *)g, n4,
                               state, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, (float)weight_decay));
    SML_LAUNCH_OK();
    return SML_OK;
}

}  // extern "C"
