// Plain matrix-factorisation step (no transfer network): fused 128 B-line row gather, warp-shuffle
// dot products, BCE-mean (model/baseline.py:188-201) or BPR-sum with item biases (MF2.forward,
// model/MF.py:129-147) and an atomic scatter-add of the three row gradients (with the L2 term
// folded in) into the dense gradient tables.  The caller finishes the step with sml_adam_dense.
// HBM-bound: 3 rows read + 3 row-gradient RED.ADDs per triple; one warp per triple.
#include "sml_common.cuh"

namespace {

constexpr int PMF_THREADS = 256;
constexpr int PMF_WARPS = PMF_THREADS / 32;

__global__ void __launch_bounds__(PMF_THREADS)
k_plain_mf(const float *__restrict__ user_tab, const float *__restrict__ item_tab, const float *__restrict__ item_bias,
           const int64_t *__restrict__ user, const int64_t *__restrict__ item, const int64_t *__restrict__ neg, int64_t B,
           int loss_kind, float l2_u, float l2_i, float *__restrict__ g_user, float *__restrict__ g_item,
           float *__restrict__ g_item_bias, float *__restrict__ loss_out, float *__restrict__ partials,
           unsigned int *__restrict__ ticket) {
    __shared__ float s_part[PMF_WARPS][2];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float invB = 1.0f / (float)B;
    float acc_loss = 0.f, acc_l2 = 0.f;
    for (int64_t b = (int64_t)blockIdx.x * PMF_WARPS + w; b < B; b += (int64_t)gridDim.x * PMF_WARPS) {
        const int64_t iu = __ldg(user + b), ii = __ldg(item + b), ij = __ldg(neg + b);
        const float *pu = user_tab + iu * SML_D, *pi = item_tab + ii * SML_D, *pj = item_tab + ij * SML_D;
        const float u[2] = {__ldg(pu + lane), __ldg(pu + lane + 32)};
        const float vi[2] = {__ldg(pi + lane), __ldg(pi + lane + 32)};
        const float vj[2] = {__ldg(pj + lane), __ldg(pj + lane + 32)};
        const float sp = warp_sum(fmaf(u[1], vi[1], u[0] * vi[0]));
        const float sn = warp_sum(fmaf(u[1], vj[1], u[0] * vj[0]));
        float dsp, dsn;
        if (loss_kind == SML_LOSS_BCE) {
            const float gp = sml_sigmoid(sp), gn = sml_sigmoid(sn);
            const float ap = gp + 1e-15f, an = (1.0f - gn) + 1e-15f;
            if (lane == 0) acc_loss += logf(ap) + logf(an);
            dsp = -(gp * (1.0f - gp)) / ap * invB;
            dsn = (gn * (1.0f - gn)) / an * invB;
            const float qu = warp_sum(u[0] * u[0] + u[1] * u[1]);
            const float qi = warp_sum(vi[0] * vi[0] + vi[1] * vi[1] + vj[0] * vj[0] + vj[1] * vj[1]);
            if (lane == 0) acc_l2 += l2_u * 0.5f * qu + l2_i * 0.5f * qi;
        } else {
            float x = sp - sn;
            if (item_bias) x += __ldg(item_bias + ii) - __ldg(item_bias + ij);   // user bias cancels (MF.py:141-143)
            if (lane == 0) acc_loss += fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
            dsp = -sml_sigmoid(-x);
            dsn = -dsp;
            if (g_item_bias && lane == 0) { atomicAdd(g_item_bias + ii, dsp); atomicAdd(g_item_bias + ij, dsn); }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = lane + 32 * h;
            atomicAdd(g_user + iu * SML_D + k, fmaf(l2_u, u[h], dsp * vi[h] + dsn * vj[h]));
            atomicAdd(g_item + ii * SML_D + k, fmaf(l2_i, vi[h], dsp * u[h]));
            atomicAdd(g_item + ij * SML_D + k, fmaf(l2_i, vj[h], dsn * u[h]));
        }
    }
    if (lane == 0) { s_part[w][0] = acc_loss; s_part[w][1] = acc_l2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, q = 0.f;
        for (int i = 0; i < PMF_WARPS; ++i) { a += s_part[i][0]; q += s_part[i][1]; }
        partials[2 * blockIdx.x] = a; partials[2 * blockIdx.x + 1] = q;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        float a = 0.f, q = 0.f;
        for (unsigned i = 0; i < gridDim.x; ++i) { a += __ldcg(partials + 2 * i); q += __ldcg(partials + 2 * i + 1); }
        const float loss = (loss_kind == SML_LOSS_BCE) ? (-(a * invB) + q) : a;
        loss_out[0] = loss;
        loss_out[1] += loss;
        *ticket = 0;
    }
}

}  // namespace

extern "C" int sml_plain_mf_grads(const float *user_tab, const float *item_tab, const float *item_bias, const int64_t *user,
                                  const int64_t *item, const int64_t *neg, int64_t batch, int d, int loss, double l2_u,
                                  double l2_i, float *g_user, float *g_item, float *g_item_bias, float *loss_out,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_plain_mf_grads: d=%d unsupported (d must be %d)", d, SML_D);
    SML_REQUIRE(user_tab && item_tab && user && item && neg && g_user && g_item && loss_out, SML_E_BADARG,
                "sml_plain_mf_grads: null pointer");
    SML_REQUIRE(loss == SML_LOSS_BCE || loss == SML_LOSS_BPR, SML_E_BADARG, "sml_plain_mf_grads: bad loss kind %d", loss);
    SML_REQUIRE(batch > 0, SML_E_BADARG, "sml_plain_mf_grads: batch must be positive");
    SML_REQUIRE(workspace && workspace_bytes >= 256 + 2 * 2048 * sizeof(float), SML_E_WORKSPACE,
                "sml_plain_mf_grads: workspace must be >= %zu bytes (zero-initialised)", 256 + 2 * 2048 * sizeof(float));
    int64_t blocks = (batch + PMF_WARPS - 1) / PMF_WARPS;
    const int64_t cap = (int64_t)sml_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks > 2048) blocks = 2048;
    unsigned int *ticket = (unsigned int *)workspace;
    float *partials = (float *)((char *)workspace + 256);
    k_plain_mf<<<(int)blocks, PMF_THREADS, 0, (cudaStream_t)stream>>>(user_tab, item_tab, item_bias, user, item, neg, batch, loss,
                                                                      (float)l2_u, (float)l2_i, g_user, g_item, g_item_bias,
                                                                      loss_out, partials, ticket);
    SML_LAUNCH_OK();
    return SML_OK;
}
