// Host-side sequencing of the kernels behind the step-level C ABI (no device code here).
// Every function only enqueues work on the caller's stream: no allocation, no synchronisation,
// so a whole step (and a whole epoch of steps) can be captured into one CUDA graph.
#include "sml_common.cuh"

namespace {

constexpr int64_t FWD_CHUNK = 8192;   // rows per pass of the SIMT transfer forward

// fc GEMM dispatch: tcgen05 3xTF32 by default, SIMT fp32 when SML_GEMM=simt
int gemm(const SmlGemmProb *probs, int n, int a_mode, int b_mode, int epi, int bn, cudaStream_t st) {
    if (sml_use_tensor_cores()) return sml_launch_umma_gemm(probs, n, a_mode, b_mode, epi, 0, bn, st);
    return sml_launch_sgemm(probs, n, a_mode, b_mode, epi, st);
}

struct StepWs {
    unsigned int *ticket;   // [64] (only [0] used), re-armed by k_loss
    float *partials;        // [3 * 1024]
    float *A, *Z1, *Y, *dY, *dZ1, *dA, *rowsq;
};

size_t step_ws_bytes(int64_t B) {
    const size_t N = (size_t)3 * B;
    return 256 + 3 * 1024 * sizeof(float) + N * (320 + 512 + 64 + 64 + 512 + 320 + 1) * sizeof(float) + 7 * 256;
}

StepWs carve(void *ws, int64_t B) {
    StepWs w;
    char *p = (char *)ws;
    auto take = [&](size_t bytes) { char *r = p; p += sml_align_up(bytes, 256); return r; };
    const size_t N = (size_t)3 * B;
    w.ticket = (unsigned int *)take(256);
    w.partials = (float *)take(3 * 1024 * sizeof(float));
    w.A = (float *)take(N * 320 * sizeof(float));
    w.Z1 = (float *)take(N * 512 * sizeof(float));
    w.Y = (float *)take(N * 64 * sizeof(float));
    w.dY = (float *)take(N * 64 * sizeof(float));
    w.dZ1 = (float *)take(N * 512 * sizeof(float));
    w.dA = (float *)take(N * 320 * sizeof(float));
    w.rowsq = (float *)take(N * sizeof(float));
    return w;
}

int check_args(const sml_step_args *a, const char *who) {
    SML_REQUIRE(a, SML_E_BADARG, "%s: null args", who);
    SML_REQUIRE(a->batch > 0, SML_E_BADARG, "%s: batch must be positive", who);
    SML_REQUIRE(a->user && a->item && a->neg, SML_E_BADARG, "%s: null id pointer", who);
    SML_REQUIRE(a->last_user && a->last_item && a->hat_user && a->hat_item && a->theta, SML_E_BADARG,
                "%s: null table/theta pointer", who);
    SML_REQUIRE(a->variant == SML_VARIANT_COM || a->variant == SML_VARIANT_CONV, SML_E_BADARG, "%s: bad variant %d",
                who, a->variant);
    SML_REQUIRE(a->loss == SML_LOSS_BCE || a->loss == SML_LOSS_BPR, SML_E_BADARG, "%s: bad loss kind %d", who, a->loss);
    SML_REQUIRE(a->workspace && a->workspace_bytes >= step_ws_bytes(a->batch), SML_E_WORKSPACE,
                "%s: workspace too small (%zu < %zu bytes)", who, a->workspace_bytes, step_ws_bytes(a->batch));
    SML_REQUIRE(a->loss_out, SML_E_BADARG, "%s: null loss_out", who);
    return SML_OK;
}

void make_groups(const sml_step_args *a, SmlRowGroup g[3]) {
    const int64_t B = a->batch;
    const float *tu = a->theta, *ti = a->theta + SML_NET_STRIDE;
    g[0] = SmlRowGroup{a->last_user, a->hat_user, a->user, tu, B, 0};
    g[1] = SmlRowGroup{a->last_item, a->hat_item, a->item, ti, B, B};
    g[2] = SmlRowGroup{a->last_item, a->hat_item, a->neg, ti, B, 2 * B};
}

// forward through fc2 + loss + data gradients down to dZ1 (shared by all step flavours)
int forward_and_loss(const sml_step_args *a, const StepWs &w, bool want_rowsq, float l2, float *scores, cudaStream_t st) {
    const int64_t B = a->batch;
    const float *tu = a->theta, *ti = a->theta + SML_NET_STRIDE;
    SmlRowGroup g[3];
    make_groups(a, g);
    int rc = sml_launch_conv_fwd(g, 3, a->variant, w.A, want_rowsq ? w.rowsq : nullptr, st);
    if (rc) return rc;
    // fc1: Z1 = A W1^T + b1   (conv_transfer.py:47), user rows with the user net, item rows with the item net
    SmlGemmProb fc1[2] = {
        {w.A, tu + SML_OFF_F1W, tu + SML_OFF_F1B, nullptr, w.Z1, (int)B, 512, 320, 320, 320, 512},
        {w.A + B * 320, ti + SML_OFF_F1W, ti + SML_OFF_F1B, nullptr, w.Z1 + B * 512, (int)(2 * B), 512, 320, 320, 320, 512}};
    rc = gemm(fc1, 2, SML_A_MK, SML_B_NK, SML_EPI_BIAS, 128, st);
    if (rc) return rc;
    // fc2: Y = g(Z1) W2^T + b2   (:48-49)
    SmlGemmProb fc2[2] = {
        {w.Z1, tu + SML_OFF_F2W, tu + SML_OFF_F2B, nullptr, w.Y, (int)B, 64, 512, 512, 512, 64},
        {w.Z1 + B * 512, ti + SML_OFF_F2W, ti + SML_OFF_F2B, nullptr, w.Y + B * 64, (int)(2 * B), 64, 512, 512, 512, 64}};
    rc = gemm(fc2, 2, SML_A_MK_GELU, SML_B_NK, SML_EPI_BIAS, 64, st);
    if (rc) return rc;
    rc = sml_launch_loss(w.Y, want_rowsq ? w.rowsq : nullptr, B, a->loss, a->variant == SML_VARIANT_CONV, l2, w.dY, scores,
                         a->loss_out, w.partials, w.ticket, st);
    if (rc) return rc;
    // dZ1 = (dY W2) * g'(Z1)
    SmlGemmProb d2[2] = {
        {w.dY, tu + SML_OFF_F2W, nullptr, w.Z1, w.dZ1, (int)B, 512, 64, 64, 512, 512},
        {w.dY + B * 64, ti + SML_OFF_F2W, nullptr, w.Z1 + B * 512, w.dZ1 + B * 512, (int)(2 * B), 512, 64, 64, 512, 512}};
    return gemm(d2, 2, SML_A_MK, SML_B_KN, SML_EPI_MUL_GELU_GRAD, 128, st);
}

int fc1_dgrad(const sml_step_args *a, const StepWs &w, cudaStream_t st) {
    const int64_t B = a->batch;
    const float *tu = a->theta, *ti = a->theta + SML_NET_STRIDE;
    SmlGemmProb d1[2] = {
        {w.dZ1, tu + SML_OFF_F1W, nullptr, nullptr, w.dA, (int)B, 320, 512, 512, 320, 320},
        {w.dZ1 + B * 512, ti + SML_OFF_F1W, nullptr, nullptr, w.dA + B * 320, (int)(2 * B), 320, 512, 512, 320, 320}};
    return gemm(d1, 2, SML_A_MK, SML_B_KN, SML_EPI_NONE, 64, st);
}

// fc1/fc2 weight + bias gradients accumulated into g_theta  (theta grads of conv_transfer.py:47-49)
int fc_wgrads(const sml_step_args *a, const StepWs &w, float *g_theta, cudaStream_t st) {
    const int64_t B = a->batch;
    float *gu = g_theta, *gi = g_theta + SML_NET_STRIDE;
    SmlGemmProb w2[2] = {
        {w.dY, w.Z1, nullptr, nullptr, gu + SML_OFF_F2W, 64, 512, (int)B, 64, 512, 512},
        {w.dY + B * 64, w.Z1 + B * 512, nullptr, nullptr, gi + SML_OFF_F2W, 64, 512, (int)(2 * B), 64, 512, 512}};
    int rc;
    if (sml_use_tensor_cores()) {
        // tensor-core tiles are 128 rows tall: compute dW2^T [512, 64] = g(Z1)^T dY and store it transposed
        SmlGemmProb w2t[2] = {
            {w.Z1, w.dY, nullptr, nullptr, gu + SML_OFF_F2W, 512, 64, (int)B, 512, 64, 512},
            {w.Z1 + B * 512, w.dY + B * 64, nullptr, nullptr, gi + SML_OFF_F2W, 512, 64, (int)(2 * B), 512, 64, 512}};
        rc = sml_launch_umma_gemm(w2t, 2, SML_A_KM_GELU, SML_B_KN, SML_EPI_ACCUM, 1, 64, st);
    } else {
        rc = sml_launch_sgemm(w2, 2, SML_A_KM, SML_B_KN_GELU, SML_EPI_ACCUM, st);
    }
    if (rc) return rc;
    SmlGemmProb w1[2] = {
        {w.dZ1, w.A, nullptr, nullptr, gu + SML_OFF_F1W, 512, 320, (int)B, 512, 320, 320},
        {w.dZ1 + B * 512, w.A + B * 320, nullptr, nullptr, gi + SML_OFF_F1W, 512, 320, (int)(2 * B), 512, 320, 320}};
    rc = gemm(w1, 2, SML_A_KM, SML_B_KN, SML_EPI_ACCUM, 64, st);
    if (rc) return rc;
    SmlColsumProb cs[4] = {{w.dY, gu + SML_OFF_F2B, (int)B, 64, 64},
                           {w.dY + B * 64, gi + SML_OFF_F2B, (int)(2 * B), 64, 64},
                           {w.dZ1, gu + SML_OFF_F1B, (int)B, 512, 512},
                           {w.dZ1 + B * 512, gi + SML_OFF_F1B, (int)(2 * B), 512, 512}};
    return sml_launch_colsum(cs, 4, st);
}

}  // namespace

extern "C" {

size_t sml_step_workspace_bytes(int64_t batch) { return batch > 0 ? step_ws_bytes(batch) : 0; }

size_t sml_transfer_fwd_workspace_bytes(int64_t n_rows) {
    const int64_t ch = n_rows < FWD_CHUNK ? n_rows : FWD_CHUNK;
    return ch > 0 ? (size_t)ch * (320 + 512) * sizeof(float) + 512 : 0;
}

int sml_transfer_fwd(const float *x_t, const float *x_hat, const int64_t *ids, int64_t n_rows, int d, int variant,
                     const float *theta_net, int normalize_out, float *out, void *workspace, size_t workspace_bytes,
                     void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_transfer_fwd: d=%d unsupported (d must be %d)", d, SML_D);
    SML_REQUIRE(variant == SML_VARIANT_COM || variant == SML_VARIANT_CONV, SML_E_BADARG, "sml_transfer_fwd: bad variant %d", variant);
    SML_REQUIRE(x_t && x_hat && theta_net && out, SML_E_BADARG, "sml_transfer_fwd: null pointer");
    SML_REQUIRE(n_rows >= 0, SML_E_BADARG, "sml_transfer_fwd: negative n_rows");
    if (n_rows == 0) return SML_OK;
    SML_REQUIRE(workspace && workspace_bytes >= sml_transfer_fwd_workspace_bytes(n_rows), SML_E_WORKSPACE,
                "sml_transfer_fwd: workspace too small (%zu < %zu bytes)", workspace_bytes,
                sml_transfer_fwd_workspace_bytes(n_rows));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ch = n_rows < FWD_CHUNK ? n_rows : FWD_CHUNK;
    float *A = (float *)workspace;
    float *Z1 = A + sml_align_up((size_t)ch * 320, 64);
    for (int64_t r0 = 0; r0 < n_rows; r0 += ch) {
        const int64_t n = (n_rows - r0) < ch ? (n_rows - r0) : ch;
        SmlRowGroup g;
        if (ids) g = SmlRowGroup{x_t, x_hat, ids + r0, theta_net, n, 0};
        else g = SmlRowGroup{x_t + r0 * SML_D, x_hat + r0 * SML_D, nullptr, theta_net, n, 0};
        rc = sml_launch_conv_fwd(&g, 1, variant, A, nullptr, st);
        if (rc) return rc;
        SmlGemmProb fc1 = {A, theta_net + SML_OFF_F1W, theta_net + SML_OFF_F1B, nullptr, Z1, (int)n, 512, 320, 320, 320, 512};
        rc = gemm(&fc1, 1, SML_A_MK, SML_B_NK, SML_EPI_BIAS, 128, st);
        if (rc) return rc;
        SmlGemmProb fc2 = {Z1, theta_net + SML_OFF_F2W, theta_net + SML_OFF_F2B, nullptr, out + r0 * SML_D, (int)n, 64, 512, 512, 512, 64};
        rc = gemm(&fc2, 1, SML_A_MK_GELU, SML_B_NK, SML_EPI_BIAS, 64, st);
        if (rc) return rc;
    }
    if (normalize_out) return sml_launch_row_normalize(out, n_rows, st);
    return SML_OK;
}

int sml_mf_step(const sml_step_args *a, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    rc = check_args(a, "sml_mf_step");
    if (rc) return rc;
    SML_REQUIRE(a->g_user && a->g_item && a->m_user && a->v_user && a->m_item && a->v_item && a->adam_state, SML_E_BADARG,
                "sml_mf_step: null gradient / Adam-state pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const StepWs w = carve(a->workspace, a->batch);
    rc = sml_adam_tick(a->adam_state, a->lr, 0.9, 0.999, stream);
    if (rc) return rc;
    rc = forward_and_loss(a, w, true, (float)a->l2, nullptr, st);
    if (rc) return rc;
    rc = fc1_dgrad(a, w, st);
    if (rc) return rc;
    SmlRowGroup g[3];
    make_groups(a, g);
    SmlConvBwdGroup bg[3] = {{g[0], a->g_user, nullptr}, {g[1], a->g_item, nullptr}, {g[2], a->g_item, nullptr}};
    rc = sml_launch_conv_bwd(bg, 3, a->variant, w.dA, (float)a->l2, nullptr, st);
    if (rc) return rc;
    // dense Adam on both latent tables, weight_decay = 0 (model/transfer.py:392); also re-zeroes the gradients
    rc = sml_adam_dense(a->hat_user, a->m_user, a->v_user, a->g_user, a->n_users * SML_D, a->adam_state, 0.9, 0.999, 1e-8, 0.0, 1, stream);
    if (rc) return rc;
    return sml_adam_dense(a->hat_item, a->m_item, a->v_item, a->g_item, a->n_items * SML_D, a->adam_state, 0.9, 0.999, 1e-8, 0.0, 1, stream);
}

int sml_tr_step(const sml_step_args *a, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    rc = check_args(a, "sml_tr_step");
    if (rc) return rc;
    SML_REQUIRE(a->g_theta && a->m_theta && a->v_theta && a->adam_state, SML_E_BADARG, "sml_tr_step: null theta-gradient / Adam-state pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const StepWs w = carve(a->workspace, a->batch);
    rc = sml_adam_tick(a->adam_state, a->lr, 0.9, 0.999, stream);
    if (rc) return rc;
    rc = forward_and_loss(a, w, false, 0.f, nullptr, st);
    if (rc) return rc;
    rc = fc_wgrads(a, w, a->g_theta, st);
    if (rc) return rc;
    rc = fc1_dgrad(a, w, st);
    if (rc) return rc;
    SmlRowGroup g[3];
    make_groups(a, g);
    float *gu = a->g_theta, *gi = a->g_theta + SML_NET_STRIDE;
    SmlConvBwdGroup bg[3] = {{g[0], nullptr, gu}, {g[1], nullptr, gi}, {g[2], nullptr, gi}};
    rc = sml_launch_conv_bwd(bg, 3, a->variant, w.dA, 0.f, nullptr, st);
    if (rc) return rc;
    // Adam with coupled L2 (weight_decay = TR_l2, model/transfer.py:393) over the whole theta block
    return sml_adam_dense(a->theta, a->m_theta, a->v_theta, a->g_theta, 2 * (int64_t)SML_NET_STRIDE, a->adam_state, 0.9, 0.999,
                          1e-8, a->l2, 1, stream);
}

int sml_run_mf_grads(const sml_step_args *a, float *d_rows, float *scores, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    rc = check_args(a, "sml_run_mf_grads");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const StepWs w = carve(a->workspace, a->batch);
    rc = forward_and_loss(a, w, false, 0.f, scores, st);
    if (rc) return rc;
    if (a->g_theta) {
        rc = fc_wgrads(a, w, a->g_theta, st);
        if (rc) return rc;
    }
    if (!d_rows && !a->g_theta) return SML_OK;
    rc = fc1_dgrad(a, w, st);
    if (rc) return rc;
    SmlRowGroup g[3];
    make_groups(a, g);
    float *gu = a->g_theta, *gi = a->g_theta ? a->g_theta + SML_NET_STRIDE : nullptr;
    SmlConvBwdGroup bg[3] = {{g[0], nullptr, gu}, {g[1], nullptr, gi}, {g[2], nullptr, gi}};
    return sml_launch_conv_bwd(bg, 3, a->variant, w.dA, 0.f, d_rows, st);
}

int sml_debug_gemm(const float *A, const float *B, const float *bias, const float *aux, float *C, int M, int N, int K, int lda,
                   int ldb, int ldc, int a_mode, int b_mode, int epi, int transpose_out, int bn, int tensor_cores, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SmlGemmProb p = {A, B, bias, aux, C, M, N, K, lda, ldb, ldc};
    if (tensor_cores) return sml_launch_umma_gemm(&p, 1, a_mode, b_mode, epi, transpose_out, bn, (cudaStream_t)stream);
    SML_REQUIRE(!transpose_out, SML_E_BADARG, "sml_debug_gemm: the SIMT kernel has no transposed store");
    return sml_launch_sgemm(&p, 1, a_mode, b_mode, epi, (cudaStream_t)stream);
}

int sml_mf_epoch(const sml_step_args *a, int64_t n_total, void *stream) {
    SML_REQUIRE(a && n_total >= 0, SML_E_BADARG, "sml_mf_epoch: bad arguments");
    sml_step_args s = *a;
    for (int64_t off = 0; off < n_total; off += a->batch) {
        s.user = a->user + off; s.item = a->item + off; s.neg = a->neg + off;
        s.batch = (n_total - off) < a->batch ? (n_total - off) : a->batch;
        int rc = sml_mf_step(&s, stream);
        if (rc) return rc;
    }
    return SML_OK;
}

int sml_tr_epoch(const sml_step_args *a, int64_t n_total, void *stream) {
    SML_REQUIRE(a && n_total >= 0, SML_E_BADARG, "sml_tr_epoch: bad arguments");
    sml_step_args s = *a;
    for (int64_t off = 0; off < n_total; off += a->batch) {
        s.user = a->user + off; s.item = a->item + off; s.neg = a->neg + off;
        s.batch = (n_total - off) < a->batch ? (n_total - off) : a->batch;
        int rc = sml_tr_step(&s, stream);
        if (rc) return rc;
    }
    return SML_OK;
}

}  // extern "C"
