// Host-side sequencing of the kernels behind the step-level C ABI (no device code here).
// Every function only enqueues work on the caller's stream: no allocation, no synchronisation,
// so a whole step (and a whole epoch of steps) can be captured into one CUDA graph.
//
// Row layout of every per-step matrix (plain [rows, .] and packed 128-row tiles alike):
//     user rows      [0, B)                     user net
//     positive items [Bp, Bp + B)               item net      Bp = B rounded up to 128
//     negative items [Bp + B, Bp + 2B)          item net
// so that no 128-row tensor-core tile straddles the two nets and the item rows stay contiguous
// (one K range for the item net's weight gradients).
#include "sml_common.cuh"
#include "umma_pack.cuh"

namespace {

constexpr int64_t FWD_CHUNK = 16384;   // rows per pass of the full-table transfer forward

inline int64_t up128(int64_t x) { return (x + 127) / 128 * 128; }

struct Rows {
    int64_t B, Bp, R;          // batch, padded user segment, total padded rows
    int user_tiles, item_tiles;
};
Rows rows_of(int64_t B) {
    Rows r;
    r.B = B; r.Bp = up128(B); r.R = r.Bp + up128(2 * B);
    r.user_tiles = (int)(r.Bp / 128); r.item_tiles = (int)(up128(2 * B) / 128);
    return r;
}

struct StepWs {
    unsigned int *ticket;   // [64] (only [0] used), re-armed by k_loss
    float *partials;        // [4 * 1024]
    float *A, *Z1, *Y, *dY, *dZ1, *dA, *rowsq;          // plain, R rows
    uint8_t *Apk, *Gpk, *dYpk, *dZpk;                   // packed operands (128-row tiles over the R rows)
    uint8_t *theta_pk;                                  // packed weights, 2 nets
};

size_t step_ws_bytes(int64_t B) {
    const Rows r = rows_of(B);
    const size_t R = (size_t)r.R, T = R / 128;
    size_t n = 256 + 4 * 1024 * sizeof(float) + R * (320 + 512 + 64 + 64 + 512 + 320 + 1) * sizeof(float);
    n += T * (size_t)(10 + 16 + 2 + 16) * pk_block_bytes(128) + 2 * SML_PK_THETA_BYTES;
    return n + 16 * 256;
}

StepWs carve(void *ws, int64_t B) {
    StepWs w;
    char *p = (char *)ws;
    auto take = [&](size_t bytes) { char *q = p; p += sml_align_up(bytes, 256); return q; };
    const Rows r = rows_of(B);
    const size_t R = (size_t)r.R, T = R / 128;
    w.ticket = (unsigned int *)take(256);
    w.partials = (float *)take(4 * 1024 * sizeof(float));
    w.A = (float *)take(R * 320 * sizeof(float));
    w.Z1 = (float *)take(R * 512 * sizeof(float));
    w.Y = (float *)take(R * 64 * sizeof(float));          // Y and dA are adjacent: one memset clears both when the
    w.dA = (float *)take(R * 320 * sizeof(float));        // split-K GEMMs accumulate into them
    w.dY = (float *)take(R * 64 * sizeof(float));
    w.dZ1 = (float *)take(R * 512 * sizeof(float));
    w.rowsq = (float *)take(R * sizeof(float));
    w.Apk = (uint8_t *)take(T * 10 * pk_block_bytes(128));
    w.Gpk = (uint8_t *)take(T * 16 * pk_block_bytes(128));
    w.dYpk = (uint8_t *)take(T * 2 * pk_block_bytes(128));
    w.dZpk = (uint8_t *)take(T * 16 * pk_block_bytes(128));
    w.theta_pk = (uint8_t *)take(2 * SML_PK_THETA_BYTES);
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
    SML_REQUIRE(a->table_pitch == 0 || (a->table_pitch >= SML_D && a->table_pitch % 4 == 0), SML_E_BADARG, "%s: bad table_pitch", who);
    return SML_OK;
}

void make_groups(const sml_step_args *a, SmlRowGroup g[3]) {
    const Rows r = rows_of(a->batch);
    const float *tu = a->theta, *ti = a->theta + SML_NET_STRIDE;
    const int64_t pitch = a->table_pitch > 0 ? a->table_pitch : SML_D;
    g[0] = SmlRowGroup{a->last_user, a->hat_user, a->user, tu, r.B, 0, pitch};
    g[1] = SmlRowGroup{a->last_item, a->hat_item, a->item, ti, r.B, r.Bp, pitch};
    g[2] = SmlRowGroup{a->last_item, a->hat_item, a->neg, ti, r.B, r.Bp + r.B, pitch};
}

// K = 512 GEMMs (fc2, d1) of small batches have only a handful of output tiles: slice K so they cover more SMs
// (measured: 119.3 -> 117.0 us for the B = 256 transfer step; at B = 1024 the atomics cost more than the slicing gains)
// (d1, which = 1: with the two weight-gradient GEMMs on their own streams the tail of the transfer step is bound by the number of
// one-CTA-per-SM tiles in flight, and an unsliced d1 -- 30 tiles instead of 120, no atomics -- wins: tools/ksplit_sweep.py, 65.1 -> 61.8 us
// together with dW1 in 2 slices)
int step_ksplit(const Rows &r, int which = 0) {
    if (sml_debug_ksplit(which) > 0) return sml_debug_ksplit(which);
    if (which == 1) return 1;
    return (r.user_tiles + r.item_tiles) <= 12 ? 4 : 1;
}

// two-net problem pair for the packed GEMMs
void pk_pair(SmlPkProb out[2], const Rows &r, const uint8_t *A, int KC, size_t w_off, int N, const float *theta, size_t bias_off,
             const float *aux, float *C, int ldc, uint8_t *Cpk, const uint8_t *theta_pk) {
    for (int net = 0; net < 2; ++net) {
        SmlPkProb &p = out[net];
        p.A = A; p.B = theta_pk + net * SML_PK_THETA_BYTES + w_off; p.KC = KC;
        p.m_tiles = net ? r.item_tiles : r.user_tiles;
        p.a_tile0 = net ? r.user_tiles : 0;
        p.row0 = net ? r.Bp : 0;
        p.M = (int)(net ? 2 * r.B : r.B);
        p.N = N;
        p.bias = bias_off ? theta + (size_t)net * SML_NET_STRIDE + bias_off : nullptr;
        p.aux = aux; p.C = C; p.ldc = ldc; p.Cpk = Cpk; p.c_tile0 = p.a_tile0; p.colsum = nullptr;
        p.B2 = nullptr; p.bias2 = nullptr; p.Y = nullptr; p.ldy = 0;
    }
}

// Fork / join helper: the two weight-gradient GEMMs of the transfer step only feed the final Adam update, so they run
// on a library-owned side stream concurrently with dA = dZ1 W1 and the conv backward (21 vs 18 us at B = 256,
// profiles/r01_tr_step_breakdown.md).  Event record / wait are stream-capturable, so inside a CUDA graph this becomes
// two parallel branches.
struct SideStream {
    cudaStream_t s = nullptr, s2 = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, fork2 = nullptr, join2 = nullptr, join3 = nullptr;
};
SideStream *side_stream() {
    static SideStream per_dev[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideStream &x = per_dev[dev];
    if (!x.s) {
        if (cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&x.s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&x.join3, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&x.fork2, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&x.join2, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &x;
}

// forward through fc2 + loss + data gradients down to dZ1 (shared by all step flavours).
// need_plain_A: the transfer step's fc1 weight gradient reads A.
// g_theta != null (tensor-core path): the fc2 / fc1 bias gradients are accumulated by the loss kernel and by the d2
// epilogue.  tick_state != null: the Adam tick rides in the theta packer's launch.
int forward_and_loss(const sml_step_args *a, const StepWs &w, bool want_rowsq, bool need_plain_A, bool pack_theta, float l2,
                     float *scores, cudaStream_t st, float *g_theta = nullptr, int64_t *tick_state = nullptr, double tick_lr = 0.0,
                     float adaptive = 0.f) {
    const Rows r = rows_of(a->batch);
    const int64_t B = r.B;
    const float *tu = a->theta, *ti = a->theta + SML_NET_STRIDE;
    const bool tc = sml_use_tensor_cores() != 0;
    SmlRowGroup g[3];
    make_groups(a, g);
    int rc;
    SideStream *ss = (tc && pack_theta && tick_state) ? side_stream() : nullptr;   // transfer step: pack theta || conv prologue
    if (ss) {
        SML_CUDA_OK(cudaEventRecord(ss->fork2, st));
        SML_CUDA_OK(cudaStreamWaitEvent(ss->s, ss->fork2, 0));
        rc = sml_launch_pack_theta(a->theta, w.theta_pk, 2, ss->s, tick_state, tick_lr);
        if (rc) return rc;
        SML_CUDA_OK(cudaEventRecord(ss->join2, ss->s));
    } else if (tc && pack_theta) {
        rc = sml_launch_pack_theta(a->theta, w.theta_pk, 2, st, tick_state, tick_lr);
        if (rc) return rc;
    } else if (tick_state) {
        rc = sml_adam_tick(tick_state, tick_lr, 0.9, 0.999, st);
        if (rc) return rc;
    }
    float *gbu2 = (tc && g_theta) ? g_theta + SML_OFF_F2B : nullptr, *gbi2 = (tc && g_theta) ? g_theta + SML_NET_STRIDE + SML_OFF_F2B : nullptr;
    // split-K slices accumulate into Y (fc2) and dA (d1): the conv prologue and the loss kernel clear the batch rows on their
    // way (a memset node in the chain would cost a launch and break the programmatic-dependent-launch overlap)
    // fc2 fused into the fc1 tiles (umma_packed.cu, SML_PK_FC1_FC2) when the batch is a handful of row tiles (the 768-row transfer
    // step: 70.8 -> 68.0 us; at 3 072 rows the 128 x 128 fc1 tiles + a separate fc2 are faster: 84.6 vs 92.1 us)
    const bool fuse_fc2 = tc && sml_use_fused_fc2() && (r.user_tiles + r.item_tiles) <= 12 && !(sml_debug_mask() & (128 | 4096));
    const bool zero_y = tc && (fuse_fc2 || step_ksplit(r, 0) > 1), zero_dA = tc && step_ksplit(r, 1) > 1;
    // loss + dL/dY inside the d2 GEMM (umma_packed.cu, SmlPkLoss): no k_loss launch in the chain (MF step at B = 1024: 84.2 -> 78.2 us; the transfer step stays at 66.7 us, its tail is bound by the two gradient branches)
    const bool fuse_loss = tc && sml_use_fused_loss() && !(sml_debug_mask() & (4 | 8192));
    rc = sml_launch_conv_fwd(g, 3, a->variant, (!tc || need_plain_A) ? w.A : nullptr, tc ? w.Apk : nullptr,
                             want_rowsq ? w.rowsq : nullptr, st, zero_y ? w.Y : nullptr, (zero_dA && fuse_loss) ? w.dA : nullptr);
    if (rc) return rc;
    if (ss) SML_CUDA_OK(cudaStreamWaitEvent(st, ss->join2, 0));
    if (sml_debug_mask() & 256) return SML_OK;
    if (tc) {
        SmlPkProb p[2];
        // fc1: Z1 = A W1^T + b1 (conv_transfer.py:47); also emits GELU(Z1) packed for fc2
        pk_pair(p, r, w.Apk, 10, SML_PK_OFF_P1, 512, a->theta, SML_OFF_F1B, nullptr, w.Z1, 512, w.Gpk, w.theta_pk);
        if (fuse_fc2) {
            // fc1 + fc2 in one launch: every 128 x 64 fc1 tile multiplies its GELU(Z1) tile with its K slice of W2 and adds into Y
            for (int net = 0; net < 2; ++net) {
                p[net].Cpk = nullptr;
                p[net].B2 = w.theta_pk + net * SML_PK_THETA_BYTES + SML_PK_OFF_P2;
                p[net].bias2 = a->theta + (size_t)net * SML_NET_STRIDE + SML_OFF_F2B;
                p[net].Y = w.Y; p[net].ldy = 64;
            }
            rc = sml_launch_umma_packed(p, 2, SML_PK_FC1_FC2, st);
            if (rc) return rc;
        } else {
            rc = sml_launch_umma_packed(p, 2, SML_PK_FC1, st);
            if (rc) return rc;
            if (sml_debug_mask() & 128) return SML_OK;
            // fc2: Y = GELU(Z1) W2^T + b2 (:48-49)
            pk_pair(p, r, w.Gpk, 16, SML_PK_OFF_P2, 64, a->theta, SML_OFF_F2B, nullptr, w.Y, 64, nullptr, w.theta_pk);
            rc = sml_launch_umma_packed(p, 2, SML_PK_FC2, st, step_ksplit(r, 0));
            if (rc) return rc;
        }
    } else {
        SmlGemmProb fc1[2] = {
            {w.A, tu + SML_OFF_F1W, tu + SML_OFF_F1B, nullptr, w.Z1, (int)B, 512, 320, 320, 320, 512},
            {w.A + r.Bp * 320, ti + SML_OFF_F1W, ti + SML_OFF_F1B, nullptr, w.Z1 + r.Bp * 512, (int)(2 * B), 512, 320, 320, 320, 512}};
        rc = sml_launch_sgemm(fc1, 2, SML_A_MK, SML_B_NK, SML_EPI_BIAS, st);
        if (rc) return rc;
        SmlGemmProb fc2[2] = {
            {w.Z1, tu + SML_OFF_F2W, tu + SML_OFF_F2B, nullptr, w.Y, (int)B, 64, 512, 512, 512, 64},
            {w.Z1 + r.Bp * 512, ti + SML_OFF_F2W, ti + SML_OFF_F2B, nullptr, w.Y + r.Bp * 64, (int)(2 * B), 64, 512, 512, 512, 64}};
        rc = sml_launch_sgemm(fc2, 2, SML_A_MK_GELU, SML_B_NK, SML_EPI_BIAS, st);
        if (rc) return rc;
    }
    const int dbg = sml_debug_mask();
    if (dbg & 8) return SML_OK;
    if (!fuse_loss) {
        rc = sml_launch_loss(w.Y, want_rowsq ? w.rowsq : nullptr, B, r.Bp, r.Bp + B, a->loss, a->variant == SML_VARIANT_CONV, l2, w.dY,
                             tc ? w.dYpk : nullptr, scores, a->loss_out, w.partials, w.ticket, st, gbu2, gbi2, zero_dA ? w.dA : nullptr, adaptive);
        if (rc) return rc;
    }
    if (dbg & 4) return SML_OK;
    // dZ1 = (dY W2) * GELU'(Z1)
    if (tc) {
        SmlPkProb p[2];
        // plain dZ1 is only read by the weight gradients (transfer step); the MF step keeps just the packed copy
        pk_pair(p, r, w.dYpk, 2, SML_PK_OFF_P3, 512, a->theta, 0, w.Z1, need_plain_A ? w.dZ1 : nullptr, 512, w.dZpk, w.theta_pk);
        if (g_theta) { p[0].colsum = g_theta + SML_OFF_F1B; p[1].colsum = g_theta + SML_NET_STRIDE + SML_OFF_F1B; }
        if (fuse_loss) {
            // plain dY is only read by the fc2 weight gradient (g_theta != null)
            SmlPkLoss L = {w.Y, want_rowsq ? w.rowsq : nullptr, B, r.Bp, r.Bp + B, a->loss, a->variant == SML_VARIANT_CONV ? 1 : 0, l2, adaptive,
                           g_theta ? w.dY : nullptr, scores, a->loss_out, w.partials, w.ticket, gbu2, gbi2};
            return sml_launch_umma_packed(p, 2, SML_PK_D2, st, 1, &L);
        }
        return sml_launch_umma_packed(p, 2, SML_PK_D2, st);
    }
    SmlGemmProb d2[2] = {
        {w.dY, tu + SML_OFF_F2W, nullptr, w.Z1, w.dZ1, (int)B, 512, 64, 64, 512, 512},
        {w.dY + r.Bp * 64, ti + SML_OFF_F2W, nullptr, w.Z1 + r.Bp * 512, w.dZ1 + r.Bp * 512, (int)(2 * B), 512, 64, 64, 512, 512}};
    return sml_launch_sgemm(d2, 2, SML_A_MK, SML_B_KN, SML_EPI_MUL_GELU_GRAD, st);
}

int fc1_dgrad(const sml_step_args *a, const StepWs &w, cudaStream_t st) {
    const Rows r = rows_of(a->batch);
    const int64_t B = r.B;
    if (sml_use_tensor_cores()) {
        SmlPkProb p[2];
        pk_pair(p, r, w.dZpk, 16, SML_PK_OFF_P4, 320, a->theta, 0, nullptr, w.dA, 320, nullptr, w.theta_pk);
        return sml_launch_umma_packed(p, 2, SML_PK_D1, st, step_ksplit(r, 1));
    }
    const float *tu = a->theta, *ti = a->theta + SML_NET_STRIDE;
    SmlGemmProb d1[2] = {
        {w.dZ1, tu + SML_OFF_F1W, nullptr, nullptr, w.dA, (int)B, 320, 512, 512, 320, 320},
        {w.dZ1 + r.Bp * 512, ti + SML_OFF_F1W, nullptr, nullptr, w.dA + r.Bp * 320, (int)(2 * B), 320, 512, 512, 320, 320}};
    return sml_launch_sgemm(d1, 2, SML_A_MK, SML_B_KN, SML_EPI_NONE, st);
}

// fc1/fc2 weight + bias gradients accumulated into g_theta  (theta grads of conv_transfer.py:47-49)
// the reduction runs over the batch rows (K = B or 2B): slice it so the few output tiles (4 for dW2,
// 20 for dW1 per net) spread over the SMs; slices combine with fp32 atomics
int wgrad_ksplit(int64_t B) { return B >= 2048 ? 16 : (B >= 128 ? 8 : 1); }

int fc2_wgrad(const sml_step_args *a, const StepWs &w, float *g_theta, cudaStream_t st) {
    const Rows r = rows_of(a->batch);
    const int64_t B = r.B, Bp = r.Bp;
    float *gu = g_theta, *gi = g_theta + SML_NET_STRIDE;
    if (sml_use_tensor_cores()) {
        // tensor-core tiles are 128 rows tall: compute dW2^T [512, 64] = g(Z1)^T dY and store it transposed
        SmlGemmProb w2t[2] = {
            {w.Z1, w.dY, nullptr, nullptr, gu + SML_OFF_F2W, 512, 64, (int)B, 512, 64, 512},
            {w.Z1 + Bp * 512, w.dY + Bp * 64, nullptr, nullptr, gi + SML_OFF_F2W, 512, 64, (int)(2 * B), 512, 64, 512}};
        return sml_launch_umma_gemm(w2t, 2, SML_A_KM_GELU, SML_B_KN, SML_EPI_ACCUM, 1, 64, st,
                                    sml_debug_ksplit(2) > 0 ? sml_debug_ksplit(2) : wgrad_ksplit(B));
    }
    SmlGemmProb w2[2] = {
        {w.dY, w.Z1, nullptr, nullptr, gu + SML_OFF_F2W, 64, 512, (int)B, 64, 512, 512},
        {w.dY + Bp * 64, w.Z1 + Bp * 512, nullptr, nullptr, gi + SML_OFF_F2W, 64, 512, (int)(2 * B), 64, 512, 512}};
    return sml_launch_sgemm(w2, 2, SML_A_KM, SML_B_KN_GELU, SML_EPI_ACCUM, st);
}

int fc1_wgrad(const sml_step_args *a, const StepWs &w, float *g_theta, cudaStream_t st) {
    const Rows r = rows_of(a->batch);
    const int64_t B = r.B, Bp = r.Bp;
    float *gu = g_theta, *gi = g_theta + SML_NET_STRIDE;
    const int ksplit = wgrad_ksplit(B);
    int rc;
    SmlGemmProb w1[2] = {
        {w.dZ1, w.A, nullptr, nullptr, gu + SML_OFF_F1W, 512, 320, (int)B, 512, 320, 320},
        {w.dZ1 + Bp * 512, w.A + Bp * 320, nullptr, nullptr, gi + SML_OFF_F1W, 512, 320, (int)(2 * B), 512, 320, 320}};
    if (sml_use_tensor_cores())
        rc = sml_launch_umma_gemm(w1, 2, SML_A_KM, SML_B_KN, SML_EPI_ACCUM, 0, 64, st,
                                  sml_debug_ksplit(3) > 0 ? sml_debug_ksplit(3) : (B >= 2048 ? ksplit / 2 : (ksplit > 2 ? ksplit / 4 : 1)));
    else rc = sml_launch_sgemm(w1, 2, SML_A_KM, SML_B_KN, SML_EPI_ACCUM, st);
    if (rc) return rc;
    if (sml_use_tensor_cores()) return SML_OK;     // bias gradients already accumulated by k_loss and the d2 epilogue
    SmlColsumProb cs[4] = {{w.dY, gu + SML_OFF_F2B, (int)B, 64, 64},
                           {w.dY + Bp * 64, gi + SML_OFF_F2B, (int)(2 * B), 64, 64},
                           {w.dZ1, gu + SML_OFF_F1B, (int)B, 512, 512},
                           {w.dZ1 + Bp * 512, gi + SML_OFF_F1B, (int)(2 * B), 512, 512}};
    return sml_launch_colsum(cs, 4, st);
}

int fc_wgrads(const sml_step_args *a, const StepWs &w, float *g_theta, cudaStream_t st) {
    int rc = fc2_wgrad(a, w, g_theta, st);
    if (rc) return rc;
    return fc1_wgrad(a, w, g_theta, st);
}

int mf_step_impl(const sml_step_args *a, bool pack_theta, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const StepWs w = carve(a->workspace, a->batch);
    int rc = sml_adam_tick(a->adam_state, a->lr, 0.9, 0.999, stream);
    if (rc) return rc;
    const bool lazy = a->stamp_user != nullptr;
    SmlAdamRows ar[3] = {{a->hat_user, a->m_user, a->v_user, a->g_user, a->stamp_user, a->user, a->batch},
                         {a->hat_item, a->m_item, a->v_item, a->g_item, a->stamp_item, a->item, a->batch},
                         {a->hat_item, a->m_item, a->v_item, a->g_item, a->stamp_item, a->neg, a->batch}};
    if (lazy) {   // row-lazy exact Adam: the rows this batch reads catch up with the zero-gradient steps they missed
        rc = sml_launch_adam_rows(ar, 3, a->adam_state, 0, 0.9, 0.999, 1e-8, st);
        if (rc) return rc;
    }
    rc = forward_and_loss(a, w, true, false, pack_theta, (float)a->l2, nullptr, st, nullptr, nullptr, 0.0, (float)a->adaptive_beta);
    if (rc) return rc;
    const int dbg = sml_debug_mask();                  // profiling aid (sml_debug_set_mask): 0 in production
    if (dbg & (4 | 8 | 128 | 256)) return SML_OK;
    if (!(dbg & 2)) {
        rc = fc1_dgrad(a, w, st);
        if (rc) return rc;
        SmlRowGroup g[3];
        make_groups(a, g);
        SmlConvBwdGroup bg[3] = {{g[0], a->g_user, nullptr, -1}, {g[1], a->g_item, nullptr, -1}, {g[2], a->g_item, nullptr, -1}};
        rc = sml_launch_conv_bwd(bg, 3, a->variant, w.dA, (float)a->l2, nullptr, st, (float)a->adaptive_beta);
        if (rc) return rc;
    }
    if (dbg & 64) return SML_OK;
    if (lazy) return sml_launch_adam_rows(ar, 3, a->adam_state, 1, 0.9, 0.999, 1e-8, st);   // this step, on the touched rows
    // dense Adam on both latent tables, weight_decay = 0 (model/transfer.py:392); also re-zeroes the gradients
    rc = sml_adam_dense(a->hat_user, a->m_user, a->v_user, a->g_user, a->n_users * SML_D, a->adam_state, 0.9, 0.999, 1e-8, 0.0, 1, stream);
    if (rc) return rc;
    return sml_adam_dense(a->hat_item, a->m_item, a->v_item, a->g_item, a->n_items * SML_D, a->adam_state, 0.9, 0.999, 1e-8, 0.0, 1, stream);
}

int mf_flush(const sml_step_args *a, void *stream) {
    int rc = sml_adam_flush(a->hat_user, a->m_user, a->v_user, a->stamp_user, a->n_users, a->adam_state, 0.9, 0.999, 1e-8, stream);
    if (rc) return rc;
    return sml_adam_flush(a->hat_item, a->m_item, a->v_item, a->stamp_item, a->n_items, a->adam_state, 0.9, 0.999, 1e-8, stream);
}

int check_mf(const sml_step_args *a) {
    int rc = sml_check_device();
    if (rc) return rc;
    rc = check_args(a, "sml_mf_step");
    if (rc) return rc;
    SML_REQUIRE(a->table_pitch == 0 || a->table_pitch == SML_D, SML_E_BADARG, "sml_mf_step: tables must be dense [rows, 64]");
    SML_REQUIRE(a->g_user && a->g_item && a->m_user && a->v_user && a->m_item && a->v_item && a->adam_state, SML_E_BADARG,
                "sml_mf_step: null gradient / Adam-state pointer");
    SML_REQUIRE((a->stamp_user != nullptr) == (a->stamp_item != nullptr), SML_E_BADARG, "sml_mf_step: give both row stamps or neither");
    return SML_OK;
}

}  // namespace

extern "C" {

size_t sml_step_workspace_bytes(int64_t batch) { return batch > 0 ? step_ws_bytes(batch) : 0; }

int64_t sml_step_rows(int64_t batch, int64_t *row_pos, int64_t *row_neg) {
    const Rows r = rows_of(batch);
    if (row_pos) *row_pos = r.Bp;
    if (row_neg) *row_neg = r.Bp + r.B;
    return r.R;
}

size_t sml_transfer_fwd_workspace_bytes(int64_t n_rows) {
    const int64_t ch = up128(n_rows < FWD_CHUNK ? n_rows : FWD_CHUNK);
    if (ch <= 0) return 0;
    return (size_t)ch * (320 + 512) * sizeof(float) + (size_t)(ch / 128) * (10 + 16) * pk_block_bytes(128) + SML_PK_THETA_BYTES + 4096;
}

int sml_transfer_fwd(const float *x_t, const float *x_hat, const int64_t *ids, int64_t n_rows, int d, int variant,
                     const float *theta_net, int normalize_out, float *out, void *workspace, size_t workspace_bytes,
                     void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_transfer_fwd: d=%d unsupported (d must be %d)", d, SML_D);
    SML_REQUIRE(variant == SML_VARIANT_COM || variant == SML_VARIANT_CONV, SML_E_BADARG, "sml_transfer_fwd: bad variant %d", variant);
    SML_REQUIRE(n_rows >= 0, SML_E_BADARG, "sml_transfer_fwd: negative n_rows");
    if (n_rows == 0) return SML_OK;
    SML_REQUIRE(x_t && x_hat && theta_net && out, SML_E_BADARG, "sml_transfer_fwd: null pointer");
    SML_REQUIRE(workspace && workspace_bytes >= sml_transfer_fwd_workspace_bytes(n_rows), SML_E_WORKSPACE,
                "sml_transfer_fwd: workspace too small (%zu < %zu bytes)", workspace_bytes,
                sml_transfer_fwd_workspace_bytes(n_rows));
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc = sml_use_tensor_cores() != 0;
    const int64_t ch = up128(n_rows < FWD_CHUNK ? n_rows : FWD_CHUNK);
    char *p = (char *)workspace;
    auto take = [&](size_t bytes) { char *q = p; p += sml_align_up(bytes, 256); return q; };
    float *A = (float *)take((size_t)ch * 320 * sizeof(float));
    float *Z1 = (float *)take((size_t)ch * 512 * sizeof(float));
    uint8_t *Apk = (uint8_t *)take((size_t)(ch / 128) * 10 * pk_block_bytes(128));
    uint8_t *Gpk = (uint8_t *)take((size_t)(ch / 128) * 16 * pk_block_bytes(128));
    uint8_t *theta_pk = (uint8_t *)take(SML_PK_THETA_BYTES);
    if (tc && sml_use_fused_fwd() && !(sml_debug_mask() & 2048)) {
        // one persistent kernel over all rows: conv -> fc1 -> GELU -> fc2 (+ normalisation), nothing but the two source rows and
        // the result row crosses HBM (umma_fused_fwd.cu); the packed weights (1.5 MB) reuse the theta_pk scratch
        static_assert(SML_PK_THETA_BYTES >= 1572864, "fused weights must fit the packed-theta scratch");
        rc = sml_launch_pack_fused(theta_net, theta_pk, st);
        if (rc) return rc;
        return sml_launch_transfer_fused(x_t, x_hat, ids, n_rows, SML_D, variant, theta_net, theta_pk, normalize_out, out, st);
    }
    if (tc) {
        rc = sml_launch_pack_theta(theta_net, theta_pk, 1, st);
        if (rc) return rc;
    }
    for (int64_t r0 = 0; r0 < n_rows; r0 += ch) {
        const int64_t n = (n_rows - r0) < ch ? (n_rows - r0) : ch;
        SmlRowGroup g;
        if (ids) g = SmlRowGroup{x_t, x_hat, ids + r0, theta_net, n, 0, SML_D};
        else g = SmlRowGroup{x_t + r0 * SML_D, x_hat + r0 * SML_D, nullptr, theta_net, n, 0, SML_D};
        rc = sml_launch_conv_fwd(&g, 1, variant, tc ? nullptr : A, tc ? Apk : nullptr, nullptr, st);
        if (rc) return rc;
        if (tc) {
            const int tiles = (int)(up128(n) / 128);
            SmlPkProb f1 = {Apk, theta_pk + SML_PK_OFF_P1, 10, tiles, 0, 0, (int)n, 512, theta_net + SML_OFF_F1B, nullptr, nullptr, 512, Gpk, 0, nullptr, nullptr, nullptr, nullptr, 0};
            rc = sml_launch_umma_packed(&f1, 1, SML_PK_FC1, st);
            if (rc) return rc;
            SmlPkProb f2 = {Gpk, theta_pk + SML_PK_OFF_P2, 16, tiles, 0, 0, (int)n, 64, theta_net + SML_OFF_F2B, nullptr, out + r0 * SML_D, 64, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0};
            rc = sml_launch_umma_packed(&f2, 1, SML_PK_FC2, st);
            if (rc) return rc;
        } else {
            SmlGemmProb fc1 = {A, theta_net + SML_OFF_F1W, theta_net + SML_OFF_F1B, nullptr, Z1, (int)n, 512, 320, 320, 320, 512};
            rc = sml_launch_sgemm(&fc1, 1, SML_A_MK, SML_B_NK, SML_EPI_BIAS, st);
            if (rc) return rc;
            SmlGemmProb fc2 = {Z1, theta_net + SML_OFF_F2W, theta_net + SML_OFF_F2B, nullptr, out + r0 * SML_D, (int)n, 64, 512, 512, 512, 64};
            rc = sml_launch_sgemm(&fc2, 1, SML_A_MK_GELU, SML_B_NK, SML_EPI_BIAS, st);
            if (rc) return rc;
        }
    }
    if (normalize_out) return sml_launch_row_normalize(out, n_rows, st);
    return SML_OK;
}

int sml_mf_step(const sml_step_args *a, void *stream) {
    int rc = check_mf(a);
    if (rc) return rc;
    return mf_step_impl(a, true, stream);
}

int sml_tr_step(const sml_step_args *a, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    rc = check_args(a, "sml_tr_step");
    if (rc) return rc;
    SML_REQUIRE(a->g_theta && a->m_theta && a->v_theta && a->adam_state, SML_E_BADARG, "sml_tr_step: null theta-gradient / Adam-state pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const StepWs w = carve(a->workspace, a->batch);
    SideStream *ss = side_stream();
    SML_REQUIRE(ss, SML_E_CUDA, "sml_tr_step: could not create the side stream");
    rc = forward_and_loss(a, w, false, true, true, 0.f, nullptr, st, a->g_theta, a->adam_state, a->lr);   // theta changes every step: re-pack
    if (rc) return rc;
    const int dbg = sml_debug_mask();
    if (dbg & (4 | 8 | 128 | 256)) return SML_OK;
    if (!(dbg & 1)) {
        // branches 1a / 1b: dW1 and dW2 are independent of each other: each on its own side stream (in the CUDA graph: two more
        // parallel branches).  One after the other they ended 8 us after the main branch (tools/tr_timeline.py: dW2 34 -> 48 us,
        // dW1 48 -> 61 us, conv backward done at 53 us) and held up the Adam update.
        SML_CUDA_OK(cudaEventRecord(ss->fork, st));
        SML_CUDA_OK(cudaStreamWaitEvent(ss->s, ss->fork, 0));
        SML_CUDA_OK(cudaStreamWaitEvent(ss->s2, ss->fork, 0));
        rc = fc1_wgrad(a, w, a->g_theta, ss->s);                           // dW1 (the longer one; + bias sums on the SIMT path)
        if (rc) return rc;
        SML_CUDA_OK(cudaEventRecord(ss->join, ss->s));
        rc = fc2_wgrad(a, w, a->g_theta, ss->s2);                          // dW2
        if (rc) return rc;
        SML_CUDA_OK(cudaEventRecord(ss->join3, ss->s2));
    }
    if (!(dbg & 2)) {
        rc = fc1_dgrad(a, w, st);                                          // branch 2: dA, conv backward
        if (rc) return rc;
        SmlRowGroup g[3];
        make_groups(a, g);
        float *gu = a->g_theta, *gi = a->g_theta + SML_NET_STRIDE;
        SmlConvBwdGroup bg[3] = {{g[0], nullptr, gu, -1}, {g[1], nullptr, gi, -1}, {g[2], nullptr, gi, -1}};
        rc = sml_launch_conv_bwd(bg, 3, a->variant, w.dA, 0.f, nullptr, st);
        if (rc) return rc;
    }
    if (!(dbg & 1)) {
        SML_CUDA_OK(cudaStreamWaitEvent(st, ss->join, 0));
        SML_CUDA_OK(cudaStreamWaitEvent(st, ss->join3, 0));
    }
    if (dbg & 64) return SML_OK;
    // Adam with coupled L2 (weight_decay = TR_l2, model/transfer.py:393) over the whole theta block
    if (a->clip_max_norm > 0.0) {
        // --clip_grad (model/transfer.py:723-727): g *= min(1, max_norm / (||g||_2 + 1e-6)) over ALL transfer parameters
        float *sumsq = reinterpret_cast<float *>(w.ticket + 2);
        rc = sml_launch_sumsq(a->g_theta, 2 * (int64_t)SML_NET_STRIDE, sumsq, w.partials, w.ticket + 1, st);
        if (rc) return rc;
        return sml_adam_dense_clipped(a->theta, a->m_theta, a->v_theta, a->g_theta, 2 * (int64_t)SML_NET_STRIDE, a->adam_state, 0.9, 0.999,
                                      1e-8, a->l2, 1, sumsq, a->clip_max_norm, stream);
    }
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
    rc = forward_and_loss(a, w, false, a->g_theta != nullptr, true, 0.f, scores, st, a->g_theta);
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
    // d_rows uses the step row layout (sml_step_rows): user rows at 0, positives at row_pos, negatives at row_neg -- or, with
    // d_rows_by_id, the order of the gathered ids (both item groups index the [2 * batch] rows that start at row_pos)
    const Rows r = rows_of(a->batch);
    const int64_t bu = a->d_rows_by_id ? 0 : -1, bi = a->d_rows_by_id ? r.Bp : -1;
    SmlConvBwdGroup bg[3] = {{g[0], nullptr, gu, bu}, {g[1], nullptr, gi, bi}, {g[2], nullptr, gi, bi}};
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
    int rc = check_mf(a);
    if (rc) return rc;
    sml_step_args s = *a;
    bool first = true;
    int since_flush = 0;
    for (int64_t off = 0; off < n_total; off += a->batch) {
        s.user = a->user + off; s.item = a->item + off; s.neg = a->neg + off;
        s.batch = (n_total - off) < a->batch ? (n_total - off) : a->batch;
        // theta is frozen during the MF epoch (model/transfer.py:428,460): pack its tensor-core operands once.
        // (the workspace carve depends on the batch size, so a short last batch re-packs into its own layout)
        rc = mf_step_impl(&s, first || s.batch != a->batch, stream);
        if (rc) return rc;
        first = false;
        if (a->stamp_user && (++since_flush >= SML_ADAM_HISTORY - 1 || off + a->batch >= n_total)) {
            // row-lazy Adam: every row up to date before the tables are read as a whole (and before the history ring wraps)
            rc = mf_flush(a, stream);
            if (rc) return rc;
            since_flush = 0;
        }
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
