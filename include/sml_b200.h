/*
 * sml_b200.h -- C ABI of libsml_b200.so: hand-written sm_100a CUDA kernels for the
 * per-period retraining hot path of SML (reference: zyang1580/SML, a pure-Python /
 * PyTorch program with no FFI layer of its own -- SURVEY.md section 8b).
 *
 * The reference has no plugin interface; the drop-in boundary is its Python class
 * surface (sml_b200/model/*.py mirrors it).  These entry points are what those
 * classes bind through ctypes, one per stock-PyTorch op sequence they replace; the
 * reference lines each one replaces are cited as (file:line under /root/reference).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the library never allocates or frees persistent memory: callers own all
 *     tensors and pass a scratch workspace (size from sml_*_workspace_bytes);
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), do not
 *     synchronise, and are CUDA-graph capturable;
 *   - return 0 on success, a negative SML_E_* code otherwise; the message is in
 *     sml_last_error() (thread-local);
 *   - latent dimension d must be 64 (the reference default, main_yelp.py:41); other
 *     values return SML_E_UNSUPPORTED;
 *   - there is no CPU path: every entry point fails with SML_E_ARCH on a device
 *     that is not compute capability 10.x.
 */
#ifndef SML_B200_H
#define SML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SML_ABI_VERSION 4

#define SML_OK 0
#define SML_E_BADARG (-1)
#define SML_E_ARCH (-2)
#define SML_E_CUDA (-3)
#define SML_E_UNSUPPORTED (-4)
#define SML_E_WORKSPACE (-5)

#define SML_D 64 /* latent dim */

/* ---- theta layout --------------------------------------------------------------
 * One transfer net (one_transfer, model/conv_transfer.py:18-50) is a flat fp32 block of
 * SML_NET_STRIDE floats; segments are padded so that every segment starts 128-byte
 * aligned (padding stays zero).  A ConvTransfer(_com) module (model/conv_transfer.py:
 * 52-56, 87-91) is [user net | item net] = 2 * SML_NET_STRIDE floats.  The Python
 * nn.Parameters are views into this block with the reference's shapes/state_dict keys.
 *   conv1.weight (10,1,R,1) at +0      (R = 3 for ConvTransfer_com, 2 for ConvTransfer)
 *   conv1.bias   (10)       at +32
 *   conv2.weight (5,10,1,1) at +64
 *   conv2.bias   (5)        at +128
 *   fc1.weight   (512,320)  at +160
 *   fc1.bias     (512)      at +164000
 *   fc2.weight   (64,512)   at +164512
 *   fc2.bias     (64)       at +197280
 */
#define SML_OFF_C1W 0
#define SML_OFF_C1B 32
#define SML_OFF_C2W 64
#define SML_OFF_C2B 128
#define SML_OFF_F1W 160
#define SML_OFF_F1B 164000
#define SML_OFF_F2W 164512
#define SML_OFF_F2B 197280
#define SML_NET_STRIDE 197344
#define SML_FC1_OUT 512
#define SML_FC1_IN 320

/* transfer variants */
#define SML_VARIANT_COM 3  /* ConvTransfer_com: rows [x_t, x_hat, x_t*x_hat/||x_t||]  (conv_transfer.py:92-110) */
#define SML_VARIANT_CONV 2 /* ConvTransfer: rows [x_t, x_hat], user output / ||.||     (conv_transfer.py:57-69)  */
/* loss kinds (ConvTransfer_com.run_MF, conv_transfer.py:122-134) */
#define SML_LOSS_BCE 0
#define SML_LOSS_BPR 1

int sml_abi_version(void);
const char *sml_last_error(void);
/* 0 if the current device is sm_100-class, SML_E_ARCH otherwise. */
int sml_device_check(void);
/* Number of SMs on the current device (grid sizing in the host code). */
int sml_sm_count(void);
/* Kernels this library has launched in this process so far (bench.py's gpu_launches evidence). */
uint64_t sml_launch_count(void);

/* ---- candidate-list evaluation ---------------------------------------------------
 * Replaces MFbasemode.test (model/MF.py:45-80): rows[n, 0] = user id, rows[n, 1..C] =
 * candidate item ids, candidate 0 is the positive.  For every test row writes
 *   gt[n] = #{j>0 : s_j > s_0}, eq[n] = #{j>0 : s_j == s_0}   (NaN sorts highest)
 * with s_j = <user_tab[u], item_tab[c_j]> in fp32.  The position of the positive in
 * torch.topk's output is gt (+eq: the reference's tie order).
 * row_stride = number of int64 per row (>= 1 + C). */
int sml_eval_candidates(const float *user_tab, const float *item_tab, int d, const int64_t *rows, int64_t n_rows,
                        int64_t row_stride, int n_cand, int32_t *gt, int32_t *eq, void *stream);
/* The same counts, bit for bit, from about half the bytes: candidates are first compared through a bf16 copy of the item
 * table (built into item_bf16, sml_eval_prefilter_bytes(n_items) bytes of caller scratch, on every call) with a rigorous
 * error bound; only the undecided few are re-scored in fp32 with the operation order of sml_eval_candidates.
 * Measured on B200 (profiles/r01_kernels_ncu.md): L2 sectors halve, instruction issue becomes the limit, same run time --
 * kept as an option, the fp32 kernel stays the default. */
size_t sml_eval_prefilter_bytes(int64_t n_items);
int sml_eval_candidates_prefilter(const float *user_tab, const float *item_tab, int64_t n_items, int d, void *item_bf16,
                                  const int64_t *rows, int64_t n_rows, int64_t row_stride, int n_cand, int32_t *gt, int32_t *eq,
                                  void *stream);
/* Per batch of `batch` consecutive test rows (evaluation2.test_model batches of 1024,
 * evalution/evaluation2.py:8-26): hits[b] = #{rank < topk}, ndcg[b] = sum 1/log2(rank+2)
 * over hits (fp32, fixed summation order), rank = gt + (tie_loses ? eq : 0). */
int sml_eval_reduce(const int32_t *gt, const int32_t *eq, int64_t n_rows, int batch, int topk, int tie_loses,
                    int32_t *hits, float *ndcg, void *stream);
/* Scores for (user, item) id pairs: MFbasemode.forward (model/MF.py:34-43).
 * score[n] = <user_tab[user[n]], item_tab[item[n]]>; if norm != 0 divided by ||user row||. */
int sml_pair_scores(const float *user_tab, const float *item_tab, int d, const int64_t *user, const int64_t *item,
                    int64_t n, int norm, float *score, void *stream);

/* ---- transfer forward -------------------------------------------------------------
 * Replaces ConvTransfer_com.forward / ConvTransfer.forward (model/conv_transfer.py:57-69,
 * 92-110) and therefore meta_train.updata (model/transfer.py:884-902):
 *   out[n] = one_transfer_theta(stack(x_t[i_n], x_hat[i_n]))   i_n = ids ? ids[n] : n
 * theta_net points at ONE net (user or item).  normalize_out != 0 divides each output row
 * by its L2 norm (ConvTransfer 'user' type).  out may not alias x_t / x_hat. */
size_t sml_transfer_fwd_workspace_bytes(int64_t n_rows);
int sml_transfer_fwd(const float *x_t, const float *x_hat, const int64_t *ids, int64_t n_rows, int d, int variant,
                     const float *theta_net, int normalize_out, float *out, void *workspace, size_t workspace_bytes,
                     void *stream);

/* ---- optimizer scalars ------------------------------------------------------------
 * torch.optim.Adam keeps one step counter per parameter and derives step_size and
 * sqrt(bias_correction2) from it on the host (model/transfer.py:392-393).  To stay
 * graph-capturable the counter lives on the device: sml_adam_tick increments
 * state[0] (int64) and writes step_size = lr/(1-b1^t) and sqrt(1-b2^t) (computed in
 * double) as two floats into the int64 slot state[1].  `state` = 4 int64 slots; if
 * state[2] != 0 the buffer continues with SML_ADAM_HISTORY more slots and the tick also
 * records the float pair of step t at slot 4 + (t mod SML_ADAM_HISTORY), which is what
 * the row-lazy update below replays. */
#define SML_ADAM_HISTORY 4096
int sml_adam_tick(int64_t *state, double lr, double beta1, double beta2, void *stream);
/* Dense Adam over n floats (torch defaults eps=1e-8, amsgrad off; coupled L2 weight_decay):
 *   g' = g + wd*p; m += (g'-m)*(1-b1); v = b2*v + (1-b2)*g'^2; p -= step_size*m/(sqrt(v)/sqrt(bc2)+eps)
 * If zero_grad != 0 the gradient buffer is zeroed after use (replaces zero_grad(),
 * model/transfer.py:464-465,702). */
int sml_adam_dense(float *p, float *m, float *v, float *g, int64_t n, const int64_t *state, double beta1, double beta2,
                   double eps, double weight_decay, int zero_grad, void *stream);
/* The same update with the gradient first scaled by min(1, max_norm / (sqrt(*sumsq) + 1e-6)) (clip_grad_norm_);
 * sml_sumsq writes sum(g^2) over n floats into *sumsq with a fixed summation order (scratch: 1024 floats + 1 uint). */
int sml_sumsq(const float *g, int64_t n, float *sumsq, void *scratch, void *stream);
int sml_adam_dense_clipped(float *p, float *m, float *v, float *g, int64_t n, const int64_t *state, double beta1, double beta2,
                           double eps, double weight_decay, int zero_grad, const float *sumsq, double max_norm, void *stream);

/* Row-lazy, bit-identical form of the dense update for embedding tables ([n_rows, 64], weight_decay 0).
 * The reference's MF optimizer is DENSE Adam over whole nn.Embedding tables (model/MF.py:21-24,
 * model/transfer.py:392): a row without gradient still moves through its momentum tail every step.
 * Instead of sweeping the tables every step, each row carries the number of the last step applied to
 * it (stamp[row], int32) and the zero-gradient steps it missed are replayed in registers -- same
 * operations in the same order as sml_adam_dense, so the tables are bit-identical to the dense sweep --
 *   sml_adam_rows(apply = 0): bring the listed rows up to step t-1, before a step reads them;
 *   sml_adam_rows(apply = 1): apply step t with the accumulated gradient rows g[id], re-zero them;
 *   sml_adam_flush:           bring EVERY row up to step t (before the table is read as a whole).
 * t = state[0]; state must carry the history (state[2] != 0) and no row may lag more than
 * SML_ADAM_HISTORY - 1 steps.  Duplicate ids are fine (first claimant updates the row).
 * stamp[row] = SML_STAMP_IDLE marks a row whose exp_avg / exp_avg_sq have been +0 since the table was created (no gradient
 * ever reached it): zero-gradient steps leave such a row unchanged bit for bit, so nothing is replayed and sml_adam_flush skips
 * it on the stamp alone.  Fresh stamps for all-zero moments should be SML_STAMP_IDLE; sml_adam_flush also sets it for rows it
 * finds with all-zero moments. */
#define SML_STAMP_IDLE 0x7fffffff
int sml_adam_rows(float *p, float *m, float *v, float *g, int32_t *stamp, const int64_t *ids, int64_t n_ids, int64_t n_rows,
                  const int64_t *state, int apply, double beta1, double beta2, double eps, void *stream);
int sml_adam_flush(float *p, float *m, float *v, int32_t *stamp, int64_t n_rows, const int64_t *state, double beta1,
                   double beta2, double eps, void *stream);

/* ---- SML steps ----------------------------------------------------------------------
 * Shared argument block for the two hot loops.  Rows [0,B) of every workspace matrix are
 * the user rows (user net), [B,2B) the positive-item rows and [2B,3B) the negative-item
 * rows (item net). */
typedef struct {
    /* batch */
    const int64_t *user, *item, *neg; /* [B] ids */
    int64_t batch;
    /* tables, [n_users|n_items, 64] fp32 row-major */
    const float *last_user, *last_item; /* w_{t-1}: x_t source (never written)                     */
    float *hat_user, *hat_item;         /* x_hat source. MF step: the live MFbase latent tables      */
                                        /* (updated in place); TR step: the w_hat snapshots (read)   */
    int64_t n_users, n_items;
    /* transfer parameters: [user net | item net] */
    float *theta;
    int variant; /* SML_VARIANT_* */
    int loss;    /* SML_LOSS_*    */
    /* MF step only: dense gradient scratch (all-zero on entry and on exit) + Adam state */
    float *g_user, *g_item, *m_user, *v_user, *m_item, *v_item;
    int64_t *adam_state; /* 4 x int64, see sml_adam_tick */
    double lr, l2;       /* MF: MF_lr, l2 (model/transfer.py:486-488); TR: TR_lr, TR_l2 (weight decay) */
    /* TR step only: theta gradient + Adam state, each 2*SML_NET_STRIDE floats */
    float *g_theta, *m_theta, *v_theta;
    /* outputs */
    float *loss_out; /* [2]: loss_out[0] = this step's loss (as the reference's loss_batch),      */
                     /*      loss_out[1] += loss (the reference's loss_all accumulation, :501,722) */
    /* scratch (zero-initialised by the caller before its first use) */
    void *workspace;
    size_t workspace_bytes;
    /* floats between consecutive rows of last_* / hat_* (0 = 64).  sml_run_mf_grads only: lets the four
     * "tables" be views into exchanged [last | hat] row pairs (pitch 128) without a de-interleave copy. */
    int64_t table_pitch;
    /* MF step only, optional: per-row "last Adam step applied" stamps ([n_users] / [n_items] int32).  Non-null
     * selects the row-lazy exact Adam (sml_adam_rows) instead of the dense sweeps: sml_mf_epoch flushes both
     * tables before it returns, after sml_mf_step the caller must (sml_adam_flush) before reading the tables. */
    int32_t *stamp_user, *stamp_item;
    /* Options that are off by default in the reference (0 = off).
     * adaptive_beta (MF step, --need_adaptive, model/transfer.py:490-499, beta = 0.1 there): adds
     *     sum over the distinct users u of the batch of beta * count_u / ||w_u||.detach() * ||w_u||^2
     *   to the loss, i.e. beta * ||w_u|| per occurrence and 2 * beta * w_u / ||w_u|| to the row gradient per occurrence.
     * clip_max_norm (transfer step, --clip_grad / --maxnorm_grad, model/transfer.py:723-727): the theta gradient is scaled
     *   by min(1, max_norm / (||g||_2 + 1e-6)) before the Adam update (torch.nn.utils.clip_grad_norm_).  The same call in
     *   the MF step (:507-510) clips gradients that no optimizer ever applies: nothing to do there. */
    double adaptive_beta, clip_max_norm;
    /* sml_run_mf_grads only.  0: d_rows uses the step row layout (sml_step_rows).  1: the gradient of the row gathered
     * through id k is written to d_rows[k] (user rows) / d_rows[*row_pos + k] (positive and negative item rows, which
     * index one [2 * batch]-row table): when the "tables" are row pairs received from an exchange and the ids are the
     * exchange's inverse permutation, d_rows comes out in the order the gradients are sent back (sml_b200/shard.py). */
    int32_t d_rows_by_id;
} sml_step_args;

size_t sml_step_workspace_bytes(int64_t batch);
/* Row layout of the per-step matrices (and of d_rows in sml_run_mf_grads): user rows at [0, batch),
 * positive-item rows at [*row_pos, +batch), negative-item rows at [*row_neg, +batch); returns the total
 * (128-padded) row count. */
int64_t sml_step_rows(int64_t batch, int64_t *row_pos_host, int64_t *row_neg_host);
/* HOT LOOP A body, model/transfer.py:463-511: gather -> transfer fwd (3 nets calls) -> BCE
 * -> + l2*0.5*sum(w_hat^2) -> row gradients through the x_hat channel -> scatter-add ->
 * dense Adam on both latent tables.  theta is read-only here. */
int sml_mf_step(const sml_step_args *args, void *stream);
/* HOT LOOP B body, model/transfer.py:701-728: same forward on snapshot rows, theta
 * gradients only, Adam with coupled L2 on theta. */
int sml_tr_step(const sml_step_args *args, void *stream);
/* One whole epoch of either loop: args->user/item/neg hold n_total triples in batch order and
 * args->batch is the nominal batch size; steps ceil(n_total / batch) times (the last batch may be
 * short, like DataLoader(drop_last=False), model/transfer.py:439-443,692-696) without returning to
 * the host language between steps.  loss_out[1] accumulates the per-step losses. */
int sml_mf_epoch(const sml_step_args *args, int64_t n_total, void *stream);
int sml_tr_epoch(const sml_step_args *args, int64_t n_total, void *stream);
/* Forward + loss + gradients without any optimizer update (ConvTransfer_com.run_MF +
 * backward, model/conv_transfer.py:113-135): writes d_rows [sml_step_rows(B), 64] = dL/d x_hat rows (no l2
 * term; row layout of sml_step_rows) if non-null and accumulates theta gradients into args->g_theta if
 * non-null. */
int sml_run_mf_grads(const sml_step_args *args, float *d_rows, float *scores /* [2B] s+, s- or null */, void *stream);

/* ---- full-catalog evaluation (north_star item 3 / config 5) ----------------------------------
 * rank of every evaluated (user, positive item) pair among ALL items: a tcgen05 3xTF32 score GEMM
 * [users x 64] x [64 x items] with a fused compare-and-count epilogue.
 *   sml_pack_rows   : rows of a [*, 64] table (gathered by ids, or rows 0..n-1 when ids is null) -> packed
 *                     tensor-core operand (sml_packed_rows_bytes(n) bytes)
 *   sml_fullcat_rank: gt[u] += #{i : s_ui > s_pos[u]}, eq[u] += #{i : s_ui == s_pos[u]} over the n_items items of
 *                     items_packed whose global ids start at item_id0; the item with id pos_id[u] is skipped.
 *                     gt / eq must be zeroed by the caller; item shards on several GPUs add up (all-reduce).
 * s_pos[u] = <user_u, item_pos_u> comes from sml_pair_scores (fp32 FFMA). */
size_t sml_packed_rows_bytes(int64_t n_rows);
int sml_pack_rows(const float *tab, const int64_t *ids, int64_t n_rows, int d, void *out, void *stream);
int sml_fullcat_rank(const void *users_packed, const void *items_packed, const float *s_pos, const int64_t *pos_id, int64_t n_users,
                     int64_t n_items, int64_t item_id0, int32_t *gt, int32_t *eq, void *stream);

/* The positive's score from the SAME tensor-core arithmetic as the catalog scores (so that the counts of sml_fullcat_rank are
 * self-consistent: a catalog item equal to the positive ties exactly): s_pos[u] = score of user row u of users_packed with row
 * u of pos_packed (= sml_pack_rows of the item table gathered by the positives' ids). */
int sml_fullcat_pos_scores(const void *users_packed, const void *pos_packed, int64_t n_users, float *s_pos, void *stream);
/* Full-catalog top-k (north_star item 3): for every user row of users_packed the k (<= 64) highest-scoring items of
 * items_packed (n_items items, global ids from item_id0), scores descending -- out_scores / out_ids [n_users, k]; the item
 * exclude_id[u] is skipped when exclude_id is non-null; NaN scores rank highest (torch.topk); entries beyond the catalog size
 * are (-inf, -1).  Fused into the score GEMM's epilogue: each epilogue thread keeps the k best of the columns it sees (a
 * register threshold, the list in scratch; ~k ln(n_items / k) insertions per row), one merge launch combines the lists.
 * Catalog shards on several GPUs: run per shard, concatenate, select again. */
size_t sml_fullcat_topk_workspace_bytes(int64_t n_users, int64_t n_items, int k);
int sml_fullcat_topk(const void *users_packed, const void *items_packed, const int64_t *exclude_id, int64_t n_users, int64_t n_items,
                     int64_t item_id0, int k, float *out_scores, int64_t *out_ids, void *workspace, size_t workspace_bytes, void *stream);

/* ---- row exchange for row-sharded tables (north_star item 4) ---------------------------------
 * Tables are sharded by id (owner = id % world, local row = id / world).  Owner side of the exchange:
 *   sml_gather_pairs : out[n] = [last[loc[n]] | hat[loc[n]]]   (2*d floats per id; answers an id request)
 *   sml_scatter_grads: g[loc[n]] += scale * d_rows[n] + l2 * hat[loc[n]]   (row gradients that came back;
 *                      the l2 term of model/transfer.py:486 is added per occurrence by the owner)
 * The NCCL all-to-all between them is issued by the host side (sml_b200/shard.py). */
int sml_gather_pairs(const float *last, const float *hat, const int64_t *loc, int64_t n, int d, float *out, void *stream);
int sml_scatter_grads(float *g, const float *hat, const int64_t *loc, const float *d_rows, int64_t n, int d, double scale,
                      double l2, void *stream);

/* ---- GPU negative sampler (throughput runs; parity runs use the host emulation) ---------------
 * neg[s] = uniform member of item_all[n_items_all] that is not an interaction of users[s] in this period
 * (keys = sorted user*span+item), Philox4x32-10, subsequence = s, so results do not depend on the launch shape.
 * Semantics of offlineDataset_withsample.__getitem__, data/dataset.py:62-71. */
int sml_philox_negatives(const int64_t *users, int64_t n, const int64_t *item_all, int64_t n_items_all, const int64_t *keys,
                         int64_t n_keys, int64_t span, uint64_t seed, uint64_t offset, int64_t *neg, void *stream);

/* ---- host helper (all pointers are HOST pointers) ------------------------------------------
 * The sequential walk of offlineDataset_withsample's rejection sampler (data/dataset.py:62-71): sample s
 * takes item_all[draws[p]] for successive p until (users[s], item) is not an interaction of the period
 * (keys = sorted user*span+item).  Returns the number of draws consumed, or -1 if n_draws was too few. */
int64_t sml_host_rejection_walk(const int64_t *draws_host, int64_t n_draws, const int64_t *users_host, int64_t n,
                                const int64_t *item_all_host, const int64_t *keys_host, int64_t n_keys, int64_t span,
                                int64_t *neg_host);

/* The same walk with the membership test through an open-addressing hash set (one or two probes instead of a
 * 17-level binary search) and resumable, so the caller draws exactly as many numbers as the reference consumes:
 *   sml_host_keyset_build: table[table_size] (power of two >= 2 n) <- the period's user*span+item keys;
 *   sml_host_rejection_walk_hashed: continues at sample *sample_io, stops when samples or draws run out, returns the
 *   draws consumed (called with as many draws as samples remain, it consumes all of them). */
int sml_host_keyset_build(const int64_t *users_host, const int64_t *items_host, int64_t n, int64_t span, int64_t *table_host,
                          int64_t table_size);
int64_t sml_host_rejection_walk_hashed(const int64_t *draws_host, int64_t n_draws, const int64_t *users_host, int64_t n,
                                       int64_t *sample_io, const int64_t *item_all_host, const int64_t *table_host,
                                       int64_t table_size, int64_t span, int64_t *neg_host);

/* ---- GEMM building block (exposed for tests and profiling) ---------------------------------
 * C[M,N] = epi(opA(A) opB(B)) with the fc-layer GEMM kernels the steps use.  a_mode: 0 A[m][k], 1 same
 * with GELU applied on load, 2 A[k][m], 3 A[k][m] + GELU.  b_mode: 0 B[n][k], 1 B[k][n], 2 B[k][n] + GELU.
 * epi: 0 store, 1 + bias[n], 2 * GELU'(aux[m][n]), 3 accumulate into C.  tensor_cores != 0 selects the
 * tcgen05 3xTF32 kernel (bn = 64 | 128, optional transposed store C[n][m]), 0 the SIMT fp32 kernel. */
/* Profiling aid (tools/tr_breakdown.py): bit mask that drops stages of sml_tr_step / sml_mf_step so that the critical
 * path can be measured by difference.  1: no weight gradients, 2: no dA / conv backward, 4: nothing after the loss,
 * 8: nothing after fc2, 64: no optimizer update, 128: nothing after fc1, 256: nothing after the conv prologue, 512: no 128 x 64 tiles for small batches, 2048: three-kernel transfer forward, 4096: separate fc2 launch, 8192: separate loss kernel.
 * Results are garbage unless the mask is 0 (the default) or 512.  Returns the previous mask. */
int sml_debug_set_mask(int mask);
int sml_debug_mask(void);
/* Tuning aid: override the split-K factors of the step GEMMs (fc2, d1, dW2, dW1); 0 = built-in choice. */
int sml_debug_set_ksplit(int fc2, int d1, int w2, int w1);
int sml_debug_ksplit(int which);
int sml_debug_gemm(const float *A, const float *B, const float *bias, const float *aux, float *C, int M, int N, int K, int lda,
                   int ldb, int ldc, int a_mode, int b_mode, int epi, int transpose_out, int bn, int tensor_cores, void *stream);

/* ---- plain MF steps (baselines / MF2; north_star kernel 1) ---------------------------
 * model/baseline.py:188-201 (BCE, mean, separate l2_u / l2_i) and MF2.forward
 * (model/MF.py:129-147, BPR sum with item biases): fused gather - dot - loss - scatter-add
 * into the dense gradient buffers, then sml_adam_dense by the caller.  bias pointers may be
 * null (BCE path ignores them). */
int sml_plain_mf_grads(const float *user_tab, const float *item_tab, const float *item_bias, const int64_t *user,
                       const int64_t *item, const int64_t *neg, int64_t batch, int d, int loss, double l2_u,
                       double l2_i, float *g_user, float *g_item, float *g_item_bias, float *loss_out,
                       void *workspace, size_t workspace_bytes, void *stream);

/* The whole plain-MF step in ONE kernel (north_star item 1): 128-bit row gather, half-warp dot products, BCE-mean
 * (model/baseline.py:188-201) or BPR-sum (model/MF.py:141-144 without the bias tables) loss, row gradients with the L2
 * term, and Adam on exactly the rows of the batch -- no table-sized gradient buffer, no second launch.
 *   optimizer = SML_OPT_ADAM_DENSE_EXACT: the reference's DENSE torch.optim.Adam (model/baseline.py:111) in its row-lazy
 *       bit-identical form (see sml_adam_rows): adam_state must carry the history ring, stamp_* the per-row stamps;
 *       call sml_adam_flush before reading a table as a whole;
 *   optimizer = SML_OPT_ADAM_SPARSE: only the rows of the batch move, moments decay only when touched (lazy Adam of
 *       large embedding tables; NOT the reference's semantic -- scaled throughput runs); stamp_* may be null.
 * head_user / head_item: int32 [n_users] / [n_items] list heads, all -1 on entry and on exit (duplicates of a row inside
 * the batch are chained through them).  The step counter in adam_state is advanced by the kernel (no sml_adam_tick).
 * loss_out[0] = this step's loss, loss_out[1] += loss.  workspace: sml_plain_mf_step_workspace_bytes(batch) bytes. */
#define SML_OPT_ADAM_DENSE_EXACT 0
#define SML_OPT_ADAM_SPARSE 1
size_t sml_plain_mf_step_workspace_bytes(int64_t batch);
int sml_plain_mf_step(float *user_tab, float *item_tab, float *m_user, float *v_user, float *m_item, float *v_item,
                      int32_t *stamp_user, int32_t *stamp_item, int32_t *head_user, int32_t *head_item, const int64_t *user,
                      const int64_t *item, const int64_t *neg, int64_t batch, int d, int loss, double l2_u, double l2_i,
                      int64_t *adam_state, double lr, int optimizer, float *loss_out, void *workspace, size_t workspace_bytes,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SML_B200_H */
