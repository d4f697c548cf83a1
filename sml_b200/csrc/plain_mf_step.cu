// Plain matrix-factorisation step, fully fused (BASELINE.json north_star item 1): ONE cooperative kernel does
//     128-bit row gather -> half-warp dot products -> BCE-mean / BPR-sum loss -> row gradients (+ the L2 term)
//     -> Adam on exactly the rows of the batch
// for a batch of (user, positive item, negative item) triples -- the step of the reference's MF baselines
// (model/baseline.py:188-201: BCE with separate l2_u / l2_i, torch.optim.Adam over nn.Embedding tables, :111) and of
// MF2.forward's BPR objective (model/MF.py:129-147, without the bias tables: sml_plain_mf_grads keeps those).
//
// HBM-bound.  Algorithmic bytes per triple: read p, m, v and write p, m, v of 3 rows = 4 608 B + 24 B ids.
// Three phases separated by grid-wide syncs (cooperative launch, every CTA resident):
//   0  ids only: every occurrence of a row in the batch is chained to the previous one through a per-row list head (one
//      int32 per table row, an atomicExch per occurrence) -- afterwards every occurrence knows whether it is alone;
//   1  one half-warp per triple: 128-bit row gather (16 lanes x 16 B per row; cp.async into per-lane shared-memory slots, one
//      triple ahead of the arithmetic), dot products, loss, the three row gradients.  A row that occurs ONCE in the batch (practically all of them on large tables) is finished on the
//      spot: p, m, v are read once, Adam is applied in registers, p, m, v are written once -- exactly the algorithmic
//      traffic.  Rows with several occurrences park their gradient rows in a compact per-occurrence buffer;
//   2  the first occurrence of such a row sums its chain and applies the update once -- embedding_dense_backward's sum
//      followed by one optimizer step, without a table-sized gradient buffer and without atomics on floats.
//
// Optimizer modes
//   SML_OPT_ADAM_DENSE_EXACT  the reference's semantic: DENSE torch.optim.Adam, in its row-lazy bit-identical form
//                             (adam.cu): a row replays the zero-gradient steps it missed since its stamp before it
//                             is read (phase A, in registers) and before it is updated (phase B).
//   SML_OPT_ADAM_SPARSE       only the rows of the batch move and their moments only decay when touched (the usual
//                             "lazy Adam" of large embedding tables).  NOT the reference's semantic: for scaled
//                             throughput runs where a dense optimizer sweep is out of the question.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "sml_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int PMS_THREADS = 256;
constexpr int PMS_HALFWARPS = PMS_THREADS / 16;
constexpr int PMS_MAX_BLOCKS = 2048;

struct PmsParams {
    float *user_tab, *item_tab, *m_user, *v_user, *m_item, *v_item;
    int32_t *stamp_user, *stamp_item, *head_user, *head_item;
    const int64_t *user, *item, *neg;
    int64_t B;
    int loss_kind, optimizer, phases;      // phases: tuning aid (SML_PMS_PHASES bit mask, 7 = all; results are garbage otherwise)
    float l2_u, l2_i;
    int64_t *state;
    double lr;
    float b1c, beta2, b2c, eps;
    float *loss_out;
    float4 *gradbuf;      // [3B][16] float4: gradient row of occurrence o = 3 b + k (k = 0 user, 1 positive, 2 negative)
    int32_t *next;        // [3B] previous occurrence of the same row in this batch, -1 = first (the owner)
    float *partials;      // [2 * gridDim.x]
    int32_t *worklist;    // [3B] first occurrences of rows that occur several times (phase 2 work)
    unsigned int *work_count;   // [0..1] entries in worklist, double-buffered by launch parity (zero between launches); [2] launches so far
                                // (the parity lives in the workspace, not in the Adam step: a workspace outlives optimizer states)
};

__device__ __forceinline__ float hw_sum(float v) {      // sum over the 16 lanes of a half-warp
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// (L2 loads: slot t of the ring is written by this very kernel, and the line it shares with the step counter sits in every
// SM's L1 from the first instruction on)
__device__ __forceinline__ float2 hist_at(const int64_t *state, int s) {
    return __ldcg(reinterpret_cast<const float2 *>(state + 4 + (s & (SML_ADAM_HISTORY - 1))));
}
__device__ __forceinline__ void adam4(float4 &p, float4 &m, float4 &v, const float4 g, const PmsParams &P, float2 c) {
    sml_adam1(p.x, m.x, v.x, g.x, P.b1c, P.beta2, P.b2c, c.x, c.y, P.eps, 0.f);
    sml_adam1(p.y, m.y, v.y, g.y, P.b1c, P.beta2, P.b2c, c.x, c.y, P.eps, 0.f);
    sml_adam1(p.z, m.z, v.z, g.z, P.b1c, P.beta2, P.b2c, c.x, c.y, P.eps, 0.f);
    sml_adam1(p.w, m.w, v.w, g.w, P.b1c, P.beta2, P.b2c, c.x, c.y, P.eps, 0.f);
}
__device__ __forceinline__ bool idle4(const float4 &m, const float4 &v) {
    return (__float_as_uint(m.x) | __float_as_uint(m.y) | __float_as_uint(m.z) | __float_as_uint(m.w) | __float_as_uint(v.x) |
            __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w)) == 0u;
}
// One table row in registers.  p is always loaded; m and v only when the row is going to be updated by this half-warp
// (solo) or has zero-gradient steps to replay first (row-lazy exact Adam).  All loads are plain L2-cached loads, not the
// non-coherent path: later phases of the same kernel write these tables.
struct RowRegs { float4 p, m, v; };
// p, m, v of a row are requested together (the m / v loads of a row that turns out to have several occurrences in the batch
// are wasted L2 traffic, but waiting for that answer first would put one more DRAM round trip on the critical path)
__device__ __forceinline__ RowRegs load_row(const float *tab, const float *mt, const float *vt, int64_t id, int l16) {
    RowRegs r;
    const size_t e = (size_t)id * SML_D;
    r.p = __ldcg(reinterpret_cast<const float4 *>(tab + e) + l16);
    r.m = __ldcg(reinterpret_cast<const float4 *>(mt + e) + l16);
    r.v = __ldcg(reinterpret_cast<const float4 *>(vt + e) + l16);
    return r;
}
// row-lazy exact Adam: replay the zero-gradient steps (stamp, t_prev] in registers
__device__ __forceinline__ void catch_up(RowRegs &r, int st, int t_prev, const PmsParams &P) {
    if (st < t_prev && !idle4(r.m, r.v)) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = st + 1; s <= t_prev; ++s) adam4(r.p, r.m, r.v, z, P, hist_at(P.state, s));
    }
}
__device__ __forceinline__ void store_row(float *tab, float *mt, float *vt, int64_t id, int l16, const RowRegs &r) {
    const size_t e = (size_t)id * SML_D;
    reinterpret_cast<float4 *>(tab + e)[l16] = r.p;
    reinterpret_cast<float4 *>(mt + e)[l16] = r.m;
    reinterpret_cast<float4 *>(vt + e)[l16] = r.v;
}

template <int MINB>      // resident CTAs per SM the register allocation aims at (2: no spills; 3: 80 registers, a few spills)
__global__ void __launch_bounds__(PMS_THREADS, MINB)
k_plain_mf_step(PmsParams P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float s_part[PMS_THREADS / 32][2];
    const int lane = threadIdx.x & 31, l16 = lane & 15, w = threadIdx.x >> 5;
    const int64_t tid = (int64_t)blockIdx.x * PMS_THREADS + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * PMS_THREADS;
    const int64_t hw0 = tid >> 4, nhw = nthreads >> 4;
    const int t_prev = (int)__ldcg(P.state);
    const int t = t_prev + 1;
    const int64_t n_occ = 3 * P.B;
    if (tid == 0) {
        // the scalars of step t (torch.optim.Adam computes them on the host in double): into the history ring now, nobody
        // reads slot t before the first grid sync; the step counter itself moves after the last one
        const float ss = (float)(P.lr / (1.0 - pow(0.9, (double)t)));
        const float bs = (float)sqrt(1.0 - pow(0.999, (double)t));
        float *h = reinterpret_cast<float *>(P.state + 4 + (t & (SML_ADAM_HISTORY - 1)));
        h[0] = ss; h[1] = bs;
    }
    // ---------------- phase 0: chain the occurrences of every row (ids only) ----------------
    // occurrence o = 3 b + k (k = 0 user, 1 positive, 2 negative); next[o] = the previous occurrence of the same row, -1 = first
    for (int64_t o = tid; o < n_occ && (P.phases & 1); o += nthreads) {
        const uint32_t b = (uint32_t)o / 3u, k = (uint32_t)o - 3u * b;
        const int64_t id = k == 0 ? __ldg(P.user + b) : (k == 1 ? __ldg(P.item + b) : __ldg(P.neg + b));
        P.next[o] = atomicExch((k == 0 ? P.head_user : P.head_item) + id, (int32_t)o);
    }
    grid.sync();
    // ---------------- phase 1: gather, scores, loss; rows that occur ONCE in the batch are updated right here ----------------
    // Software pipeline per half-warp, one triple per iteration:  ids two triples ahead (registers)  ->  the nine 256 B row
    // chunks (p, m, v of the three rows) of the NEXT triple by cp.async into this lane's own shared-memory slots, together with
    // its chain state (next / head / stamp, lanes 0..2)  ->  compute + update of the CURRENT triple from shared memory.
    // Every lane only ever reads back the 16 bytes it copied itself, so the staging needs no synchronisation beyond the lane's
    // own cp.async.wait_group; it is "registers filled asynchronously": twice the bytes in flight at half the register count.
    extern __shared__ float4 stage[];                       // [2 buffers][9 chunks][PMS_THREADS lanes]
    const float2 ct = hist_at(P.state, t);
    const float invB = 1.0f / (float)P.B;
    float acc_loss = 0.f, acc_l2 = 0.f;
    const int64_t B_up = (P.B + 1) & ~(int64_t)1;          // both half-warps of a warp stay in the loop for the shuffles
    const unsigned int seq = __ldcg(P.work_count + 2);     // every CTA reads it before the second grid sync; it moves after that one
    unsigned int *work_count = P.work_count + (seq & 1u);
    auto slot = [&](int buf, int j) -> float4 * { return stage + (size_t)(buf * 9 + j) * PMS_THREADS + threadIdx.x; };
    auto ids_of = [&](int64_t b, int64_t &iu, int64_t &ii, int64_t &ij) {
        const int64_t bb = b < P.B ? b : P.B - 1;          // triples past the end recompute the last one (never stored)
        iu = __ldg(P.user + bb); ii = __ldg(P.item + bb); ij = __ldg(P.neg + bb);
    };
    auto issue = [&](int buf, int64_t iu, int64_t ii, int64_t ij) {
        const float *src[9] = {P.user_tab + iu * SML_D, P.m_user + iu * SML_D, P.v_user + iu * SML_D,
                               P.item_tab + ii * SML_D, P.m_item + ii * SML_D, P.v_item + ii * SML_D,
                               P.item_tab + ij * SML_D, P.m_item + ij * SML_D, P.v_item + ij * SML_D};
#pragma unroll
        for (int j = 0; j < 9; ++j)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(slot(buf, j))),
                         "l"(reinterpret_cast<const float4 *>(src[j]) + l16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // chain state of occurrence 3 b + l16 (lanes 0..2): x = next, y = head of its row, z = stamp of its row
    auto meta_of = [&](int64_t b, int64_t iu, int64_t ii, int64_t ij) -> int3 {
        int3 m = make_int3(0, 0, t_prev);
        if (l16 < 3) {
            const int64_t bb = b < P.B ? b : P.B - 1;
            const int64_t id = l16 == 0 ? iu : (l16 == 1 ? ii : ij);
            m.x = __ldcg(P.next + 3 * bb + l16);
            m.y = __ldcg((l16 == 0 ? P.head_user : P.head_item) + id);
            if (P.optimizer == SML_OPT_ADAM_DENSE_EXACT) m.z = __ldcg((l16 == 0 ? P.stamp_user : P.stamp_item) + id);
        }
        return m;
    };
    if (P.phases & 2) {
        int64_t u0, i0, j0, u1, i1, j1, u2, i2, j2;
        ids_of(hw0, u0, i0, j0);
        ids_of(hw0 + nhw, u1, i1, j1);
        issue(0, u0, i0, j0);
        int3 meta0 = meta_of(hw0, u0, i0, j0);
        int it = 0;
        for (int64_t b = hw0; b < B_up; b += nhw, ++it) {
            const int buf = it & 1;
            ids_of(b + 2 * nhw, u2, i2, j2);
            issue(buf ^ 1, u1, i1, j1);                      // (a harmless re-read of the last triple when b + nhw is past the end)
            const int3 meta1 = meta_of(b + nhw, u1, i1, j1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            const bool live = b < P.B;
            const int32_t o = (int32_t)(3 * (live ? b : P.B - 1)) + l16;
            const int solo_l = live && l16 < 3 && meta0.x == -1 && meta0.y == o;        // my occurrence is the only one of its row
            const bool solo_u = __shfl_sync(0xffffffffu, solo_l, 0, 16), solo_i = __shfl_sync(0xffffffffu, solo_l, 1, 16),
                       solo_j = __shfl_sync(0xffffffffu, solo_l, 2, 16);
            RowRegs ru, ri, rj;
            ru.p = *slot(buf, 0); ri.p = *slot(buf, 3); rj.p = *slot(buf, 6);
            if (P.optimizer == SML_OPT_ADAM_DENSE_EXACT) {
                const int su = __shfl_sync(0xffffffffu, meta0.z, 0, 16), si = __shfl_sync(0xffffffffu, meta0.z, 1, 16),
                          sj = __shfl_sync(0xffffffffu, meta0.z, 2, 16);
                ru.m = *slot(buf, 1); ru.v = *slot(buf, 2); catch_up(ru, su, t_prev, P);
                ri.m = *slot(buf, 4); ri.v = *slot(buf, 5); catch_up(ri, si, t_prev, P);
                rj.m = *slot(buf, 7); rj.v = *slot(buf, 8); catch_up(rj, sj, t_prev, P);
            }
            const float4 u = ru.p, vi = ri.p, vj = rj.p;
            const float sp = hw_sum(fmaf(u.w, vi.w, fmaf(u.z, vi.z, fmaf(u.y, vi.y, u.x * vi.x))));
            const float sn = hw_sum(fmaf(u.w, vj.w, fmaf(u.z, vj.z, fmaf(u.y, vj.y, u.x * vj.x))));
            float dsp, dsn;
            if (P.loss_kind == SML_LOSS_BCE) {                 // model/baseline.py:197-199
                const float gp = sml_sigmoid(sp), gn = sml_sigmoid(sn);
                const float ap = gp + 1e-15f, an = (1.0f - gn) + 1e-15f;
                dsp = -(gp * (1.0f - gp)) / ap * invB;
                dsn = (gn * (1.0f - gn)) / an * invB;
                const float qu = hw_sum(u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w);
                const float qi = hw_sum(vi.x * vi.x + vi.y * vi.y + vi.z * vi.z + vi.w * vi.w + vj.x * vj.x + vj.y * vj.y + vj.z * vj.z + vj.w * vj.w);
                if (live && l16 == 0) { acc_loss += logf(ap) + logf(an); acc_l2 += P.l2_u * 0.5f * qu + P.l2_i * 0.5f * qi; }
            } else {                                           // model/MF.py:141-144 without the bias tables
                const float x = sp - sn;
                if (live && l16 == 0) acc_loss += fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
                dsp = -sml_sigmoid(-x);
                dsn = -dsp;
            }
            if (live) {
                float4 gu, gi, gj;
                gu.x = fmaf(P.l2_u, u.x, dsp * vi.x + dsn * vj.x); gu.y = fmaf(P.l2_u, u.y, dsp * vi.y + dsn * vj.y);
                gu.z = fmaf(P.l2_u, u.z, dsp * vi.z + dsn * vj.z); gu.w = fmaf(P.l2_u, u.w, dsp * vi.w + dsn * vj.w);
                gi.x = fmaf(P.l2_i, vi.x, dsp * u.x); gi.y = fmaf(P.l2_i, vi.y, dsp * u.y); gi.z = fmaf(P.l2_i, vi.z, dsp * u.z); gi.w = fmaf(P.l2_i, vi.w, dsp * u.w);
                gj.x = fmaf(P.l2_i, vj.x, dsn * u.x); gj.y = fmaf(P.l2_i, vj.y, dsn * u.y); gj.z = fmaf(P.l2_i, vj.z, dsn * u.z); gj.w = fmaf(P.l2_i, vj.w, dsn * u.w);
                float4 *gb = P.gradbuf + (size_t)(3 * b) * 16 + l16;
                // a solo row: p, m, v were fetched once -- apply step t and write them back, once
                if (solo_u) {
                    if (P.optimizer != SML_OPT_ADAM_DENSE_EXACT) { ru.m = *slot(buf, 1); ru.v = *slot(buf, 2); }
                    adam4(ru.p, ru.m, ru.v, gu, P, ct); store_row(P.user_tab, P.m_user, P.v_user, u0, l16, ru);
                } else gb[0] = gu;
                if (solo_i) {
                    if (P.optimizer != SML_OPT_ADAM_DENSE_EXACT) { ri.m = *slot(buf, 4); ri.v = *slot(buf, 5); }
                    adam4(ri.p, ri.m, ri.v, gi, P, ct); store_row(P.item_tab, P.m_item, P.v_item, i0, l16, ri);
                } else gb[16] = gi;
                if (solo_j) {
                    if (P.optimizer != SML_OPT_ADAM_DENSE_EXACT) { rj.m = *slot(buf, 7); rj.v = *slot(buf, 8); }
                    adam4(rj.p, rj.m, rj.v, gj, P, ct); store_row(P.item_tab, P.m_item, P.v_item, j0, l16, rj);
                } else gb[32] = gj;
                if (l16 < 3) {
                    const int64_t id = l16 == 0 ? u0 : (l16 == 1 ? i0 : j0);
                    if (solo_l) {                              // re-arm the list head, stamp the row
                        (l16 == 0 ? P.head_user : P.head_item)[id] = -1;
                        if (P.optimizer == SML_OPT_ADAM_DENSE_EXACT) (l16 == 0 ? P.stamp_user : P.stamp_item)[id] = t;
                    } else if (meta0.x == -1) {                // first of several occurrences of its row: phase 2 work
                        P.worklist[atomicAdd(work_count, 1u)] = o;
                    }
                }
            }
            u0 = u1; i0 = i1; j0 = j1; u1 = u2; i1 = i2; j1 = j2; meta0 = meta1;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    acc_loss += __shfl_xor_sync(0xffffffffu, acc_loss, 16);
    acc_l2 += __shfl_xor_sync(0xffffffffu, acc_l2, 16);
    if (lane == 0) { s_part[w][0] = acc_loss; s_part[w][1] = acc_l2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, q = 0.f;
        for (int i = 0; i < PMS_THREADS / 32; ++i) { a += s_part[i][0]; q += s_part[i][1]; }
        P.partials[2 * blockIdx.x] = a; P.partials[2 * blockIdx.x + 1] = q;
    }
    if (tid == 0) P.work_count[(seq + 1u) & 1u] = 0;       // the next launch's counter (this launch's was zeroed by the previous one)
    grid.sync();
    // ---------------- phase 2: rows with several occurrences -- the first one sums the chain and applies step t ----------------
    // phase 1 listed those first occurrences; one half-warp per entry
    const int64_t n_work = (P.phases & 4) ? (int64_t)__ldcg(work_count) : 0;
    for (int64_t idx = hw0; idx < n_work; idx += nhw) {
        const int32_t o = __ldcg(P.worklist + idx);
        const uint32_t b = (uint32_t)o / 3u, k = (uint32_t)o - 3u * b;
        const int64_t id = k == 0 ? __ldg(P.user + b) : (k == 1 ? __ldg(P.item + b) : __ldg(P.neg + b));
        float *tab = k == 0 ? P.user_tab : P.item_tab, *mt = k == 0 ? P.m_user : P.m_item, *vt = k == 0 ? P.v_user : P.v_item;
        int32_t *stamp = k == 0 ? P.stamp_user : P.stamp_item, *head = (k == 0 ? P.head_user : P.head_item) + id;
        // ONE lane reads the list head and the stamp and hands them to the other 15: lane 0 overwrites both below, and the
        // lanes of a half-warp are not guaranteed to run in lockstep
        const unsigned hmask = 0xFFFFu << (lane & 16);
        const int first = __shfl_sync(hmask, l16 == 0 ? __ldcg(head) : 0, 0, 16);
        const int st = __shfl_sync(hmask, (l16 == 0 && P.optimizer == SML_OPT_ADAM_DENSE_EXACT) ? __ldcg(stamp + id) : 0, 0, 16);
        RowRegs r = load_row(tab, mt, vt, id, l16);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = first; c != -1; c = __ldcg(P.next + c)) {                   // newest occurrence first
            const float4 x = __ldcg(P.gradbuf + (size_t)c * 16 + l16);
            g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
        }
        if (P.optimizer == SML_OPT_ADAM_DENSE_EXACT) catch_up(r, st, t_prev, P);
        adam4(r.p, r.m, r.v, g, P, ct);
        store_row(tab, mt, vt, id, l16, r);
        if (l16 == 0) {
            *head = -1;                                        // list heads are all -1 again when the kernel ends
            if (P.optimizer == SML_OPT_ADAM_DENSE_EXACT) stamp[id] = t;
        }
    }
    if (tid == 0) {
        float a = 0.f, q = 0.f;
        for (unsigned i = 0; i < gridDim.x; ++i) { a += __ldcg(P.partials + 2 * i); q += __ldcg(P.partials + 2 * i + 1); }
        const float loss = (P.loss_kind == SML_LOSS_BCE) ? (-(a * invB) + q) : a;
        P.loss_out[0] = loss;
        P.loss_out[1] += loss;
        P.work_count[2] = seq + 1u;
        P.state[0] = t;                                    // every CTA read the old counter before the first grid sync
        float *f = reinterpret_cast<float *>(P.state + 1);
        f[0] = ct.x; f[1] = ct.y;
    }
}

size_t ws_bytes(int64_t B) {
    return 256 + 2 * PMS_MAX_BLOCKS * sizeof(float) + sml_align_up((size_t)3 * B * SML_D * sizeof(float), 256) +
           2 * sml_align_up((size_t)3 * B * sizeof(int32_t), 256) + 256;
}
constexpr size_t PMS_SMEM = (size_t)2 * 9 * PMS_THREADS * sizeof(float4);      // 73 728 B

}  // namespace

extern "C" {

size_t sml_plain_mf_step_workspace_bytes(int64_t batch) { return batch > 0 ? ws_bytes(batch) : 0; }

int sml_plain_mf_step(float *user_tab, float *item_tab, float *m_user, float *v_user, float *m_item, float *v_item,
                      int32_t *stamp_user, int32_t *stamp_item, int32_t *head_user, int32_t *head_item, const int64_t *user,
                      const int64_t *item, const int64_t *neg, int64_t batch, int d, int loss, double l2_u, double l2_i,
                      int64_t *adam_state, double lr, int optimizer, float *loss_out, void *workspace, size_t workspace_bytes,
                      void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_plain_mf_step: d=%d unsupported (d must be %d)", d, SML_D);
    SML_REQUIRE(user_tab && item_tab && m_user && v_user && m_item && v_item && head_user && head_item && user && item && neg &&
                adam_state && loss_out, SML_E_BADARG, "sml_plain_mf_step: null pointer");
    SML_REQUIRE(loss == SML_LOSS_BCE || loss == SML_LOSS_BPR, SML_E_BADARG, "sml_plain_mf_step: bad loss kind %d", loss);
    SML_REQUIRE(optimizer == SML_OPT_ADAM_DENSE_EXACT || optimizer == SML_OPT_ADAM_SPARSE, SML_E_BADARG,
                "sml_plain_mf_step: bad optimizer mode %d", optimizer);
    SML_REQUIRE(optimizer != SML_OPT_ADAM_DENSE_EXACT || (stamp_user && stamp_item), SML_E_BADARG,
                "sml_plain_mf_step: the exact dense-Adam mode needs the per-row stamps");
    SML_REQUIRE(batch > 0 && 3 * batch < (int64_t)1 << 31, SML_E_BADARG, "sml_plain_mf_step: bad batch size");
    SML_REQUIRE(workspace && workspace_bytes >= ws_bytes(batch), SML_E_WORKSPACE, "sml_plain_mf_step: workspace too small (%zu < %zu bytes)",
                workspace_bytes, ws_bytes(batch));
    PmsParams P;
    P.user_tab = user_tab; P.item_tab = item_tab; P.m_user = m_user; P.v_user = v_user; P.m_item = m_item; P.v_item = v_item;
    P.stamp_user = stamp_user; P.stamp_item = stamp_item; P.head_user = head_user; P.head_item = head_item;
    P.user = user; P.item = item; P.neg = neg; P.B = batch; P.loss_kind = loss; P.optimizer = optimizer;
    P.l2_u = (float)l2_u; P.l2_i = (float)l2_i; P.state = adam_state; P.lr = lr;
    P.b1c = (float)(1.0 - 0.9); P.beta2 = (float)0.999; P.b2c = (float)(1.0 - 0.999); P.eps = 1e-8f;
    P.loss_out = loss_out;
    { const char *e = getenv("SML_PMS_PHASES"); P.phases = e ? atoi(e) : 7; }
    // fixed-position items first (the two work counters must be zero between steps whatever the batch size was)
    char *p = (char *)workspace;
    P.work_count = (unsigned int *)p; p += 256;
    P.partials = (float *)p; p += 2 * PMS_MAX_BLOCKS * sizeof(float);
    P.gradbuf = (float4 *)p; p += sml_align_up((size_t)3 * batch * SML_D * sizeof(float), 256);
    P.next = (int32_t *)p; p += sml_align_up((size_t)3 * batch * sizeof(int32_t), 256);
    P.worklist = (int32_t *)p;
    static int per_sm = 0, minb = 0;
    if (!per_sm) {
        const char *e = getenv("SML_PMS_MINB");            // tuning aid
        minb = (e && atoi(e) == 3) ? 3 : 2;
        SML_CUDA_OK(cudaFuncSetAttribute(k_plain_mf_step<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PMS_SMEM));
        SML_CUDA_OK(cudaFuncSetAttribute(k_plain_mf_step<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PMS_SMEM));
        if (minb == 3) SML_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_plain_mf_step<3>, PMS_THREADS, PMS_SMEM));
        else SML_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_plain_mf_step<2>, PMS_THREADS, PMS_SMEM));
        if (per_sm < 1) per_sm = 1;
    }
    int64_t blocks = (int64_t)sml_sm_count() * per_sm;      // cooperative launch: every CTA must be resident
    const int64_t need = (3 * batch + PMS_HALFWARPS - 1) / PMS_HALFWARPS;
    if (blocks > need) blocks = need;
    if (blocks > PMS_MAX_BLOCKS) blocks = PMS_MAX_BLOCKS;
    void *args[] = {&P};
    void *kern = minb == 3 ? (void *)k_plain_mf_step<3> : (void *)k_plain_mf_step<2>;
    SML_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(PMS_THREADS), args, PMS_SMEM, (cudaStream_t)stream));
    SML_LAUNCH_OK();
    return SML_OK;
}

}  // extern "C"
