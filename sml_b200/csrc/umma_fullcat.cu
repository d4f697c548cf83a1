// Full-catalog evaluation (BASELINE.json north_star item 3, config 5): for every evaluated (user, positive item)
// pair, the rank of the positive among ALL items of the catalog,
//     gt[u] = #{i != pos_u : <U_u, I_i> > <U_u, I_pos_u>},   eq[u] = #{i != pos_u : ... == ...},
// from which recall@K / NDCG@K follow exactly as in the candidate-list evaluation (model/MF.py:59-80 of the
// reference, which only ever ranks 1 + 999 candidates).  With one positive per row a count is enough: no
// sort, no top-k selection.
//
// This is a [users x 64] x [64 x items] score GEMM whose epilogue is a compare-and-count.  K = 64 is only two
// 32-wide chunks, so the kernel is epilogue / operand-streaming bound, not MMA bound (hard
// part 7).  Structure (persistent over the item tiles of one user tile):
//   * the 128-user A tile (packed 3xTF32 operand, 72 KB) is loaded once per CTA; item tiles (packed, 128 items,
//     72 KB) stream through a 2-stage cp.async.bulk ring;
//   * warp 1 issues 24 tcgen05.mma.kind::tf32 per item tile into one of TWO 128-column TMEM accumulators, so the
//     MMAs of tile j+1 overlap the epilogue of tile j (acc_full / acc_empty mbarriers);
//   * 8 epilogue warps read the scores with tcgen05.ld (one user row per thread), compare against the row's
//     positive score and count; the positive's own column is skipped by item id.  Counts accumulate in registers
//     over all item tiles and leave the SM as one atomicAdd per row.
// Items can be sharded across GPUs: every rank counts over its shard and the counts are summed (all-reduce).
#include "sml_common.cuh"
#include "umma_ptx.cuh"

namespace {

using namespace ptx;

constexpr int FC_THREADS = 320;
constexpr int FC_STAGES = 2;
constexpr uint32_t FC_TILE_BYTES = 2 * pk_block_bytes(128);   // 2 K chunks (K = 64), hi + lo: 73 728 B
constexpr uint32_t FC_SMEM = (1 + FC_STAGES) * FC_TILE_BYTES + 128;

// MODE 0: rank counts of the positive (gt / eq).  MODE 1: the diagonal of the score tile -- score of user u with the u-th row of
// the "item" operand (its gathered positive item), i.e. the positive's score from the very same tensor-core arithmetic the
// catalog scores come from.  MODE 2: per-thread top-K candidate lists (merged by k_fullcat_topk_merge).
struct FcTopk {
    float *scores;      // [gridDim.y * gridDim.x * 256][K]  (CTA-major, thread, slot)
    int64_t *ids;
    int *counts;        // [gridDim.y * gridDim.x * 256]
    int K;
};

template <int MODE>
__global__ void __launch_bounds__(FC_THREADS, 1)
k_fullcat_rank(const uint8_t *__restrict__ Upk, const uint8_t *__restrict__ Ipk, const float *__restrict__ s_pos,
               const int64_t *__restrict__ pos_id, int64_t n_users, int64_t n_items, int64_t item_id0, int tiles_per_split,
               int *__restrict__ gt, int *__restrict__ eq, float *__restrict__ diag_out, FcTopk T) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;
    uint8_t *sB = smem + FC_TILE_BYTES;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + (1 + FC_STAGES) * FC_TILE_BYTES);
    uint64_t *b_full = a_full + 1, *b_empty = b_full + FC_STAGES, *acc_full = b_empty + FC_STAGES, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_item_tiles = (n_items + 127) / 128;
    const int ut = blockIdx.y;
    const int64_t t0 = MODE == 1 ? (int64_t)ut : (int64_t)blockIdx.x * tiles_per_split;      // diagonal: user tile t against "item" tile t
    const int64_t t1 = MODE == 1 ? t0 + 1 : (t0 + tiles_per_split < n_item_tiles ? t0 + tiles_per_split : n_item_tiles);
    const int ntiles = (int)(t1 - t0);
    if (ntiles <= 0) {
        if (MODE == 2 && threadIdx.x >= 64) T.counts[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 256 + (threadIdx.x - 64)] = 0;
        return;
    }

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int i = 0; i < FC_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(a_full, FC_TILE_BYTES);
            bulk_g2s(sA, Upk + (size_t)ut * FC_TILE_BYTES, FC_TILE_BYTES, a_full);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % FC_STAGES;
                if (j >= FC_STAGES) mbar_wait(&b_empty[s], ((j / FC_STAGES) - 1) & 1);
                mbar_expect_tx(&b_full[s], FC_TILE_BYTES);
                bulk_g2s(sB + s * FC_TILE_BYTES, Ipk + (size_t)(t0 + j) * FC_TILE_BYTES, FC_TILE_BYTES, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        // the whole warp runs the loop (descriptors in uniform registers), one elected lane issues
        constexpr uint32_t IDESC = idesc_tf32(128);
        constexpr uint32_t DHI = desc_hi(PK_SBO), KSTEP = (2 * PK_LBO) >> 4;
        const bool leader = elect_one();
        mbar_wait(a_full, 0);
        for (int j = 0; j < ntiles; ++j) {
            const int s = j % FC_STAGES, acc = j & 1;
            mbar_wait(&b_full[s], (j / FC_STAGES) & 1);
            if (j >= 2) mbar_wait(&acc_empty[acc], ((j >> 1) - 1) & 1);          // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t d = tmem + acc * 128;
            if (leader) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t a_hi = desc_lo(smem_u32(sA) + c * pk_block_bytes(128), PK_LBO), a_lo = desc_lo(smem_u32(sA) + c * pk_block_bytes(128) + pk_half_bytes(128), PK_LBO);
                    const uint32_t b_hi = desc_lo(smem_u32(sB + s * FC_TILE_BYTES) + c * pk_block_bytes(128), PK_LBO),
                                   b_lo = desc_lo(smem_u32(sB + s * FC_TILE_BYTES) + c * pk_block_bytes(128) + pk_half_bytes(128), PK_LBO);
#pragma unroll
                    for (int k = 0; k < PK_BK / 8; ++k) {
                        umma_tf32_h(d, a_lo + k * KSTEP, b_hi + k * KSTEP, DHI, IDESC, (c | k) != 0);
                        umma_tf32_h(d, a_hi + k * KSTEP, b_lo + k * KSTEP, DHI, IDESC, 1);
                        umma_tf32_h(d, a_hi + k * KSTEP, b_hi + k * KSTEP, DHI, IDESC, 1);
                    }
                }
                umma_commit(&b_empty[s]);
                umma_commit(&acc_full[acc]);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        const int64_t u = (int64_t)ut * 128 + quarter * 32 + lane;
        const bool live = u < n_users;
        const float sp = (MODE == 0 && live) ? __ldg(s_pos + u) : 0.f;
        const int64_t pid = (MODE != 1 && live && pos_id) ? __ldg(pos_id + u) : -1;
        const bool sp_nan = sp != sp;
        int c_gt = 0, c_eq = 0;
        // MODE 2: this thread's candidate list (unsorted, the minimum tracked): K slots in global scratch, touched only on the
        // rare insertion (~K ln(N / K) times over N items)
        const size_t tslot = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 256 + (threadIdx.x - 64);
        float *lsc = MODE == 2 ? T.scores + tslot * T.K : nullptr;
        int64_t *lid = MODE == 2 ? T.ids + tslot * T.K : nullptr;
        float thr = -INFINITY;
        int cnt = 0, amin = 0;
        for (int j = 0; j < ntiles; ++j) {
            const int acc = j & 1;
            mbar_wait(&acc_full[acc], (j >> 1) & 1);
            tc_fence_after();
            const int64_t item0 = item_id0 + (t0 + j) * 128;
#pragma unroll
            for (int c0 = chalf * 64; c0 < chalf * 64 + 64; c0 += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + acc * 128 + c0, v);
                if (MODE == 1) {
                    // the diagonal: user row r of the tile against item row r of the tile
                    const int r = quarter * 32 + lane;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i == r && live && j == 0) diag_out[u] = v[i];
                } else if (MODE == 0) {
                    const int64_t first = item0 + c0;
                    // a chunk that lies inside the catalog and does not hold the positive itself (all but one or two per row): two
                    // compares per score -- !(s <= sp) is "greater, or NaN" (NaN ranks highest), s == sp never holds for NaN
                    const bool plain = !sp_nan && (first + 32 - item_id0) <= n_items && (pid < first || pid >= first + 32);
                    if (plain) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { c_gt += !(v[i] <= sp); c_eq += (v[i] == sp); }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int64_t it = first + i;
                            const bool ok = (it - item_id0) < n_items && it != pid;
                            const bool vn = v[i] != v[i];
                            c_gt += ok && ((v[i] > sp) || (vn && !sp_nan));
                            c_eq += ok && ((v[i] == sp) || (vn && sp_nan));
                        }
                    }
                } else {
                    // hot path: one compare per score against the list's current minimum (NaN fails `<=` and is taken: it ranks
                    // highest, like torch.topk); only a chunk that holds a candidate, the excluded item or the end of the catalog
                    // goes through the per-element path
                    bool any = false;
#pragma unroll
                    for (int i = 0; i < 32; ++i) any |= !(v[i] <= thr);
                    const int64_t first = item0 + c0;
                    if (live && any) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {                                   // (unrolled: v[] must stay in registers)
                            if (v[i] <= thr) continue;
                            const int64_t it = first + i;
                            if ((it - item_id0) >= n_items || it == pid) continue;
                            const float sc = (v[i] != v[i]) ? INFINITY : v[i];
                            if (cnt < T.K) {
                                lsc[cnt] = sc; lid[cnt] = it; ++cnt;
                            } else {
                                lsc[amin] = sc; lid[amin] = it;
                            }
                            if (cnt == T.K) {                                            // new minimum of the full list
                                float m = lsc[0]; int am = 0;
                                for (int k = 1; k < T.K; ++k) { const float x = lsc[k]; if (x < m) { m = x; am = k; } }
                                thr = m; amin = am;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
        if (MODE == 0 && live) {
            if (c_gt) atomicAdd(gt + u, c_gt);
            if (c_eq) atomicAdd(eq + u, c_eq);
        }
        if (MODE == 2) T.counts[tslot] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem);
}

// Merge of the per-thread candidate lists of one user (2 column halves x `splits` item ranges) into its top K, descending:
// one warp per user, K rounds of a warp-wide arg-max over at most 2 * splits * K candidates.
__global__ void __launch_bounds__(256)
k_fullcat_topk_merge(FcTopk T, int splits, int64_t n_users, float *__restrict__ out_scores, int64_t *__restrict__ out_ids) {
    const int lane = threadIdx.x & 31;
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= n_users) return;
    const int64_t ut = u >> 7;
    const int r = (int)(u & 127), quarter = r >> 5, l = r & 31;
    const int ncand = 2 * splits * T.K;
    for (int k = 0; k < T.K; ++k) {
        float best = -INFINITY;
        int bc = -1;
        for (int c = lane; c < ncand; c += 32) {
            const int list = c / T.K, slot = c - list * T.K, split = list >> 1, chalf = list & 1;
            const size_t tslot = ((size_t)ut * splits + split) * 256 + (size_t)(chalf * 4 + ((quarter + 2) & 3)) * 32 + l;   // epilogue warp 2 + e: quarter = warp & 3
            if (slot < __ldg(T.counts + tslot)) {
                const float x = T.scores[tslot * T.K + slot];
                if (bc < 0 || x > best) { best = x; bc = c; }
            }
        }
        // warp arg-max (ties: the lowest candidate index, deterministic)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (oc >= 0 && (bc < 0 || ob > best || (ob == best && oc < bc))) { best = ob; bc = oc; }
        }
        if (bc < 0) {                                       // fewer than K items in the catalog
            if (lane == 0) { out_scores[u * T.K + k] = -INFINITY; out_ids[u * T.K + k] = -1; }
            continue;
        }
        const int list = bc / T.K, slot = bc - list * T.K, split = list >> 1, chalf = list & 1;
        const size_t tslot = ((size_t)ut * splits + split) * 256 + (size_t)(chalf * 4 + ((quarter + 2) & 3)) * 32 + l;
        if (lane == 0) {
            out_scores[u * T.K + k] = best;
            out_ids[u * T.K + k] = T.ids[tslot * T.K + slot];
            T.scores[tslot * T.K + slot] = -INFINITY;       // taken; -inf entries never win again unless nothing else is left
            T.ids[tslot * T.K + slot] = -1;
        }
        __syncwarp();
    }
}

// rows of a table (optionally gathered by ids) -> packed K-major operand with K = 64 (2 chunks), 128-row tiles
__global__ void __launch_bounds__(256)
k_pack_rows(const float *__restrict__ tab, const int64_t *__restrict__ ids, int64_t n, uint8_t *__restrict__ out) {
    // one 16-lane half-warp per row, lane q owns K quad q (floats 4q..4q+3)
    const int q = threadIdx.x & 15;
    const int64_t tiles = (n + 127) / 128;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; r < tiles * 128; r += ((int64_t)gridDim.x * blockDim.x) >> 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n) {
            const int64_t id = ids ? __ldg(ids + r) : r;
            t = __ldg(reinterpret_cast<const float4 *>(tab + id * SML_D) + q);
        }
        float h[4], l[4];
        pk_split(t.x, h[0], l[0]); pk_split(t.y, h[1], l[1]); pk_split(t.z, h[2], l[2]); pk_split(t.w, h[3], l[3]);
        const int64_t tile = r >> 7;
        const int rr = (int)(r & 127), k = 4 * q;
        uint8_t *blk = out + ((size_t)tile * 2 + (k >> 5)) * pk_block_bytes(128) + pk_elem_off(rr, k & 31);
        *reinterpret_cast<float4 *>(blk) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(blk + pk_half_bytes(128)) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

}  // namespace

extern "C" {

size_t sml_packed_rows_bytes(int64_t n_rows) { return (size_t)((n_rows + 127) / 128) * FC_TILE_BYTES; }

int sml_pack_rows(const float *tab, const int64_t *ids, int64_t n_rows, int d, void *out, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_pack_rows: d=%d unsupported (d must be %d)", d, SML_D);
    if (n_rows <= 0) return SML_OK;
    SML_REQUIRE(tab && out, SML_E_BADARG, "sml_pack_rows: null pointer");
    int64_t blocks = ((n_rows + 127) / 128 * 128 * 16 + 255) / 256;
    const int64_t cap = (int64_t)sml_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_pack_rows<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(tab, ids, n_rows, (uint8_t *)out);
    SML_LAUNCH_OK();
    return SML_OK;
}

static int fullcat_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        SML_CUDA_OK(cudaFuncSetAttribute(k_fullcat_rank<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_SMEM));
        SML_CUDA_OK(cudaFuncSetAttribute(k_fullcat_rank<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_SMEM));
        SML_CUDA_OK(cudaFuncSetAttribute(k_fullcat_rank<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_SMEM));
        attr_set = true;
    }
    return SML_OK;
}

int sml_fullcat_rank(const void *users_packed, const void *items_packed, const float *s_pos, const int64_t *pos_id, int64_t n_users,
                     int64_t n_items, int64_t item_id0, int32_t *gt, int32_t *eq, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    if (n_users <= 0 || n_items <= 0) return SML_OK;
    SML_REQUIRE(users_packed && items_packed && s_pos && pos_id && gt && eq, SML_E_BADARG, "sml_fullcat_rank: null pointer");
    rc = fullcat_attr();
    if (rc) return rc;
    const int64_t ut = (n_users + 127) / 128, it = (n_items + 127) / 128;
    // split the item tiles so that user_tiles x splits covers the SMs a few times over, but keep >= 8 tiles per CTA
    // to amortise the A-tile load and the pipeline fill
    const int sms = sml_sm_count();
    int64_t splits = (4 * sms + ut - 1) / ut;
    if (splits < 1) splits = 1;
    if (splits > it) splits = it;
    int64_t per = (it + splits - 1) / splits;
    if (per < 8 && it >= 8) per = 8;
    splits = (it + per - 1) / per;
    SML_REQUIRE(ut <= 65535, SML_E_UNSUPPORTED, "sml_fullcat_rank: at most 65535*128 users per call (got %lld)", (long long)n_users);
    dim3 grid((unsigned)splits, (unsigned)ut);
    k_fullcat_rank<0><<<grid, FC_THREADS, FC_SMEM, (cudaStream_t)stream>>>((const uint8_t *)users_packed, (const uint8_t *)items_packed, s_pos,
                                                                           pos_id, n_users, n_items, item_id0, (int)per, gt, eq, nullptr, FcTopk{});
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_fullcat_pos_scores(const void *users_packed, const void *pos_packed, int64_t n_users, float *s_pos, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    if (n_users <= 0) return SML_OK;
    SML_REQUIRE(users_packed && pos_packed && s_pos, SML_E_BADARG, "sml_fullcat_pos_scores: null pointer");
    rc = fullcat_attr();
    if (rc) return rc;
    const int64_t ut = (n_users + 127) / 128;
    SML_REQUIRE(ut <= 65535, SML_E_UNSUPPORTED, "sml_fullcat_pos_scores: at most 65535*128 users per call");
    k_fullcat_rank<1><<<dim3(1, (unsigned)ut), FC_THREADS, FC_SMEM, (cudaStream_t)stream>>>(
        (const uint8_t *)users_packed, (const uint8_t *)pos_packed, nullptr, nullptr, n_users, ut * 128, 0, 1, nullptr, nullptr, s_pos, FcTopk{});
    SML_LAUNCH_OK();
    return SML_OK;
}

static int64_t topk_splits(int64_t ut, int64_t it) {
    // as few item ranges per user tile as still fill the SMs (every list pays ~K ln(items per list / K) insertions while it warms
    // up, which is what a small catalog spends its time on), at most 8 (the merge scans 2 * splits * K candidates per user)
    const int sms = sml_sm_count();
    int64_t splits = (sms + ut - 1) / ut;
    if (splits > 8) splits = 8;
    if (splits < 1) splits = 1;
    if (splits > it) splits = it;
    return splits;
}

size_t sml_fullcat_topk_workspace_bytes(int64_t n_users, int64_t n_items, int k) {
    if (n_users <= 0 || n_items <= 0 || k <= 0) return 0;
    const int64_t ut = (n_users + 127) / 128, it = (n_items + 127) / 128;
    const size_t lists = (size_t)ut * topk_splits(ut, it) * 256;
    return lists * k * (sizeof(float) + sizeof(int64_t)) + lists * sizeof(int) + 512;
}

int sml_fullcat_topk(const void *users_packed, const void *items_packed, const int64_t *exclude_id, int64_t n_users, int64_t n_items,
                     int64_t item_id0, int k, float *out_scores, int64_t *out_ids, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    if (n_users <= 0) return SML_OK;
    SML_REQUIRE(users_packed && items_packed && out_scores && out_ids && n_items > 0, SML_E_BADARG, "sml_fullcat_topk: bad arguments");
    SML_REQUIRE(k >= 1 && k <= 64, SML_E_UNSUPPORTED, "sml_fullcat_topk: k must be in 1..64 (got %d)", k);
    SML_REQUIRE(workspace && workspace_bytes >= sml_fullcat_topk_workspace_bytes(n_users, n_items, k), SML_E_WORKSPACE,
                "sml_fullcat_topk: workspace too small");
    rc = fullcat_attr();
    if (rc) return rc;
    const int64_t ut = (n_users + 127) / 128, it = (n_items + 127) / 128;
    SML_REQUIRE(ut <= 65535, SML_E_UNSUPPORTED, "sml_fullcat_topk: at most 65535*128 users per call");
    const int64_t splits = topk_splits(ut, it);
    const int64_t per = (it + splits - 1) / splits;
    const size_t lists = (size_t)ut * splits * 256;
    FcTopk T;
    char *p = (char *)workspace;
    T.ids = (int64_t *)p; p += lists * k * sizeof(int64_t);
    T.scores = (float *)p; p += lists * k * sizeof(float);
    T.counts = (int *)p;
    T.K = k;
    k_fullcat_rank<2><<<dim3((unsigned)splits, (unsigned)ut), FC_THREADS, FC_SMEM, (cudaStream_t)stream>>>(
        (const uint8_t *)users_packed, (const uint8_t *)items_packed, nullptr, exclude_id, n_users, n_items, item_id0, (int)per, nullptr, nullptr,
        nullptr, T);
    SML_LAUNCH_OK();
    const int64_t threads = n_users * 32;
    k_fullcat_topk_merge<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, (int)splits, n_users, out_scores, out_ids);
    SML_LAUNCH_OK();
    return SML_OK;
}

}  // extern "C"
