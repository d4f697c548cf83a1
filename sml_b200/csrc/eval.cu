// Candidate-list evaluation: fused gather - dot - rank-count.
//
// Replaces MFbasemode.test (model/MF.py:45-80 of the reference): instead of materialising
// [b, 1000, 64] gathered item rows (262 MB per 1024-row batch), a [b, 1000] score matrix and a
// topk, every test row is handled by one CTA that streams its 1 + C item rows once and counts
// how many candidates beat the positive.  With one positive per row, "index 0 is in the top-K"
// <=> rank(positive) < K, so no sort is needed and one pass serves every K.
//
// Memory behaviour: per test row 1 user row + C item rows of 256 B + (1+C) int64 ids
// (264 264 B at C = 1000).  The Yelp / Adressa item tables (31 MB / 5 MB) stay L2 resident, so
// the kernel is bound by L2->SM bandwidth; 16 lanes x 16 B cover one 256 B row, i.e. every
// LDG.128 warp instruction moves two full rows (4 x 128 B lines, fully sector-efficient).
#include <cuda_bf16.h>

#include "sml_common.cuh"

namespace {

constexpr int EVAL_THREADS = 256;
constexpr int EVAL_HW = EVAL_THREADS / 16;  // half-warps per CTA
constexpr int EVAL_UNROLL = 4;

__device__ __forceinline__ float hw_sum(float v, unsigned mask) {
    // reduce inside a 16-lane half-warp; `mask` names only that half's lanes because the two
    // halves of a warp can run different trip counts of the candidate loop
    v += __shfl_xor_sync(mask, v, 8);
    v += __shfl_xor_sync(mask, v, 4);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 1);
    return v;
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    float s = a.x * b.x;
    s = fmaf(a.y, b.y, s);
    s = fmaf(a.z, b.z, s);
    s = fmaf(a.w, b.w, s);
    return s;
}

__global__ void __launch_bounds__(EVAL_THREADS)
k_eval_candidates(const float *__restrict__ user_tab, const float *__restrict__ item_tab,
                  const int64_t *__restrict__ rows, int64_t n_rows, int64_t row_stride, int n_cand,
                  int32_t *__restrict__ gt_out, int32_t *__restrict__ eq_out) {
    extern __shared__ int64_t s_ids[];  // [n_cand]
    __shared__ int s_gt, s_eq;
    const int tid = threadIdx.x;
    const int hw = tid >> 4;        // half-warp id in the CTA
    const int l16 = tid & 15;       // lane inside the half-warp: owns floats [4*l16, 4*l16+4)
    const unsigned hmask = 0xffffu << (16 * (hw & 1));
    const float4 *item4 = reinterpret_cast<const float4 *>(item_tab);
    const float4 *user4 = reinterpret_cast<const float4 *>(user_tab);

    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t *row = rows + r * row_stride;
        if (tid == 0) { s_gt = 0; s_eq = 0; }
        for (int c = tid; c < n_cand; c += EVAL_THREADS) s_ids[c] = __ldg(row + 1 + c);
        const float4 u = __ldg(user4 + __ldg(row) * (SML_D / 4) + l16);
        __syncthreads();
        const float s0 = hw_sum(dot4(u, __ldg(item4 + s_ids[0] * (SML_D / 4) + l16)), hmask);
        const bool s0_nan = s0 != s0;
        int gt = 0, eq = 0;
        // candidates 1.. are dealt round-robin to the half-warps, EVAL_UNROLL loads in flight each
        int c = 1 + hw;
        for (; c + (EVAL_UNROLL - 1) * EVAL_HW < n_cand; c += EVAL_UNROLL * EVAL_HW) {
            float4 v[EVAL_UNROLL];
#pragma unroll
            for (int k = 0; k < EVAL_UNROLL; ++k) v[k] = __ldg(item4 + s_ids[c + k * EVAL_HW] * (SML_D / 4) + l16);
#pragma unroll
            for (int k = 0; k < EVAL_UNROLL; ++k) {
                const float s = hw_sum(dot4(u, v[k]), hmask);
                const bool s_nan = s != s;
                gt += (s > s0) || (s_nan && !s0_nan);
                eq += (s == s0) || (s_nan && s0_nan);
            }
        }
        for (; c < n_cand; c += EVAL_HW) {
            const float s = hw_sum(dot4(u, __ldg(item4 + s_ids[c] * (SML_D / 4) + l16)), hmask);
            const bool s_nan = s != s;
            gt += (s > s0) || (s_nan && !s0_nan);
            eq += (s == s0) || (s_nan && s0_nan);
        }
        if (l16 == 0) {
            if (gt) atomicAdd(&s_gt, gt);
            if (eq) atomicAdd(&s_eq, eq);
        }
        __syncthreads();
        if (tid == 0) { gt_out[r] = s_gt; eq_out[r] = s_eq; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Same counts from half the bytes: bf16 pre-filter + exact fp32 fallback.
//
// k_eval_candidates runs at the L2->SM limit, so the only way to go faster is to read fewer bytes.  A candidate only has to be
// COMPARED with the positive: s_j is first estimated from a bf16 copy of the item table (128 B per row instead of 256 B)
// together with a rigorous bound on the estimate's error, |s_j - s~_j| <= 0.002 * sum_k |u_k| |i~_k| + 1e-30 (bf16 rounds to
// nearest: relative error <= 2^-9 per element; the fp32 summation errors of both dot products are three orders smaller).
// If |s~_j - s_0| exceeds the bound the comparison is decided; otherwise (a few % of the candidates, and every NaN / Inf
// case, because the test is written so that NaN fails it) the candidate's fp32 row is fetched and its score is recomputed
// with EXACTLY the operation order of k_eval_candidates (lane l of 8 owns floats [4l, 4l+4) and [4l+32, 4l+36): its two
// partial dots are what lanes l and l+8 of the 16-lane version hold, the first xor-8 shuffle step becomes a local add).
// gt / eq are therefore identical to the exact kernel's, bit for bit, for every input.
__global__ void __launch_bounds__(256) k_to_bf16(const float4 *__restrict__ src, uint2 *__restrict__ dst, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + i);
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        dst[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b));
    }
}

constexpr int EVAL_QW = EVAL_THREADS / 8;   // quarter-warps per CTA

__device__ __forceinline__ float qw_sum(float v, unsigned mask) {
    v += __shfl_xor_sync(mask, v, 4);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 1);
    return v;
}

// exact score with the summation order of k_eval_candidates (see above)
__device__ __forceinline__ float exact_score8(const float4 *__restrict__ item4, int64_t id, int l8, const float4 u_lo, const float4 u_hi,
                                              unsigned mask) {
    const float4 v_lo = __ldg(item4 + id * (SML_D / 4) + l8), v_hi = __ldg(item4 + id * (SML_D / 4) + l8 + 8);
    return qw_sum(dot4(u_lo, v_lo) + dot4(u_hi, v_hi), mask);
}

__global__ void __launch_bounds__(EVAL_THREADS)
k_eval_prefilter(const float *__restrict__ user_tab, const float *__restrict__ item_tab, const uint4 *__restrict__ item_bf16,
                 const int64_t *__restrict__ rows, int64_t n_rows, int64_t row_stride, int n_cand,
                 int32_t *__restrict__ gt_out, int32_t *__restrict__ eq_out) {
    extern __shared__ int64_t s_ids[];  // [n_cand]
    __shared__ int s_gt, s_eq;
    const int tid = threadIdx.x;
    const int qw = tid >> 3;        // quarter-warp id in the CTA
    const int l8 = tid & 7;         // lane inside the quarter-warp
    const unsigned qmask = 0xffu << (8 * (qw & 3));
    const float4 *item4 = reinterpret_cast<const float4 *>(item_tab);
    const float4 *user4 = reinterpret_cast<const float4 *>(user_tab);

    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t *row = rows + r * row_stride;
        if (tid == 0) { s_gt = 0; s_eq = 0; }
        for (int c = tid; c < n_cand; c += EVAL_THREADS) s_ids[c] = __ldg(row + 1 + c);
        const int64_t uid = __ldg(row);
        // exact path operands: floats [4 l8, +4) and [4 l8 + 32, +4); estimate operands: floats [8 l8, +8)
        const float4 u_lo = __ldg(user4 + uid * (SML_D / 4) + l8), u_hi = __ldg(user4 + uid * (SML_D / 4) + l8 + 8);
        const float4 ua = __ldg(user4 + uid * (SML_D / 4) + 2 * l8), ub = __ldg(user4 + uid * (SML_D / 4) + 2 * l8 + 1);
        __syncthreads();
        const float s0 = exact_score8(item4, s_ids[0], l8, u_lo, u_hi, qmask);
        const bool s0_nan = s0 != s0;
        int gt = 0, eq = 0;
        for (int c0 = 1 + qw; c0 < n_cand; c0 += EVAL_UNROLL * EVAL_QW) {
            uint4 w[EVAL_UNROLL];
#pragma unroll
            for (int k = 0; k < EVAL_UNROLL; ++k) {
                const int c = c0 + k * EVAL_QW;
                w[k] = c < n_cand ? __ldg(item_bf16 + s_ids[c] * (SML_D / 8) + l8) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int k = 0; k < EVAL_UNROLL; ++k) {
                const int c = c0 + k * EVAL_QW;
                if (c >= n_cand) break;                      // uniform over the quarter-warp
                const float i0 = __uint_as_float(w[k].x << 16), i1 = __uint_as_float(w[k].x & 0xffff0000u);
                const float i2 = __uint_as_float(w[k].y << 16), i3 = __uint_as_float(w[k].y & 0xffff0000u);
                const float i4 = __uint_as_float(w[k].z << 16), i5 = __uint_as_float(w[k].z & 0xffff0000u);
                const float i6 = __uint_as_float(w[k].w << 16), i7 = __uint_as_float(w[k].w & 0xffff0000u);
                float se = ua.x * i0, ae = fabsf(ua.x) * fabsf(i0);
                se = fmaf(ua.y, i1, se); ae = fmaf(fabsf(ua.y), fabsf(i1), ae);
                se = fmaf(ua.z, i2, se); ae = fmaf(fabsf(ua.z), fabsf(i2), ae);
                se = fmaf(ua.w, i3, se); ae = fmaf(fabsf(ua.w), fabsf(i3), ae);
                se = fmaf(ub.x, i4, se); ae = fmaf(fabsf(ub.x), fabsf(i4), ae);
                se = fmaf(ub.y, i5, se); ae = fmaf(fabsf(ub.y), fabsf(i5), ae);
                se = fmaf(ub.z, i6, se); ae = fmaf(fabsf(ub.z), fabsf(i6), ae);
                se = fmaf(ub.w, i7, se); ae = fmaf(fabsf(ub.w), fabsf(i7), ae);
                se = qw_sum(se, qmask); ae = qw_sum(ae, qmask);
                const float d = se - s0;
                if (fabsf(d) > fmaf(0.002f, ae, 1e-30f)) {   // decided by the estimate (false for any NaN / Inf involved)
                    gt += d > 0.f;
                } else {
                    const float s = exact_score8(item4, s_ids[c], l8, u_lo, u_hi, qmask);
                    const bool s_nan = s != s;
                    gt += (s > s0) || (s_nan && !s0_nan);
                    eq += (s == s0) || (s_nan && s0_nan);
                }
            }
        }
        if (l8 == 0) {
            if (gt) atomicAdd(&s_gt, gt);
            if (eq) atomicAdd(&s_eq, eq);
        }
        __syncthreads();
        if (tid == 0) { gt_out[r] = s_gt; eq_out[r] = s_eq; }
        __syncthreads();
    }
}

// hits / NDCG per batch of test rows, fixed summation order (one CTA per batch, tree reduce).
__global__ void __launch_bounds__(256)
k_eval_reduce(const int32_t *__restrict__ gt, const int32_t *__restrict__ eq, int64_t n_rows, int batch, int topk,
              int tie_loses, int32_t *__restrict__ hits, float *__restrict__ ndcg) {
    __shared__ float s_nd[256];
    __shared__ int s_hit[256];
    const int64_t lo = (int64_t)blockIdx.x * batch;
    const int64_t hi = min(lo + (int64_t)batch, n_rows);
    float nd = 0.0f;
    int h = 0;
    for (int64_t r = lo + threadIdx.x; r < hi; r += blockDim.x) {
        const int rank = gt[r] + (tie_loses ? eq[r] : 0);
        if (rank < topk) {
            h += 1;
            nd += 1.0f / log2f((float)rank + 2.0f);   // model/MF.py:74
        }
    }
    s_nd[threadIdx.x] = nd;
    s_hit[threadIdx.x] = h;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_nd[threadIdx.x] += s_nd[threadIdx.x + o];
            s_hit[threadIdx.x] += s_hit[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { hits[blockIdx.x] = s_hit[0]; ndcg[blockIdx.x] = s_nd[0]; }
}

// MFbasemode.forward scores (model/MF.py:34-43): one half-warp per (user, item) pair.
__global__ void __launch_bounds__(256)
k_pair_scores(const float *__restrict__ user_tab, const float *__restrict__ item_tab, const int64_t *__restrict__ user,
              const int64_t *__restrict__ item, int64_t n, int norm, float *__restrict__ score) {
    const int l16 = threadIdx.x & 15;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (p >= n) return;   // whole half-warps exit together; shuffles below use the full mask of live lanes
    const float4 u = __ldg(reinterpret_cast<const float4 *>(user_tab) + __ldg(user + p) * (SML_D / 4) + l16);
    const float4 v = __ldg(reinterpret_cast<const float4 *>(item_tab) + __ldg(item + p) * (SML_D / 4) + l16);
    const unsigned mask = __activemask();
    float s = dot4(u, v), q = dot4(u, u);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        s += __shfl_xor_sync(mask, s, o);
        q += __shfl_xor_sync(mask, q, o);
    }
    if (l16 == 0) score[p] = norm ? s / sqrtf(q) : s;
}

}  // namespace

extern "C" {

int sml_eval_candidates(const float *user_tab, const float *item_tab, int d, const int64_t *rows, int64_t n_rows,
                        int64_t row_stride, int n_cand, int32_t *gt, int32_t *eq, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_eval_candidates: d=%d unsupported (d must be %d)", d, SML_D);
    SML_REQUIRE(n_rows >= 0, SML_E_BADARG, "sml_eval_candidates: negative n_rows");
    if (n_rows == 0) return SML_OK;   // empty test file: nothing to do (pointers of empty tensors may be null)
    SML_REQUIRE(user_tab && item_tab && rows && gt && eq, SML_E_BADARG, "sml_eval_candidates: null pointer");
    SML_REQUIRE(n_cand >= 1 && row_stride >= 1 + (int64_t)n_cand, SML_E_BADARG,
                "sml_eval_candidates: need n_cand >= 1 and row_stride >= 1 + n_cand (got %d, %lld)", n_cand,
                (long long)row_stride);
    const size_t smem = (size_t)n_cand * sizeof(int64_t);
    SML_REQUIRE(smem <= 200 * 1024, SML_E_UNSUPPORTED, "sml_eval_candidates: n_cand=%d too large", n_cand);
    if (smem > 48 * 1024)
        SML_CUDA_OK(cudaFuncSetAttribute(k_eval_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sml_sm_count();
    const int64_t max_grid = (int64_t)sms * 8;
    const int grid = (int)(n_rows < max_grid ? n_rows : max_grid);
    k_eval_candidates<<<grid, EVAL_THREADS, smem, (cudaStream_t)stream>>>(user_tab, item_tab, rows, n_rows, row_stride,
                                                                         n_cand, gt, eq);
    SML_LAUNCH_OK();
    return SML_OK;
}

size_t sml_eval_prefilter_bytes(int64_t n_items) { return n_items > 0 ? (size_t)n_items * SML_D * 2 : 0; }

int sml_eval_candidates_prefilter(const float *user_tab, const float *item_tab, int64_t n_items, int d, void *item_bf16,
                                  const int64_t *rows, int64_t n_rows, int64_t row_stride, int n_cand, int32_t *gt, int32_t *eq,
                                  void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_eval_candidates_prefilter: d=%d unsupported (d must be %d)", d, SML_D);
    SML_REQUIRE(n_rows >= 0 && n_items >= 0, SML_E_BADARG, "sml_eval_candidates_prefilter: negative size");
    if (n_rows == 0) return SML_OK;
    SML_REQUIRE(user_tab && item_tab && item_bf16 && rows && gt && eq, SML_E_BADARG, "sml_eval_candidates_prefilter: null pointer");
    SML_REQUIRE(((uintptr_t)item_bf16 & 15) == 0, SML_E_BADARG, "sml_eval_candidates_prefilter: scratch must be 16-byte aligned");
    SML_REQUIRE(n_cand >= 1 && row_stride >= 1 + (int64_t)n_cand, SML_E_BADARG,
                "sml_eval_candidates_prefilter: need n_cand >= 1 and row_stride >= 1 + n_cand (got %d, %lld)", n_cand,
                (long long)row_stride);
    const size_t smem = (size_t)n_cand * sizeof(int64_t);
    SML_REQUIRE(smem <= 200 * 1024, SML_E_UNSUPPORTED, "sml_eval_candidates_prefilter: n_cand=%d too large", n_cand);
    if (smem > 48 * 1024)
        SML_CUDA_OK(cudaFuncSetAttribute(k_eval_prefilter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sml_sm_count();
    // the bf16 copy is rebuilt on every call (47 MB of traffic at the Yelp shape, ~1 % of the scoring pass): the tables are
    // updated through raw pointers, so a cached copy could go stale
    const int64_t n4 = n_items * (SML_D / 4);
    int64_t cb = (n4 + 255) / 256;
    if (cb > (int64_t)sms * 16) cb = (int64_t)sms * 16;
    k_to_bf16<<<(int)cb, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(item_tab), reinterpret_cast<uint2 *>(item_bf16), n4);
    SML_LAUNCH_OK();
    const int64_t max_grid = (int64_t)sms * 8;
    const int grid = (int)(n_rows < max_grid ? n_rows : max_grid);
    k_eval_prefilter<<<grid, EVAL_THREADS, smem, (cudaStream_t)stream>>>(user_tab, item_tab, reinterpret_cast<const uint4 *>(item_bf16), rows,
                                                                        n_rows, row_stride, n_cand, gt, eq);
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_eval_reduce(const int32_t *gt, const int32_t *eq, int64_t n_rows, int batch, int topk, int tie_loses,
                    int32_t *hits, float *ndcg, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(batch >= 1 && topk >= 1 && n_rows >= 0, SML_E_BADARG, "sml_eval_reduce: bad batch/topk/n_rows");
    if (n_rows == 0) return SML_OK;
    SML_REQUIRE(gt && eq && hits && ndcg, SML_E_BADARG, "sml_eval_reduce: null pointer");
    const int64_t nb = (n_rows + batch - 1) / batch;
    k_eval_reduce<<<(int)nb, 256, 0, (cudaStream_t)stream>>>(gt, eq, n_rows, batch, topk, tie_loses, hits, ndcg);
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_pair_scores(const float *user_tab, const float *item_tab, int d, const int64_t *user, const int64_t *item,
                    int64_t n, int norm, float *score, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_pair_scores: d=%d unsupported (d must be %d)", d, SML_D);
    SML_REQUIRE(user_tab && item_tab && user && item && score, SML_E_BADARG, "sml_pair_scores: null pointer");
    if (n <= 0) return SML_OK;
    const int64_t threads = n * 16;
    k_pair_scores<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(user_tab, item_tab, user, item, n,
                                                                                 norm, score);
    SML_LAUNCH_OK();
    return SML_OK;
}

}  // extern "C"
