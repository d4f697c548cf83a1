// Row-exchange helpers for row-sharded tables (BASELINE.json north_star item 4).
// Tables are sharded by id across ranks (owner = id % world, local row = id / world); a step needs the
// (w_{t-1}, w_hat) pair of every id in its batch and returns one row gradient per id:
//   * k_gather_pairs : owner side, out[n] = [last[loc_n] | hat[loc_n]]  (512 B per requested id, ready for the
//                      NCCL all-to-all that answers the id request)
//   * k_scatter_grads: owner side, g[loc_n] += scale * d_row[n] + l2 * hat[loc_n]  (the dense-gradient scatter of
//                      model/transfer.py:486-502 applied to the rows that came back over NVLink)
// Both are pure 128-bit streaming kernels (16 lanes x 16 B per row), HBM / NVLink bound.
#include "sml_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_gather_pairs(const float4 *__restrict__ last, const float4 *__restrict__ hat, const int64_t *__restrict__ loc, int64_t n,
               float4 *__restrict__ out) {
    const int l16 = threadIdx.x & 15;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; r < n; r += ((int64_t)gridDim.x * blockDim.x) >> 4) {
        const int64_t id = __ldg(loc + r);
        out[r * 32 + l16] = __ldg(last + id * 16 + l16);
        out[r * 32 + 16 + l16] = __ldg(hat + id * 16 + l16);
    }
}

__global__ void __launch_bounds__(256)
k_scatter_grads(float *__restrict__ g, const float4 *__restrict__ hat, const int64_t *__restrict__ loc, const float4 *__restrict__ d_rows,
                int64_t n, float scale, float l2) {
    const int l16 = threadIdx.x & 15;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4; r < n; r += ((int64_t)gridDim.x * blockDim.x) >> 4) {
        const int64_t id = __ldg(loc + r);
        const float4 d = __ldg(d_rows + r * 16 + l16);
        const float4 h = __ldg(hat + id * 16 + l16);
        float *dst = g + id * SML_D + 4 * l16;
        atomicAdd(dst + 0, fmaf(l2, h.x, scale * d.x));
        atomicAdd(dst + 1, fmaf(l2, h.y, scale * d.y));
        atomicAdd(dst + 2, fmaf(l2, h.z, scale * d.z));
        atomicAdd(dst + 3, fmaf(l2, h.w, scale * d.w));
    }
}

int grid_rows(int64_t n) {
    int64_t b = (n * 16 + 255) / 256;
    const int64_t cap = (int64_t)sml_sm_count() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int sml_gather_pairs(const float *last, const float *hat, const int64_t *loc, int64_t n, int d, float *out, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_gather_pairs: d=%d unsupported (d must be %d)", d, SML_D);
    if (n <= 0) return SML_OK;
    SML_REQUIRE(last && hat && loc && out, SML_E_BADARG, "sml_gather_pairs: null pointer");
    k_gather_pairs<<<grid_rows(n), 256, 0, (cudaStream_t)stream>>>((const float4 *)last, (const float4 *)hat, loc, n, (float4 *)out);
    SML_LAUNCH_OK();
    return SML_OK;
}

int sml_scatter_grads(float *g, const float *hat, const int64_t *loc, const float *d_rows, int64_t n, int d, double scale,
                      double l2, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    SML_REQUIRE(d == SML_D, SML_E_UNSUPPORTED, "sml_scatter_grads: d=%d unsupported (d must be %d)", d, SML_D);
    if (n <= 0) return SML_OK;
    SML_REQUIRE(g && hat && loc && d_rows, SML_E_BADARG, "sml_scatter_grads: null pointer");
    k_scatter_grads<<<grid_rows(n), 256, 0, (cudaStream_t)stream>>>(g, (const float4 *)hat, loc, (const float4 *)d_rows, n,
                                                                     (float)scale, (float)l2);
    SML_LAUNCH_OK();
    return SML_OK;
}

}  // extern "C"
