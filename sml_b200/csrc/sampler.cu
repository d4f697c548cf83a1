// GPU negative sampler for throughput runs: the semantics of
// offlineDataset_withsample.__getitem__ (data/dataset.py:62-71 of the reference) -- draw a uniform item of
// this period's item set, redraw while it is one of the user's items of the period -- with a counter-based
// Philox4x32-10 stream per sample instead of the host's sequential numpy generator.  Parity runs keep using
// the host emulation (bit-identical to the reference); this kernel trades that for ~10^9 samples/s.
#include <curand_kernel.h>

#include "sml_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_philox_negatives(const int64_t *__restrict__ users, int64_t n, const int64_t *__restrict__ item_all, int64_t n_items_all,
                   const int64_t *__restrict__ keys, int64_t n_keys, int64_t span, uint64_t seed, uint64_t offset,
                   int64_t *__restrict__ neg) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)s, offset, &st);          // subsequence = sample index
        const int64_t u = users[s];
        int64_t it = 0;
        for (int tries = 0; tries < 1024; ++tries) {
            const uint4 r = curand4(&st);
            bool done = false;
#pragma unroll
            for (int j = 0; j < 4 && !done; ++j) {
                const uint32_t x = j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w;
                it = item_all[(int64_t)(((uint64_t)x * (uint64_t)n_items_all) >> 32)];   // Lemire multiply-shift
                const int64_t key = u * span + it;
                int64_t lo = 0, hi = n_keys;
                while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
                done = !(lo < n_keys && keys[lo] == key);
            }
            if (done) break;
        }
        neg[s] = it;
    }
}

}  // namespace

extern "C" int sml_philox_negatives(const int64_t *users, int64_t n, const int64_t *item_all, int64_t n_items_all,
                                    const int64_t *keys, int64_t n_keys, int64_t span, uint64_t seed, uint64_t offset,
                                    int64_t *neg, void *stream) {
    int rc = sml_check_device();
    if (rc) return rc;
    if (n <= 0) return SML_OK;
    SML_REQUIRE(users && item_all && keys && neg && n_items_all > 0, SML_E_BADARG, "sml_philox_negatives: bad arguments");
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sml_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_philox_negatives<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(users, n, item_all, n_items_all, keys, n_keys, span, seed, offset, neg);
    SML_LAUNCH_OK();
    return SML_OK;
}
