// Host-side helper (no device code): the sequential part of the reference's on-the-fly negative
// sampler.  offlineDataset_withsample.__getitem__ (data/dataset.py:62-71 of the reference) draws
// np.random.choice(item_all, 1) per sample and redraws while the item is one of the user's items of
// the period.  The random draws themselves are produced by numpy (so the global RNG stream stays
// bit-identical with the reference); this function only walks them: sample s consumes draws until one
// is accepted.  Membership is a binary search in the sorted (user * span + item) key array.
#include <stdint.h>

#include "sml_b200.h"

extern "C" int64_t sml_host_rejection_walk(const int64_t *draws, int64_t n_draws, const int64_t *users, int64_t n,
                                           const int64_t *item_all, const int64_t *keys, int64_t n_keys, int64_t span,
                                           int64_t *neg) {
    int64_t p = 0;
    for (int64_t s = 0; s < n; ++s) {
        for (;;) {
            if (p >= n_draws) return -1;                 // ran out of draws: caller retries with more
            const int64_t it = item_all[draws[p++]];
            const int64_t key = users[s] * span + it;
            int64_t lo = 0, hi = n_keys;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (keys[mid] < key) lo = mid + 1; else hi = mid;
            }
            if (!(lo < n_keys && keys[lo] == key)) { neg[s] = it; break; }
        }
    }
    return p;                                            // number of draws consumed
}
