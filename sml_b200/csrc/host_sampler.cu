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

// ---- hashed variant ---------------------------------------------------------------------------------
// The binary search above costs ~17 dependent cache misses per draw (15 ms per 75 000-sample epoch on the host).  The
// same membership test through an open-addressing hash set of the period's (user * span + item) keys takes one or two
// probes.  table: int64[table_size], table_size a power of two >= 2 * n, filled here (-1 = empty).
namespace {
inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
}  // namespace

extern "C" int sml_host_keyset_build(const int64_t *users, const int64_t *items, int64_t n, int64_t span, int64_t *table,
                                     int64_t table_size) {
    if (n < 0 || table_size < 2 || (table_size & (table_size - 1)) != 0 || table_size < 2 * n || (n > 0 && (!users || !items)) || !table)
        return SML_E_BADARG;
    const uint64_t mask = (uint64_t)table_size - 1;
    for (int64_t i = 0; i < table_size; ++i) table[i] = -1;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t key = users[i] * span + items[i];
        uint64_t h = mix64((uint64_t)key) & mask;
        while (table[h] != -1 && table[h] != key) h = (h + 1) & mask;
        table[h] = key;
    }
    return SML_OK;
}

// Resumable walk: starts at sample *sample_io, stops when the samples or the draws run out, leaves the next sample in
// *sample_io and returns the number of draws consumed.  Called with exactly as many draws as samples remain it always
// consumes all of them (every sample needs at least one), so the caller never has to rewind the generator.
extern "C" int64_t sml_host_rejection_walk_hashed(const int64_t *draws, int64_t n_draws, const int64_t *users, int64_t n,
                                                  int64_t *sample_io, const int64_t *item_all, const int64_t *table,
                                                  int64_t table_size, int64_t span, int64_t *neg) {
    const uint64_t mask = (uint64_t)table_size - 1;
    int64_t p = 0, s = *sample_io;
    while (s < n && p < n_draws) {
        const int64_t it = item_all[draws[p++]];
        const int64_t key = users[s] * span + it;
        uint64_t h = mix64((uint64_t)key) & mask;
        bool hit = false;
        while (table[h] != -1) {
            if (table[h] == key) { hit = true; break; }
            h = (h + 1) & mask;
        }
        if (!hit) neg[s++] = it;
    }
    *sample_io = s;
    return p;
}
