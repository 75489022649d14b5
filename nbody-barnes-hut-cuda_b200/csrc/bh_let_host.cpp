// bh_let_host.cpp — host-side decisions of the locally-essential-tree mode (include/bh.h): how a
// Morton-key range is cut into octree-aligned intervals, and how the ranks' key samples become new key ranges.
// Pure host code: every rank evaluates these on all-gathered data and must arrive at the same answer, so there
// is no floating-point freedom here (sequential double sums, stable sort).  The reference has no multi-GPU path.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <set>
#include <vector>

#include "../../include/bh.h"

namespace {
constexpr int64_t KEY_END = (int64_t)1 << 30;   // one past the largest 30-bit key

inline int64_t ceil_to(int64_t v, int64_t step) { return (v + step - 1) / step * step; }
}  // namespace

extern "C" {

// Interior: the 8..64 whole cells of size S = 8^j (largest with span/S >= 8) inside [k_lo, k_hi); the two ragged
// ends (parts of one S-cell each) are cut again at S/64.  Whole-cell intervals are convex and owned by one rank.
int bh_let_domain_cuts(uint32_t k_lo32, uint32_t k_hi32, int max_boxes, uint32_t* cuts) {
    const int64_t k_lo = k_lo32, k_hi = k_hi32;
    if (!cuts || max_boxes < 1 || k_lo > k_hi || k_hi > KEY_END) return BH_E_INVAL;
    std::set<int64_t> c = {k_lo, k_hi};
    const int64_t span = k_hi - k_lo;
    if (span > 0) {
        int64_t S = 1;
        while (S * 64 <= span) S *= 8;
        const int64_t first = ceil_to(k_lo, S), last = k_hi / S * S;
        const int64_t fine = std::max<int64_t>(S / 64, 1);
        if (first <= last) {
            for (int64_t v = first; v <= last; v += S) c.insert(v);
            const int64_t ends[2][2] = {{k_lo, first}, {last, k_hi}};
            for (auto& e : ends)
                if (e[1] - e[0] > fine)
                    for (int64_t v = ceil_to(e[0], fine); v < e[1]; v += fine) c.insert(v);
        } else {   // the whole range lies inside one S-cell
            for (int64_t v = ceil_to(k_lo, fine); v < k_hi; v += fine) c.insert(v);
        }
    }
    std::vector<int64_t> out;
    for (int64_t v : c)
        if (v >= k_lo && v <= k_hi) out.push_back(v);
    if ((int)out.size() > max_boxes + 1) return BH_E_INVAL;
    for (int i = 0; i <= max_boxes; ++i) cuts[i] = (uint32_t)(i < (int)out.size() ? out[i] : k_hi);   // empty intervals pad
    return 0;
}

// samples: world x nsample keys, row r = keys of rank r at equal increments of its cumulative work, so that every
// sample stands for work[r]/nsample (rows with work <= 0 are ignored).  edges[r] = the key at r/world of the pooled
// work; edges[0] = 0, edges[world] = 2^30, non-decreasing.
int bh_let_elect_splitters(const int64_t* samples, int nsample, const double* work, int world, int64_t* edges) {
    if (!samples || !work || !edges || nsample < 1 || world < 1) return BH_E_INVAL;
    edges[0] = 0;
    bool any = false;
    for (int r = 0; r < world; ++r) any |= work[r] > 0.0;
    if (!any) {
        for (int r = 1; r <= world; ++r) edges[r] = KEY_END;
        return 0;
    }
    const size_t m = (size_t)world * nsample;
    std::vector<size_t> order(m);
    std::iota(order.begin(), order.end(), (size_t)0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return samples[a] < samples[b]; });
    std::vector<double> cum(m);
    double acc = 0.0;
    for (size_t i = 0; i < m; ++i) {
        const int r = (int)(order[i] / nsample);
        acc += work[r] > 0.0 ? work[r] / nsample : 0.0;
        cum[i] = acc;
    }
    for (int r = 1; r < world; ++r) {
        const double target = cum[m - 1] * r / world;
        size_t idx = (size_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        if (idx > m - 1) idx = m - 1;
        edges[r] = std::max(samples[order[idx]], edges[r - 1]);
    }
    edges[world] = KEY_END;
    return 0;
}

}  // extern "C"
