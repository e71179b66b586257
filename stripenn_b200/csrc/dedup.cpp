// getStripe.RemoveRedundant (getStripe.py:1116-1196) — host code, no device work.
// Every unordered pair of rows of one chromosome whose window ids differ by at most one is
// tested once (earlier row first); a pair overlaps when the shared fraction of both the x and
// the y extent exceeds 0.2; the loser is marked, marks accumulate (non-greedy).
//   size  : smaller h/w loses, tie -> the earlier row
//   pvalue: larger p loses,    tie -> the later row
//   score : smaller Stripiness loses, tie -> the earlier row
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/stripenn_b200.h"

extern "C" int snn_remove_redundant(int64_t n, const int32_t* group, const int64_t* num, const int64_t* pos,
                                    const double* h, const double* w, const double* key, int by,
                                    uint8_t* keep) {
    if (n < 0 || by < 0 || by > 2) return SNN_EINVAL;
    if (n == 0) return SNN_OK;
    if (!group || !num || !pos || !keep || (by == 0 && (!h || !w)) || (by != 0 && !key)) return SNN_EINVAL;
    for (int64_t i = 0; i < n; ++i) keep[i] = 1;
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        if (group[a] != group[b]) return group[a] < group[b];
        return num[a] < num[b];
    });
    std::vector<double> ratio;
    if (by == 0) {
        ratio.resize(n);
        for (int64_t i = 0; i < n; ++i) ratio[i] = h[i] / w[i];
    }
    // lo[a]: first position in `order` with the same group and num >= num[a]-1
    int64_t lo = 0;
    for (int64_t a = 0; a < n; ++a) {
        const int64_t i = order[a];
        while (lo < a && (group[order[lo]] != group[i] || num[order[lo]] < num[i] - 1)) ++lo;
        const int64_t p1 = pos[i * 4], p2 = pos[i * 4 + 1], p3 = pos[i * 4 + 2], p4 = pos[i * 4 + 3];
        for (int64_t b = lo; b < n; ++b) {
            const int64_t j = order[b];
            if (group[j] != group[i] || num[j] > num[i] + 1) break;
            if (j <= i) continue;  // each unordered pair once, i = the earlier row
            const int64_t q1 = pos[j * 4], q2 = pos[j * 4 + 1], q3 = pos[j * 4 + 2], q4 = pos[j * 4 + 3];
            int64_t ix = std::min(p2, q2) - std::max(p1, q1) + 1;
            int64_t iy = std::min(p4, q4) - std::max(p3, q3) + 1;
            if (ix < 0) ix = 0;
            if (iy < 0) iy = 0;
            const double sx = (double)ix / (double)std::min(p2 - p1, q2 - q1);
            const double sy = (double)iy / (double)std::min(p4 - p3, q4 - q3);
            if (!(sx > 0.2 && sy > 0.2)) continue;
            bool first;
            if (by == 0) first = ratio[i] <= ratio[j];
            else if (by == 2) first = key[i] <= key[j];
            else first = key[i] > key[j];
            if (first) keep[i] = 0;
            else keep[j] = 0;
        }
    }
    return SNN_OK;
}
