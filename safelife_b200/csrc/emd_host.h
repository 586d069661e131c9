// emd_host.h -- exact earth mover's distance between two small histograms on
// the host (the last stage of side_effects.side_effect_score).
//
// The reference calls pyemd.emd(a, b, dist, extra_mass_penalty)
// (side_effects.py:57; pyemd 0.5.1, an un-vendored third-party wrapper of
// Pele & Werman's FastEMD, entry point emd_hat_gd_metric).  Its published
// definition (Pele & Werman, ICCV 2009, "EMD-hat") is
//
//     emd_hat(a, b) = min_F  sum_ij F_ij * dist_ij  +  penalty * |sum(a) - sum(b)|
//     s.t.  F >= 0,  sum_j F_ij <= a_i,  sum_i F_ij <= b_j,
//           sum_ij F_ij = min(sum(a), sum(b))
//
// with penalty = max(dist) when extra_mass_penalty < 0, for a ground distance
// that is a metric: dist_ii = 0 and the (directed) triangle inequality hold, so
// the mass common to a_i and b_i stays where it is.  The reference's ground
// distance (tanh of a wrapped manhattan distance, side_effects.py:45-56) is such
// a (quasi-)metric, including its signed-wrap asymmetry.
//
// This file solves that transportation problem exactly in double precision:
// common mass is pre-flowed (as FastEMD's gd_metric entry does), every remaining
// point is a pure supplier or a pure consumer, the surplus side drains into a
// zero-cost dummy, and the rest is successive shortest augmenting paths with
// node potentials on the dense bipartite graph (Dijkstra stops at the nearest
// unsaturated consumer).  FastEMD itself rounds masses and distances to 1e-6 of
// their maxima before an integer min-cost flow, so the two agree to about 1e-6
// relative, not bit for bit: parity at this boundary is unpinned (DESIGN.md
// section 4).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <thread>
#include <vector>

namespace sl {

// a[k], b[k] >= 0, dist[k*k] row-major (dist[i*k+j] = cost of moving a_i to b_j).
inline double emd_hat(const double *a, const double *b, const double *dist, int k,
                      double extra_mass_penalty) {
    double sum_a = 0.0, sum_b = 0.0, max_d = 0.0;
    for (int i = 0; i < k; i++) {
        sum_a += a[i];
        sum_b += b[i];
    }
    for (long long i = 0; i < (long long)k * k; i++) max_d = std::max(max_d, dist[i]);
    if (extra_mass_penalty < 0.0) extra_mass_penalty = max_d;
    const double extra = std::fabs(sum_a - sum_b) * extra_mass_penalty;
    const double scale = std::max(sum_a, sum_b);
    if (!(scale > 0.0)) return extra;
    const double tol = 1e-13 * scale;

    // pre-flow: what a_i and b_i have in common does not move
    std::vector<int> src, dst;
    std::vector<double> supply, demand;
    double left_a = 0.0, left_b = 0.0;
    for (int i = 0; i < k; i++) {
        const double d = a[i] - b[i];
        if (d > tol) {
            src.push_back(i);
            supply.push_back(d);
            left_a += d;
        } else if (-d > tol) {
            dst.push_back(i);
            demand.push_back(-d);
            left_b += -d;
        }
    }
    if (src.empty() || dst.empty()) return extra;

    // Balanced problem: the heavier side ships its surplus to a dummy at zero cost.
    const int real_rows = (int)src.size(), real_cols = (int)dst.size();
    if (left_a > left_b + tol) demand.push_back(left_a - left_b);
    else if (left_b > left_a + tol) supply.push_back(left_b - left_a);
    const int R = (int)supply.size(), C = (int)demand.size();
    std::vector<double> cost((size_t)R * C, 0.0);
    for (int r = 0; r < real_rows; r++)
        for (int c = 0; c < real_cols; c++) cost[(size_t)r * C + c] = dist[(size_t)src[r] * k + dst[c]];

    std::vector<double> flow((size_t)R * C, 0.0), pr(R, 0.0), pc(C, 0.0), dr(R), dc(C);
    std::vector<int> from_r(R), from_c(C);        // predecessor column of a row / row of a column
    std::vector<int> open_r, open_c;               // nodes not yet finalised by the current search
    std::vector<int> seen_r, seen_c;               // nodes finalised by the current search
    const double INF = std::numeric_limits<double>::infinity();
    double total = 0.0;

    for (int s = 0; s < R; s++) {
        while (supply[s] > tol) {
            // Dijkstra from row s over reduced costs cost - pr[r] + pc[c] >= 0
            std::fill(dr.begin(), dr.end(), INF);
            std::fill(dc.begin(), dc.end(), INF);
            open_r.clear();
            open_c.clear();
            seen_r.clear();
            seen_c.clear();
            for (int r = 0; r < R; r++) if (r != s) open_r.push_back(r);
            for (int c = 0; c < C; c++) open_c.push_back(c);
            dr[s] = 0.0;
            from_r[s] = -1;
            int target = -1;
            int cur = s;
            bool cur_is_row = true;
            double cur_d = 0.0;
            for (;;) {
                if (cur_is_row) {
                    seen_r.push_back(cur);
                    const double *crow = &cost[(size_t)cur * C];
                    const double base = cur_d - pr[cur];
                    for (int c : open_c) {
                        const double nd = std::max(cur_d, base + crow[c] + pc[c]);
                        if (nd < dc[c]) {
                            dc[c] = nd;
                            from_c[c] = cur;
                        }
                    }
                } else {
                    seen_c.push_back(cur);
                    if (demand[cur] > tol) {
                        target = cur;
                        break;
                    }
                    for (int r : open_r) {      // backward edges carry existing flow
                        if (flow[(size_t)r * C + cur] <= tol) continue;
                        const double nd = std::max(cur_d, cur_d - (cost[(size_t)r * C + cur] - pr[r] + pc[cur]));
                        if (nd < dr[r]) {
                            dr[r] = nd;
                            from_r[r] = cur;
                        }
                    }
                }
                // next: the closest open node
                double best = INF;
                int bi = -1;
                bool brow = false;
                for (size_t i = 0; i < open_c.size(); i++)
                    if (dc[open_c[i]] < best) {
                        best = dc[open_c[i]];
                        bi = (int)i;
                    }
                for (size_t i = 0; i < open_r.size(); i++)
                    if (dr[open_r[i]] < best) {
                        best = dr[open_r[i]];
                        bi = (int)i;
                        brow = true;
                    }
                if (bi < 0) break;
                cur_is_row = brow;
                cur_d = best;
                std::vector<int> &open = brow ? open_r : open_c;
                cur = open[bi];
                open[bi] = open.back();
                open.pop_back();
            }
            if (target < 0) {     // numerically exhausted
                supply[s] = 0.0;
                break;
            }
            const double dt = dc[target];
            for (int r : seen_r) pr[r] += dt - dr[r];
            for (int c : seen_c) pc[c] += dt - dc[c];
            // bottleneck along the path target -> ... -> s
            double push = std::min(supply[s], demand[target]);
            for (int c = target;;) {
                const int r = from_c[c];
                if (r == s) break;
                const int pcol = from_r[r];
                push = std::min(push, flow[(size_t)r * C + pcol]);
                c = pcol;
            }
            for (int c = target;;) {
                const int r = from_c[c];
                flow[(size_t)r * C + c] += push;
                total += push * cost[(size_t)r * C + c];
                if (r == s) break;
                const int pcol = from_r[r];
                flow[(size_t)r * C + pcol] -= push;
                total -= push * cost[(size_t)r * C + pcol];
                c = pcol;
            }
            supply[s] -= push;
            demand[target] -= push;
        }
    }
    return total + extra;
}

// n independent problems on up to n_threads host threads (n_threads <= 0: one per
// hardware thread).  k[i] points each; a[i], b[i], dist[i] as for emd_hat.
inline void emd_hat_batch(const double *const *a, const double *const *b,
                          const double *const *dist, const int *k, int n,
                          double extra_mass_penalty, double *out, int n_threads) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    n_threads = std::max(1, std::min(n_threads, n));
    auto work = [&](int t) {
        for (int i = t; i < n; i += n_threads)
            out[i] = k[i] == 0 ? 0.0 : emd_hat(a[i], b[i], dist[i], k[i], extra_mass_penalty);
    };
    if (n_threads == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(work, t);
    for (std::thread &th : pool) th.join();
}

}  // namespace sl
