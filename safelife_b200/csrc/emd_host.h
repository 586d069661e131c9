// emd_host.h -- exact earth mover's distance between two small histograms on
// the host (the last stage of side_effects.side_effect_score).
//
// The reference calls pyemd.emd(a, b, dist, extra_mass_penalty)
// (side_effects.py:57; pyemd 0.5.1, an un-vendored third-party wrapper of
// Pele & Werman's FastEMD).  Its published definition (Pele & Werman, ICCV
// 2009, "EMD-hat") is
//
//     emd_hat(a, b) = min_F  sum_ij F_ij * dist_ij  +  penalty * |sum(a) - sum(b)|
//     s.t.  F >= 0,  sum_j F_ij <= a_i,  sum_i F_ij <= b_j,
//           sum_ij F_ij = min(sum(a), sum(b))
//
// with penalty = max(dist) when extra_mass_penalty < 0.  This file solves that
// transportation problem exactly in double precision (successive shortest
// augmenting paths with node potentials on the dense bipartite graph, the
// surplus side drained into a zero-cost dummy node).  FastEMD itself rounds
// masses and distances to 1e-6 of their maxima before an integer min-cost
// flow, so the two agree to about 1e-6 relative, not bit for bit: parity at
// this boundary is unpinned.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

namespace sl {

// a[k], b[k] >= 0, dist[k*k] row-major (dist[i*k+j] = cost of moving a_i to b_j).
inline double emd_hat(const double *a, const double *b, const double *dist, int k,
                      double extra_mass_penalty) {
    std::vector<int> src, dst;
    double sum_a = 0.0, sum_b = 0.0, max_d = 0.0;
    for (int i = 0; i < k; i++) {
        if (a[i] > 0.0) { src.push_back(i); sum_a += a[i]; }
        if (b[i] > 0.0) { dst.push_back(i); sum_b += b[i]; }
    }
    for (int i = 0; i < k * k; i++) max_d = std::max(max_d, dist[i]);
    if (extra_mass_penalty < 0.0) extra_mass_penalty = max_d;
    const double extra = std::fabs(sum_a - sum_b) * extra_mass_penalty;
    if (src.empty() || dst.empty()) return extra;

    // Balanced problem: the heavier side ships its surplus to a dummy node at
    // zero cost.  Rows = suppliers, columns = consumers.
    std::vector<double> supply, demand;
    for (int i : src) supply.push_back(a[i]);
    for (int j : dst) demand.push_back(b[j]);
    const int real_rows = (int)src.size(), real_cols = (int)dst.size();
    const double tol = 1e-12 * std::max(sum_a, sum_b);
    if (sum_a > sum_b + tol) demand.push_back(sum_a - sum_b);
    else if (sum_b > sum_a + tol) supply.push_back(sum_b - sum_a);
    const int R = (int)supply.size(), C = (int)demand.size();
    auto cost = [&](int r, int c) -> double {
        return (r < real_rows && c < real_cols) ? dist[src[r] * k + dst[c]] : 0.0;
    };

    std::vector<double> flow((size_t)R * C, 0.0), pr(R, 0.0), pc(C, 0.0);
    std::vector<double> dr(R), dc(C);
    std::vector<int> from_r(R), from_c(C);        // predecessor column of a row / row of a column
    std::vector<char> done_r(R), done_c(C);
    const double INF = std::numeric_limits<double>::infinity();
    double total = 0.0;

    for (int s = 0; s < R; s++) {
        while (supply[s] > tol) {
            // Dijkstra from row s over reduced costs cost - pr[r] + pc[c] >= 0
            std::fill(dr.begin(), dr.end(), INF);
            std::fill(dc.begin(), dc.end(), INF);
            std::fill(done_r.begin(), done_r.end(), 0);
            std::fill(done_c.begin(), done_c.end(), 0);
            dr[s] = 0.0;
            from_r[s] = -1;
            int target = -1;
            for (;;) {
                int best = -1;
                bool is_row = false;
                double bd = INF;
                for (int r = 0; r < R; r++)
                    if (!done_r[r] && dr[r] < bd) { bd = dr[r]; best = r; is_row = true; }
                for (int c = 0; c < C; c++)
                    if (!done_c[c] && dc[c] < bd) { bd = dc[c]; best = c; is_row = false; }
                if (best < 0) break;
                if (is_row) {
                    done_r[best] = 1;
                    for (int c = 0; c < C; c++) {
                        if (done_c[c]) continue;
                        double w = std::max(0.0, cost(best, c) - pr[best] + pc[c]);
                        if (bd + w < dc[c]) { dc[c] = bd + w; from_c[c] = best; }
                    }
                } else {
                    done_c[best] = 1;
                    if (demand[best] > tol) { target = best; break; }
                    for (int r = 0; r < R; r++) {      // backward edges carry existing flow
                        if (done_r[r] || flow[(size_t)r * C + best] <= tol) continue;
                        double w = std::max(0.0, -(cost(r, best) - pr[r] + pc[best]));
                        if (bd + w < dr[r]) { dr[r] = bd + w; from_r[r] = best; }
                    }
                }
            }
            if (target < 0) { supply[s] = 0.0; break; }     // numerically exhausted
            const double dt = dc[target];
            for (int r = 0; r < R; r++) if (done_r[r]) pr[r] += dt - dr[r];
            for (int c = 0; c < C; c++) if (done_c[c]) pc[c] += dt - dc[c];
            // bottleneck along the path target -> ... -> s
            double push = std::min(supply[s], demand[target]);
            for (int c = target;;) {
                int r = from_c[c];
                if (r == s) break;
                int pcol = from_r[r];
                push = std::min(push, flow[(size_t)r * C + pcol]);
                c = pcol;
            }
            for (int c = target;;) {
                int r = from_c[c];
                flow[(size_t)r * C + c] += push;
                total += push * cost(r, c);
                if (r == s) break;
                int pcol = from_r[r];
                flow[(size_t)r * C + pcol] -= push;
                total -= push * cost(r, pcol);
                c = pcol;
            }
            supply[s] -= push;
            demand[target] -= push;
        }
    }
    return total + extra;
}

}  // namespace sl
