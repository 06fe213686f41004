// Row analysis shared by the two tile planners (tiles.cu, segs.cu).  Works on the graph alone:
//   rows   = maximal chains i, i+1, ... of mutually linked consecutive ids (in the reference's
//            numbering these are the raster rows of one layer, R/growthEnvironment.R:86-117);
//   every ordered pair of adjacent rows gets an alignment offset (position of a point's lowest
//   neighbour in the other row minus its own position) and a "pair" flag (some point has two
//   neighbours in the other row: the half-cell shifted rows of the adjacent layers);
//   a breadth-first sweep over the row graph turns the offsets into a column coordinate
//   u = xs[row] + position for every point and a level for every row.
#include <algorithm>

#include "common.h"

namespace eut {

void analyse_rows(const eut_plan *p, RowAnalysis &ra)
{
    const int64_t N = p->N;
    auto has_in = [&](int64_t dst, int32_t src) {
        for (int64_t k = p->row_ptr[dst]; k < p->row_ptr[dst + 1]; k++)
            if (p->col[k] == src) return true;
        return false;
    };
    ra.row_of.assign(N, 0);
    ra.row_start.clear();
    for (int64_t i = 0; i < N; i++) {
        const bool linked = i > 0 && has_in(i, (int32_t)(i - 1)) && has_in(i - 1, (int32_t)i);
        if (!linked) ra.row_start.push_back((int32_t)i);
        ra.row_of[i] = (int32_t)ra.row_start.size() - 1;
    }
    const int32_t R = (int32_t)ra.row_start.size();
    ra.n_rows = R;
    ra.row_start.push_back((int32_t)N);

    ra.adj_ptr.assign(R + 1, 0);
    ra.adj_row.clear(); ra.adj_delta.clear(); ra.adj_pair.clear();
    {
        struct Loc { int32_t row, delta, pair; };
        std::vector<Loc> local;
        for (int32_t r = 0; r < R; r++) {
            local.clear();
            for (int32_t i = ra.row_start[r]; i < ra.row_start[r + 1]; i++) {
                const int32_t ki = i - ra.row_start[r];
                int32_t seen_row[kEllMax];   // rows this point already has a neighbour in
                int n_seen = 0;
                for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                    const int32_t j = p->col[k];
                    const int32_t rj = ra.row_of[j];
                    if (rj == r) continue;
                    const int32_t delta = (j - ra.row_start[rj]) - ki;
                    Loc *hit = nullptr;
                    for (auto &e : local)
                        if (e.row == rj) { hit = &e; break; }
                    if (!hit) { local.push_back({rj, delta, 0}); hit = &local.back(); }
                    else hit->delta = std::min(hit->delta, delta);
                    bool twice = false;
                    for (int s = 0; s < n_seen; s++) twice |= (seen_row[s] == rj);
                    if (twice) hit->pair = 1;
                    else if (n_seen < kEllMax) seen_row[n_seen++] = rj;
                }
            }
            std::sort(local.begin(), local.end(), [](const Loc &a, const Loc &b) { return a.row < b.row; });
            for (auto &e : local) {
                ra.adj_row.push_back(e.row); ra.adj_delta.push_back(e.delta); ra.adj_pair.push_back((uint8_t)e.pair);
            }
            ra.adj_ptr[r + 1] = (int64_t)ra.adj_row.size();
        }
    }
    ra.level.assign(R, -1);
    ra.xs.assign(R, 0);
    ra.xmin = 0;
    std::vector<int32_t> queue;
    queue.reserve(R);
    int32_t level_base = 0;   // components are stacked along the level axis
    for (int32_t seed = 0; seed < R; seed++) {
        if (ra.level[seed] >= 0) continue;
        size_t head = queue.size();
        ra.level[seed] = level_base; ra.xs[seed] = 0;
        queue.push_back(seed);
        int32_t maxl = level_base;
        while (head < queue.size()) {
            const int32_t r = queue[head++];
            maxl = std::max(maxl, ra.level[r]);
            ra.xmin = std::min(ra.xmin, ra.xs[r]);
            for (int64_t a = ra.adj_ptr[r]; a < ra.adj_ptr[r + 1]; a++) {
                const int32_t rr = ra.adj_row[a];
                if (ra.level[rr] >= 0) continue;
                ra.level[rr] = ra.level[r] + 1;
                ra.xs[rr] = ra.xs[r] - ra.adj_delta[a];   // neighbour at k_other = k_own + delta shares u
                queue.push_back(rr);
            }
        }
        level_base = maxl + 2;
    }
}

}  // namespace eut
