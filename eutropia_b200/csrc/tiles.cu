// Tile planner: derives, from the neighbour graph alone, an internal point ordering made of compact
// tiles, and for every tile the shared-memory image the tiled substep kernel works from.
//
// Why: in the reference's numbering (layer-major raster, R/growthEnvironment.R:86-117) the twelve
// in-neighbours of a point live in seven different address runs, up to a whole layer apart.  A CTA
// that takes "the next 256 ids" therefore pulls ~7x its own bytes through L2->L1 (measured:
// profiles/r01_v1_ell_substep_ncu.csv).  A tile that is compact in all three lattice directions
// needs its own points once plus a thin halo.
//
// How (no geometry is used -- mat.in carries none; a graph that is not lattice-like just produces
// poor tiles and the caller falls back to the ELL kernel):
//   1. rows   = maximal chains i, i+1, ... of mutually linked consecutive ids (raster rows);
//   2. every pair of adjacent rows gets an alignment offset (position of a point's neighbour in the
//      other row minus its own position); a breadth-first sweep over the row graph turns these
//      into a column coordinate u for every point and a level for every row;
//   3. tile  = (band of `levels_per_band` consecutive levels) x (block of `tile_cols` columns);
//      points are renumbered tile by tile, row by row inside a tile;
//   4. per tile: halo = in-neighbours outside the tile, grouped into runs of consecutive internal
//      ids; the shared-memory image is [own points | halo runs | zero cell]; a copy list (16-byte
//      and 8-byte cp.async granules) and 16-bit local neighbour offsets are emitted.
#include <algorithm>
#include <numeric>

#include "common.h"

namespace eut {

namespace {
struct Piece {           // a run of one row inside one tile column
    int32_t band, ucol, level, row;
    int32_t first;       // reference id of the first point
    int32_t count;
};
}  // namespace

int build_tiles(eut_plan *p, const TileParams &tp)
{
    const int64_t N = p->N;
    TilePlanHost &T = p->tiles;
    T = TilePlanHost();
    if (N == 0) return EUT_ERR_UNSUPPORTED;

    // ---- 1. rows ------------------------------------------------------------------------------
    auto has_in = [&](int64_t dst, int32_t src) {
        for (int64_t k = p->row_ptr[dst]; k < p->row_ptr[dst + 1]; k++)
            if (p->col[k] == src) return true;
        return false;
    };
    std::vector<int32_t> row_of(N);
    std::vector<int32_t> row_start;
    for (int64_t i = 0; i < N; i++) {
        const bool linked = i > 0 && has_in(i, (int32_t)(i - 1)) && has_in(i - 1, (int32_t)i);
        if (!linked) row_start.push_back((int32_t)i);
        row_of[i] = (int32_t)row_start.size() - 1;
    }
    const int32_t R = (int32_t)row_start.size();
    row_start.push_back((int32_t)N);

    // ---- 2. row adjacency with alignment offsets, BFS -> level, column origin ------------------
    // adjacency in CSR form: for each row the distinct neighbour rows and min(k_other - k_own)
    std::vector<int64_t> adj_ptr(R + 1, 0);
    std::vector<int32_t> adj_row, adj_delta;
    {
        std::vector<std::pair<int32_t, int32_t>> local;
        for (int32_t r = 0; r < R; r++) {
            local.clear();
            for (int32_t i = row_start[r]; i < row_start[r + 1]; i++) {
                const int32_t ki = i - row_start[r];
                for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                    const int32_t j = p->col[k];
                    const int32_t rj = row_of[j];
                    if (rj == r) continue;
                    const int32_t delta = (j - row_start[rj]) - ki;
                    bool found = false;
                    for (auto &e : local)
                        if (e.first == rj) { e.second = std::min(e.second, delta); found = true; break; }
                    if (!found) local.emplace_back(rj, delta);
                }
            }
            std::sort(local.begin(), local.end());
            for (auto &e : local) { adj_row.push_back(e.first); adj_delta.push_back(e.second); }
            adj_ptr[r + 1] = (int64_t)adj_row.size();
        }
    }
    std::vector<int32_t> level(R, -1), xs(R, 0), comp(R, -1);
    std::vector<int32_t> comp_rows, comp_maxlevel;
    {
        std::vector<int32_t> queue;
        queue.reserve(R);
        for (int32_t seed = 0; seed < R; seed++) {
            if (level[seed] >= 0) continue;
            const int32_t cid = (int32_t)comp_rows.size();
            comp_rows.push_back(0);
            comp_maxlevel.push_back(0);
            size_t head = queue.size();
            level[seed] = 0; xs[seed] = 0; comp[seed] = cid;
            queue.push_back(seed);
            while (head < queue.size()) {
                const int32_t r = queue[head++];
                comp_rows[cid]++;
                comp_maxlevel[cid] = std::max(comp_maxlevel[cid], level[r]);
                for (int64_t a = adj_ptr[r]; a < adj_ptr[r + 1]; a++) {
                    const int32_t rr = adj_row[a];
                    if (level[rr] >= 0) continue;
                    level[rr] = level[r] + 1;
                    xs[rr] = xs[r] - adj_delta[a];   // neighbour at k_other = k_own + delta shares u
                    comp[rr] = cid;
                    queue.push_back(rr);
                }
            }
        }
    }
    // normalise u >= 0 per component; bands per component, stacked
    const int n_comp = (int)comp_rows.size();
    std::vector<int32_t> comp_xmin(n_comp, INT32_MAX), comp_band0(n_comp, 0), comp_lpb(n_comp, 1);
    for (int32_t r = 0; r < R; r++) comp_xmin[comp[r]] = std::min(comp_xmin[comp[r]], xs[r]);
    {
        int32_t band_base = 0;
        for (int c = 0; c < n_comp; c++) {
            const double rows_per_level = (double)comp_rows[c] / (comp_maxlevel[c] + 1);
            int lpb = tp.levels_per_band > 0
                          ? tp.levels_per_band
                          : (int)std::max(1.0, std::floor(tp.target_rows / std::max(1.0, rows_per_level) + 0.5));
            comp_lpb[c] = lpb;
            comp_band0[c] = band_base;
            band_base += comp_maxlevel[c] / lpb + 1;
        }
    }

    // ---- 3. pieces -> tiles -> permutation ------------------------------------------------------
    const int X = tp.tile_cols;
    std::vector<Piece> pieces;
    pieces.reserve((size_t)(N / std::max(1, X / 2)) + R);
    for (int32_t r = 0; r < R; r++) {
        const int c = comp[r];
        const int32_t u0 = xs[r] - comp_xmin[c];
        const int32_t len = row_start[r + 1] - row_start[r];
        const int32_t band = comp_band0[c] + level[r] / comp_lpb[c];
        int32_t k = 0;
        while (k < len) {
            const int32_t ucol = (u0 + k) / X;
            const int32_t kend = std::min<int32_t>(len, (ucol + 1) * X - u0);
            pieces.push_back({band, ucol, level[r], r, row_start[r] + k, kend - k});
            k = kend;
        }
    }
    std::sort(pieces.begin(), pieces.end(), [](const Piece &a, const Piece &b) {
        if (a.band != b.band) return a.band < b.band;
        if (a.ucol != b.ucol) return a.ucol < b.ucol;
        if (a.level != b.level) return a.level < b.level;
        return a.first < b.first;
    });
    p->perm.assign(N, 0);
    T.tile_start.clear();
    {
        int32_t pos = 0, cur_band = -1, cur_ucol = -1, in_tile = 0;
        for (const Piece &pc : pieces) {
            int32_t done = 0;
            while (done < pc.count) {
                const bool new_tile = pc.band != cur_band || pc.ucol != cur_ucol || in_tile >= tp.max_tile_points;
                if (new_tile) {
                    T.tile_start.push_back(pos);
                    cur_band = pc.band; cur_ucol = pc.ucol; in_tile = 0;
                }
                const int32_t take = std::min(pc.count - done, tp.max_tile_points - in_tile);
                for (int32_t k = 0; k < take; k++) p->perm[pc.first + done + k] = pos + k;
                pos += take; done += take; in_tile += take;
            }
        }
        T.tile_start.push_back(pos);
    }
    const int32_t n_tiles = (int32_t)T.tile_start.size() - 1;
    p->Np = ((N + kPadPoints - 1) / kPadPoints) * kPadPoints;
    p->iperm.assign(p->Np, -1);
    for (int64_t i = 0; i < N; i++) p->iperm[p->perm[i]] = (int32_t)i;

    // ---- 4. per-tile shared-memory image, copy list, local offsets -----------------------------
    const int W = kEllMax;   // the tiled kernel is compiled for 12 neighbour slots
    T.loc.assign((size_t)W * p->Np, 0);
    T.copy_ptr.assign(n_tiles + 1, 0);
    T.copies.clear();
    T.img_points = 0; T.max_tile_points = 0; T.max_copies = 0;
    int64_t halo_total = 0;
    std::vector<int32_t> halo;
    auto emit = [&](int32_t g, int32_t s, int32_t len) {
        // granules: 16 bytes where the global (and, by construction, the shared) offset is even
        int32_t k = 0;
        if (len > 0 && ((g + k) & 1)) { T.copies.push_back({g + k, (s + k) | (1 << 16)}); k++; }
        for (; k + 2 <= len; k += 2) T.copies.push_back({g + k, (s + k) | (2 << 16)});
        if (k < len) T.copies.push_back({g + k, (s + k) | (1 << 16)});
    };
    for (int32_t t = 0; t < n_tiles; t++) {
        const int32_t t0 = T.tile_start[t], t1 = T.tile_start[t + 1];
        halo.clear();
        for (int32_t q = t0; q < t1; q++) {
            const int32_t i = p->iperm[q];
            for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                const int32_t nq = p->perm[p->col[k]];
                if (nq < t0 || nq >= t1) halo.push_back(nq);
            }
        }
        std::sort(halo.begin(), halo.end());
        halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
        // image: cells 0 and 1 stay zero (target of unused neighbour slots); own points follow at
        // an offset whose parity equals the global one so that 16-byte granules line up
        const int32_t own_off = 2 + (t0 & 1);
        emit(t0, own_off, t1 - t0);
        int32_t next = own_off + (t1 - t0);
        std::vector<int32_t> halo_off(halo.size());
        for (size_t h = 0; h < halo.size();) {
            size_t e = h + 1;
            while (e < halo.size() && halo[e] == halo[e - 1] + 1) e++;
            if ((next & 1) != (halo[h] & 1)) next++;
            for (size_t k = h; k < e; k++) halo_off[k] = next + (int32_t)(k - h);
            emit(halo[h], next, (int32_t)(e - h));
            next += (int32_t)(e - h);
            h = e;
        }
        const int32_t zero_cell = 0;      // never written by a copy, zeroed once by the kernel
        if (next > tp.max_img_points || next * 8 > 65535) return EUT_ERR_UNSUPPORTED;
        T.img_points = std::max(T.img_points, next);
        T.max_tile_points = std::max(T.max_tile_points, t1 - t0);
        T.copy_ptr[t + 1] = (int32_t)T.copies.size();
        T.max_copies = std::max(T.max_copies, T.copy_ptr[t + 1] - T.copy_ptr[t]);
        halo_total += (int64_t)halo.size();
        // local offsets (in points) of every in-neighbour, mat.in order; unused slots -> zero cell
        for (int32_t q = t0; q < t1; q++) {
            const int32_t i = p->iperm[q];
            const int d = (int)(p->row_ptr[i + 1] - p->row_ptr[i]);
            for (int k = 0; k < W; k++) {
                int32_t off = zero_cell;
                if (k < d) {
                    const int32_t nq = p->perm[p->col[p->row_ptr[i] + k]];
                    if (nq >= t0 && nq < t1) off = own_off + (nq - t0);
                    else off = halo_off[std::lower_bound(halo.begin(), halo.end(), nq) - halo.begin()];
                }
                T.loc[(size_t)k * p->Np + q] = (uint16_t)(off * 8);   // byte offset inside the image
            }
        }
    }
    T.halo_ratio = (double)halo_total / (double)N;
    if (T.max_copies > tp.max_copies) return EUT_ERR_UNSUPPORTED;
    p->n_tiles = n_tiles;
    return EUT_OK;
}

}  // namespace eut
