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
//   3. rows are cut into pieces at multiples of `tile_cols` in u; pieces of one u-column that touch
//      each other form a strip (one per arm of a non-convex domain); every strip is cut along the
//      level axis into tiles of at most `max_tile_points` points;
//   4. points are renumbered tile by tile, level by level and row by row inside a tile;
//   5. per tile: halo = in-neighbours outside the tile, grouped into runs of consecutive internal
//      ids; the shared-memory image is [2 zero cells | own points | halo runs]; emitted are a list
//      of bulk copies (16-byte aligned runs, for cp.async.bulk), a list of single 8-byte copies
//      (odd run ends, lone halo points) and 16-bit image offsets of every neighbour slot.
#include <algorithm>
#include <numeric>

#include "common.h"

namespace eut {

namespace {
struct Piece {           // a run of one row inside one u-column
    int32_t ucol, level;
    int32_t first;       // reference id of the first point
    int32_t count;
    int32_t strip;
};

struct Dsu {
    std::vector<int32_t> p;
    explicit Dsu(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int32_t find(int32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int32_t a, int32_t b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};
}  // namespace

int build_tiles(eut_plan *p, const TileParams &tp)
{
    const int64_t N = p->N;
    TilePlanHost &T = p->tiles;
    T = TilePlanHost();
    if (N == 0) return EUT_ERR_UNSUPPORTED;

    // ---- 1./2. rows, row adjacency, levels and column origins (rows.cu) -------------------------
    RowAnalysis ra;
    analyse_rows(p, ra);
    const std::vector<int32_t> &row_start = ra.row_start, &level = ra.level, &xs = ra.xs;
    const int32_t R = ra.n_rows, xmin = ra.xmin;

    // ---- 3. pieces, strips, tiles -------------------------------------------------------------
    const int X = std::max(2, tp.tile_cols);
    std::vector<Piece> pieces;
    pieces.reserve((size_t)(N / std::max(1, X / 2)) + R);
    std::vector<int32_t> piece_of(N);
    for (int32_t r = 0; r < R; r++) {
        const int32_t u0 = xs[r] - xmin;
        const int32_t len = row_start[r + 1] - row_start[r];
        // cuts at multiples of X in u; a remainder shorter than X/2 at either end of the row is
        // merged into its neighbour piece (no skinny tile columns with more halo than content)
        int32_t k = 0;
        while (k < len) {
            int32_t ucol = (u0 + k) / X;
            int32_t kend = std::min<int32_t>(len, (ucol + 1) * X - u0);
            if (k == 0 && kend < len && kend - k < X / 2) {          // short head: extend over the next column
                ucol += 1;
                kend = std::min<int32_t>(len, (ucol + 1) * X - u0);
            }
            if (len - kend > 0 && len - kend < X / 2) kend = len;    // short tail: absorb it
            for (int32_t i = k; i < kend; i++) piece_of[row_start[r] + i] = (int32_t)pieces.size();
            pieces.push_back({ucol, level[r], row_start[r] + k, kend - k, 0});
            k = kend;
        }
    }
    const int32_t P = (int32_t)pieces.size();
    {
        Dsu dsu(P);
        for (int64_t i = 0; i < N; i++) {
            const int32_t a = piece_of[i];
            for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                const int32_t b = piece_of[p->col[k]];
                if (a != b && pieces[a].ucol == pieces[b].ucol) dsu.unite(a, b);
            }
        }
        for (int32_t a = 0; a < P; a++) pieces[a].strip = dsu.find(a);
    }
    // order pieces: strip by strip (strips by coarse level band, then u-column), level by level
    std::vector<int32_t> strip_minlevel(P, INT32_MAX);
    for (int32_t a = 0; a < P; a++)
        strip_minlevel[pieces[a].strip] = std::min(strip_minlevel[pieces[a].strip], pieces[a].level);
    std::vector<int32_t> ord(P);
    std::iota(ord.begin(), ord.end(), 0);
    const int band = std::max(1, tp.band_levels);
    std::sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) {
        const Piece &A = pieces[a], &B = pieces[b];
        if (A.strip != B.strip) {
            const int32_t ba = strip_minlevel[A.strip] / band, bb = strip_minlevel[B.strip] / band;
            if (ba != bb) return ba < bb;
            if (A.ucol != B.ucol) return A.ucol < B.ucol;
            return A.strip < B.strip;
        }
        if (A.level != B.level) return A.level < B.level;
        return A.first < B.first;
    });
    // cut strips into tiles at level boundaries; a level group that does not fit is split
    p->perm.assign(N, 0);
    T.tile_start.clear();
    {
        int32_t pos = 0, in_tile = 0, cur_strip = -1;
        size_t i = 0;
        while (i < ord.size()) {
            size_t e = i;   // level group [i, e)
            int32_t gcount = 0;
            while (e < ord.size() && pieces[ord[e]].strip == pieces[ord[i]].strip &&
                   pieces[ord[e]].level == pieces[ord[i]].level) { gcount += pieces[ord[e]].count; e++; }
            if (pieces[ord[i]].strip != cur_strip || in_tile + gcount > tp.max_tile_points) {
                T.tile_start.push_back(pos);
                in_tile = 0;
                cur_strip = pieces[ord[i]].strip;
            }
            for (size_t k = i; k < e; k++) {
                const Piece &pc = pieces[ord[k]];
                int32_t done = 0;
                while (done < pc.count) {
                    if (in_tile >= tp.max_tile_points) { T.tile_start.push_back(pos); in_tile = 0; }
                    const int32_t take = std::min(pc.count - done, tp.max_tile_points - in_tile);
                    for (int32_t q = 0; q < take; q++) p->perm[pc.first + done + q] = pos + q;
                    pos += take; done += take; in_tile += take;
                }
            }
            i = e;
        }
        T.tile_start.push_back(pos);
        T.tile_start.erase(std::unique(T.tile_start.begin(), T.tile_start.end()), T.tile_start.end());
    }
    const int32_t n_tiles = (int32_t)T.tile_start.size() - 1;
    p->Np = ((N + kPadPoints - 1) / kPadPoints) * kPadPoints;
    p->iperm.assign(p->Np, -1);
    for (int64_t i = 0; i < N; i++) p->iperm[p->perm[i]] = (int32_t)i;

    // ---- 4. per-tile shared-memory image, copy lists, local offsets -----------------------------
    const int W = kEllMax;   // the tiled kernel is compiled for 12 neighbour slots
    T.loc.assign((size_t)W * p->Np, 0);
    T.meta.assign((size_t)n_tiles * 8, 0);
    T.copies.clear();
    T.img_points = 0; T.max_tile_points = 0; T.max_bulk = 0; T.max_single = 0;
    int64_t halo_total = 0;
    std::vector<int32_t> halo, halo_off;
    std::vector<int2> bulk, single;
    auto emit = [&](int32_t g, int32_t s, int32_t len) {
        // [odd head] [16-byte aligned body] [odd tail]; global and image offsets share their parity
        if (len > 0 && (g & 1)) { single.push_back({g, s}); g++; s++; len--; }
        const int32_t body = len & ~1;
        if (body > 0) bulk.push_back({g, s | (body << 16)});
        if (len & 1) single.push_back({g + body, s + body});
    };
    for (int32_t t = 0; t < n_tiles; t++) {
        const int32_t t0 = T.tile_start[t], t1 = T.tile_start[t + 1];
        halo.clear(); bulk.clear(); single.clear();
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
        // an offset whose parity equals the global one so that 16-byte copies line up
        const int32_t own_off = 2 + (t0 & 1);
        emit(t0, own_off, t1 - t0);
        int32_t next = own_off + (t1 - t0);
        halo_off.assign(halo.size(), -1);
        // (a) runs of three or more consecutive ids: contiguous in the image, bulk copied
        for (size_t h = 0; h < halo.size();) {
            size_t e = h + 1;
            while (e < halo.size() && halo[e] == halo[e - 1] + 1) e++;
            if (e - h >= 3) {
                if ((next & 1) != (halo[h] & 1)) next++;
                for (size_t k = h; k < e; k++) halo_off[k] = next + (int32_t)(k - h);
                emit(halo[h], next, (int32_t)(e - h));
                next += (int32_t)(e - h);
            }
            h = e;
        }
        // (b) lone halo points (the x-1 / x+1 and diagonal neighbours at the ends of a row piece):
        // a warp reads such a point together with 31 contiguous cells of a row, so it is placed at
        // an image offset congruent (mod 16 cells = 128 bytes = all banks) to "one before / one
        // after" the cells its warp mates read -- the 32 addresses then cover every bank once and
        // the load is conflict free.  Layout: a 16-column grid after the runs, column = residue.
        {
            auto off_of = [&](int32_t nq) -> int32_t {
                if (nq >= t0 && nq < t1) return own_off + (nq - t0);
                return halo_off[std::lower_bound(halo.begin(), halo.end(), nq) - halo.begin()];
            };
            std::vector<int32_t> want(halo.size(), -1);
            for (int32_t q = t0; q < t1; q++) {
                const int32_t i = p->iperm[q];
                const int d = (int)(p->row_ptr[i + 1] - p->row_ptr[i]);
                for (int k = 0; k < d; k++) {
                    const int32_t nq = p->perm[p->col[p->row_ptr[i] + k]];
                    if (nq >= t0 && nq < t1) continue;
                    const size_t hi_ = std::lower_bound(halo.begin(), halo.end(), nq) - halo.begin();
                    if (halo_off[hi_] >= 0 || want[hi_] >= 0) continue;
                    for (int side = 1; side >= -1 && want[hi_] < 0; side -= 2) {   // warp mate q+1, then q-1
                        const int32_t q2 = q + side;
                        if (q2 < t0 || q2 >= t1) continue;
                        const int32_t i2 = p->iperm[q2];
                        if (i2 != i + side) continue;   // not the x-neighbour in the same raster row
                        if ((int)(p->row_ptr[i2 + 1] - p->row_ptr[i2]) <= k) continue;
                        const int32_t o2 = off_of(p->perm[p->col[p->row_ptr[i2] + k]]);
                        if (o2 >= 0) want[hi_] = ((o2 - side) % 16 + 16) % 16;
                    }
                }
            }
            next = (next + 15) / 16 * 16;
            int32_t fill[16] = {0};
            int32_t rr = 0;
            for (size_t h = 0; h < halo.size(); h++) {
                if (halo_off[h] >= 0) continue;
                const int32_t r = want[h] >= 0 ? want[h] : (rr++ & 15);
                halo_off[h] = next + fill[r] * 16 + r;
                fill[r]++;
                single.push_back({halo[h], halo_off[h]});
            }
            next += 16 * *std::max_element(fill, fill + 16);
        }
        if (next > tp.max_img_points || next * 8 > 65535) {
            if (getenv("EUT_TRACE"))
                fprintf(stderr, "[eut] tile planner: tile %d (%d points, %zu halo) needs an image of %d points > %d\n",
                        t, t1 - t0, halo.size(), next, tp.max_img_points);
            return EUT_ERR_UNSUPPORTED;
        }
        T.img_points = std::max(T.img_points, next);
        T.max_tile_points = std::max(T.max_tile_points, t1 - t0);
        T.max_bulk = std::max(T.max_bulk, (int)bulk.size());
        T.max_single = std::max(T.max_single, (int)single.size());
        halo_total += (int64_t)halo.size();
        int32_t tx = 0;
        for (const int2 &b : bulk) tx += (b.y >> 16) * 8;
        int32_t *m = &T.meta[(size_t)t * 8];
        m[0] = t0; m[1] = t1 - t0;
        m[2] = (int32_t)T.copies.size(); m[3] = (int32_t)bulk.size();
        T.copies.insert(T.copies.end(), bulk.begin(), bulk.end());
        m[4] = (int32_t)T.copies.size(); m[5] = (int32_t)single.size();
        T.copies.insert(T.copies.end(), single.begin(), single.end());
        m[6] = tx; m[7] = next;
        // image offsets (bytes) of every in-neighbour, mat.in order; unused slots -> zero cell
        for (int32_t q = t0; q < t1; q++) {
            const int32_t i = p->iperm[q];
            const int d = (int)(p->row_ptr[i + 1] - p->row_ptr[i]);
            for (int k = 0; k < W; k++) {
                int32_t off = 0;
                if (k < d) {
                    const int32_t nq = p->perm[p->col[p->row_ptr[i] + k]];
                    if (nq >= t0 && nq < t1) off = own_off + (nq - t0);
                    else off = halo_off[std::lower_bound(halo.begin(), halo.end(), nq) - halo.begin()];
                }
                T.loc[(size_t)k * p->Np + q] = (uint16_t)(off * 8);
            }
        }
    }
    T.halo_ratio = (double)halo_total / (double)N;
    T.n_copies = (int64_t)T.copies.size();
    if (T.max_bulk > tp.max_bulk || T.max_single > tp.max_single) {
        if (getenv("EUT_TRACE"))
            fprintf(stderr, "[eut] tile planner: %d bulk / %d single copies in one tile exceed %d / %d\n",
                    T.max_bulk, T.max_single, tp.max_bulk, tp.max_single);
        return EUT_ERR_UNSUPPORTED;
    }
    p->n_tiles = n_tiles;
    return EUT_OK;
}

}  // namespace eut
