// Segment planner: tiles for the register-reuse substep kernel (stencil_seg.cu).
//
// The first tiled kernel reads 13 shared-memory values per point update and is bound by the
// shared-memory pipe (profiles/r01_v6_tile_lean_ncu.csv: 75 % of the LSU data pipe).  On the
// rhombic-dodecahedral mesh the 12 in-neighbours of R consecutive points of a raster row are not
// 12*R different values: they lie in seven "streams" of consecutive cells --
//     A, B : the two rows of the layer below   (cells u+s .. u+s+R, two neighbours per point)
//     C    : the previous row of the own layer (cells u   .. u+R-1)
//     O    : the own row                        (cells u-1 .. u+R: left, self, right)
//     D    : the next row of the own layer
//     E, F : the two rows of the layer above
// so a thread that owns a *segment* of R points needs 7R+6 loads instead of 13R, and sums every
// point's inflow in the canonical order A A B B C O- O+ D E E F F, which is the ascending-`from`
// order of mat.in (R/growthEnvironment.R:140-149).
//
// Nothing here assumes the geometry.  The streams of a segment are guessed from the row analysis
// (rows.cu) and then VERIFIED point by point: the canonical cell sequence, restricted to cells that
// exist as points, must equal the point's mat.in neighbour list exactly (same ids, same order).
// Cells that do not exist (polygon boundary, ragged row ends) read the zero header of the image;
// adding +0.0 to a chain that started at +0.0 is exact, so border points are regular too.  Points
// that fail alone are "generic" points and keep twelve explicit slots.
//
// Per tile the shared-memory image is [zero header | the tile's own points, dense, in internal order
// | halo runs | lone halo points].  Internal order inside a tile is level by level, so a tile's own
// block is ONE bulk copy and the rows it needs from the neighbouring tiles of its strip are a few
// contiguous runs (the copy count, not the bytes, bounds the producer warp: ~60 cycles per
// cp.async.bulk).  A stream must be contiguous in this image except for its two end cells -- the
// left / right neighbours of a row piece live in the next tile column -- which get explicit offsets.
#include <algorithm>
#include <numeric>

#include "common.h"

namespace eut {

namespace {
struct Piece {
    int32_t ucol, level, first, count, strip;
};
struct Dsu {
    std::vector<int32_t> p;
    explicit Dsu(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int32_t find(int32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int32_t a, int32_t b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};
enum { SA = 0, SB, SC, SD, SE_, SF, NSTREAM };   // streams other than the own row
struct Pattern {
    int32_t row[NSTREAM];
    int32_t sh[NSTREAM];      // u of the stream's first cell minus u of the point
};
}  // namespace

int build_segs(eut_plan *p, const SegParams &sp)
{
    const int64_t N = p->N;
    SegPlanHost &S = p->segs;
    S = SegPlanHost();
    const int R = sp.R;
    if (N == 0 || R < 1 || R > 15) return EUT_ERR_UNSUPPORTED;
    RowAnalysis ra;
    analyse_rows(p, ra);
    const int32_t NR = ra.n_rows;
    auto row_len = [&](int32_t r) { return ra.row_start[r + 1] - ra.row_start[r]; };
    auto u_of = [&](int32_t i) { const int32_t r = ra.row_of[i]; return ra.xs[r] + (i - ra.row_start[r]); };

    // ---- pieces, strips (as tiles.cu) -----------------------------------------------------------
    const int X = std::max(2, sp.tile_cols);
    std::vector<Piece> pieces;
    std::vector<int32_t> piece_of(N);
    for (int32_t r = 0; r < NR; r++) {
        const int32_t u0 = ra.xs[r] - ra.xmin;
        const int32_t len = row_len(r);
        int32_t k = 0;
        while (k < len) {
            int32_t ucol = (u0 + k) / X;
            int32_t kend = std::min<int32_t>(len, (ucol + 1) * X - u0);
            if (k == 0 && kend < len && kend - k < X / 2) {
                ucol += 1;
                kend = std::min<int32_t>(len, (ucol + 1) * X - u0);
            }
            if (len - kend > 0 && len - kend < X / 2) kend = len;
            for (int32_t i = k; i < kend; i++) piece_of[ra.row_start[r] + i] = (int32_t)pieces.size();
            // the column label is taken at the centre of the piece: a head that is one cell shorter or
            // longer than the neighbouring row's (the half-cell shift between layers along a ragged
            // boundary) must not relabel it, or the strips would fall apart into single rows
            ucol = (u0 + (k + kend) / 2) / X;
            pieces.push_back({ucol, ra.level[r], ra.row_start[r] + k, kend - k, 0});
            k = kend;
        }
    }
    const int32_t P = (int32_t)pieces.size();
    if (getenv("EUT_TRACE_PIECES"))
        for (int32_t a = 0; a < std::min(P, 60); a++)
            fprintf(stderr, "[eut] piece %d: row %d level %d u %d count %d ucol %d\n", a, ra.row_of[pieces[a].first],
                    pieces[a].level, u_of(pieces[a].first) - ra.xmin, pieces[a].count, pieces[a].ucol);
    {
        Dsu dsu(P);
        for (int64_t i = 0; i < N; i++) {
            const int32_t a = piece_of[i];
            for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                const int32_t b = piece_of[p->col[k]];
                if (a != b && pieces[a].ucol == pieces[b].ucol) dsu.unite(a, b);
            }
        }
        // Tiny strips join a neighbour, whatever its column label.  The ghost points of a slab (halo.cu)
        // are the typical case: they have no in-edges of their own, so each is a one-point row at the end
        // of a raster row with a column label of its own, and would otherwise become a tile of its own.
        {
            std::vector<int32_t> strip_pts(P, 0);
            for (int32_t a = 0; a < P; a++) strip_pts[dsu.find(a)] += pieces[a].count;
            const int32_t tiny = 4 * X;
            for (int64_t i = 0; i < N; i++) {
                const int32_t a = piece_of[i];
                for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                    const int32_t b = piece_of[p->col[k]];
                    const int32_t ra_ = dsu.find(a), rb_ = dsu.find(b);
                    if (ra_ == rb_) continue;
                    if (strip_pts[ra_] < tiny || strip_pts[rb_] < tiny) {
                        const int32_t tot = strip_pts[ra_] + strip_pts[rb_];
                        dsu.unite(ra_, rb_);
                        strip_pts[dsu.find(ra_)] = tot;
                    }
                }
            }
        }
        for (int32_t a = 0; a < P; a++) pieces[a].strip = dsu.find(a);
    }

    // ---- segments -------------------------------------------------------------------------------
    // pattern of a set of points of row r: the rows they touch, classified by side (lower / higher
    // ids) and by the pair flag of the row analysis
    auto adj_find = [&](int32_t r, int32_t rr) -> int64_t {
        const auto b = ra.adj_row.begin() + ra.adj_ptr[r], e = ra.adj_row.begin() + ra.adj_ptr[r + 1];
        const auto it = std::lower_bound(b, e, rr);
        return (it != e && *it == rr) ? (it - ra.adj_row.begin()) : -1;
    };
    auto derive = [&](int32_t first, int32_t len, Pattern &pat) -> bool {
        const int32_t r = ra.row_of[first];
        int32_t bp[2], bs[1], as[1], ap[2];
        int nbp = 0, nbs = 0, nas = 0, nap = 0;
        auto add = [](int32_t *v, int &n, int cap, int32_t x) -> bool {
            for (int i = 0; i < n; i++) if (v[i] == x) return true;
            if (n == cap) return false;
            v[n++] = x;
            return true;
        };
        for (int32_t i = first; i < first + len; i++)
            for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                const int32_t rr = ra.row_of[p->col[k]];
                if (rr == r) continue;
                const int64_t a = adj_find(r, rr);
                if (a < 0) return false;
                bool ok;
                if (rr < r) ok = ra.adj_pair[a] ? add(bp, nbp, 2, rr) : add(bs, nbs, 1, rr);
                else ok = ra.adj_pair[a] ? add(ap, nap, 2, rr) : add(as, nas, 1, rr);
                if (!ok) return false;
            }
        if (nbp == 2 && bp[0] > bp[1]) std::swap(bp[0], bp[1]);
        if (nap == 2 && ap[0] > ap[1]) std::swap(ap[0], ap[1]);
        for (int s = 0; s < NSTREAM; s++) { pat.row[s] = -1; pat.sh[s] = 0; }
        if (nbp > 0) pat.row[SA] = bp[0];
        if (nbp > 1) pat.row[SB] = bp[1];
        if (nbs > 0) pat.row[SC] = bs[0];
        if (nas > 0) pat.row[SD] = as[0];
        if (nap > 0) pat.row[SE_] = ap[0];
        if (nap > 1) pat.row[SF] = ap[1];
        for (int s = 0; s < NSTREAM; s++)
            if (pat.row[s] >= 0) {
                const int64_t a = adj_find(r, pat.row[s]);
                pat.sh[s] = ra.xs[pat.row[s]] - ra.xs[r] + ra.adj_delta[a];
            }
        return true;
    };
    // the exact check: canonical cells that exist as points == mat.in neighbour list of point i
    auto check_point = [&](int32_t i, const Pattern &pat) -> bool {
        const int32_t r = ra.row_of[i];
        const int32_t u = u_of(i);
        int32_t want[12];
        int n = 0;
        auto cell = [&](int s, int extra) {
            const int32_t rr = pat.row[s];
            if (rr < 0) return;
            const int32_t k = u + pat.sh[s] + extra - ra.xs[rr];
            if (k < 0 || k >= row_len(rr)) return;
            want[n++] = ra.row_start[rr] + k;
        };
        cell(SA, 0); cell(SA, 1); cell(SB, 0); cell(SB, 1); cell(SC, 0);
        if (i > ra.row_start[r]) want[n++] = i - 1;
        if (i + 1 < ra.row_start[r + 1]) want[n++] = i + 1;
        cell(SD, 0); cell(SE_, 0); cell(SE_, 1); cell(SF, 0); cell(SF, 1);
        if ((int64_t)n != p->row_ptr[i + 1] - p->row_ptr[i]) return false;
        for (int k = 0; k < n; k++)
            if (p->col[p->row_ptr[i] + k] != want[k]) return false;
        return true;
    };
    // ---- order pieces, cut strips into tiles ----------------------------------------------------
    std::vector<int32_t> strip_minlevel(P, INT32_MAX);
    for (int32_t a = 0; a < P; a++)
        strip_minlevel[pieces[a].strip] = std::min(strip_minlevel[pieces[a].strip], pieces[a].level);
    std::vector<int32_t> ord(P);
    std::iota(ord.begin(), ord.end(), 0);
    const int band = std::max(1, sp.band_levels);
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
    // capacity is counted in lanes: a piece of n points needs about ceil(n / R) segments (ragged
    // pieces a few more; the overflow becomes generic points below)
    auto units = [&](const Piece &pc) { return (pc.count + R - 1) / R; };
    // capacity in points is checked against an upper bound of the stored length (see below)
    auto stored = [&](int32_t count) { return sp.pad_rows ? count + 15 : count; };
    p->perm.assign(N, 0);
    std::vector<int32_t> tile_piece0;    // index into ord of every tile's first piece
    {
        int32_t in_units = 0, in_pts = 0, cur_strip = -1;
        size_t i = 0;
        auto open_tile = [&](size_t at) { tile_piece0.push_back((int32_t)at); in_units = 0; in_pts = 0; };
        while (i < ord.size()) {
            size_t e = i;
            int32_t gunits = 0, gpts = 0;
            while (e < ord.size() && pieces[ord[e]].strip == pieces[ord[i]].strip &&
                   pieces[ord[e]].level == pieces[ord[i]].level) { gunits += units(pieces[ord[e]]); gpts += stored(pieces[ord[e]].count); e++; }
            if (pieces[ord[i]].strip != cur_strip || in_units + gunits > sp.max_tile_segs ||
                in_pts + gpts > sp.max_tile_points) {
                open_tile(i);
                cur_strip = pieces[ord[i]].strip;
            }
            for (size_t k = i; k < e; k++) {
                const Piece &pc = pieces[ord[k]];
                if (in_pts > 0 && (in_units + units(pc) > sp.max_tile_segs || in_pts + stored(pc.count) > sp.max_tile_points))
                    open_tile(k);
                if (units(pc) > sp.max_tile_segs || stored(pc.count) > sp.max_tile_points) return EUT_ERR_UNSUPPORTED;
                in_pts += stored(pc.count); in_units += units(pc);
            }
            i = e;
        }
        tile_piece0.push_back((int32_t)ord.size());
    }
    const int32_t n_tiles = (int32_t)tile_piece0.size() - 1;
    if (getenv("EUT_TRACE")) {
        std::vector<int32_t> st;
        for (int32_t a = 0; a < P; a++) st.push_back(pieces[a].strip);
        std::sort(st.begin(), st.end());
        const size_t n_strips = std::unique(st.begin(), st.end()) - st.begin();
        int32_t max_level = 0;
        for (int32_t r = 0; r < NR; r++) max_level = std::max(max_level, ra.level[r]);
        fprintf(stderr, "[eut] segment planner: %d rows, %d pieces, %zu strips, %d levels, %d tiles\n", NR, P, n_strips, max_level + 1, n_tiles);
    }
    // Storage inside a tile.  Pieces follow the reference order (layer by layer, row by row), so the
    // rows a segment's streams read are a fixed number of pieces away from its own row: the previous
    // / next piece for C / D, a constant piece count for the other layers.  The own block of a tile
    // is dense in the image, so cell(row k, column u) = base_k + (u - ustart_k).  If
    //     rho_k = (base_k - ustart_k) mod 16   advances by the same step s from piece to piece,
    // all seven streams of every segment move together mod 16 and half warps can be packed without
    // bank conflicts (lane assignment below).  Regular tiles have this for free (equal ustart, 56
    // cells per piece: s = 8).  Ragged tiles (polygon border: rows start and end anywhere) get gap
    // cells after each piece, 0..15 of them, and the step s that needs the fewest.  Gap cells are
    // never read as neighbours, hold 0, and cost their share of HBM traffic (n_padded vs n_points).
    // Row index inside its layer: rows linked by the C / D relation (row adjacency without the pair
    // flag) are one apart, the one with the higher ids above.  A row and the row one of its streams
    // reads then differ by an index offset that is the same for all rows of its layer, as long as no
    // raster row is missing inside the tile -- which is what "all streams move together" needs.
    std::vector<int32_t> krow(NR, INT32_MIN);
    int64_t krow_clash = 0;
    {
        std::vector<int32_t> stack;
        for (int32_t seed = 0; seed < NR; seed++) {
            if (krow[seed] != INT32_MIN) continue;
            krow[seed] = 0;
            stack.push_back(seed);
            while (!stack.empty()) {
                const int32_t r = stack.back();
                stack.pop_back();
                for (int64_t a = ra.adj_ptr[r]; a < ra.adj_ptr[r + 1]; a++) {
                    if (ra.adj_pair[a]) continue;
                    const int32_t rr = ra.adj_row[a];
                    const int32_t want = krow[r] + (rr > r ? 1 : -1);
                    if (krow[rr] == INT32_MIN) { krow[rr] = want; stack.push_back(rr); }
                    else if (krow[rr] != want) krow_clash++;
                }
            }
        }
    }
    // Layer of every row, relative: rows linked without the pair flag share a layer, a pair-flagged neighbour with
    // higher ids lies one layer up (ids are layer major, R/growthEnvironment.R:104-117).  Only used to keep the
    // residue statistics of the gap rule apart per layer.
    constexpr int kMaxSegLayers = 16;
    std::vector<int32_t> layer_of_row(NR, INT32_MIN);
    int32_t min_layer = 0;
    {
        std::vector<int32_t> stack;
        for (int32_t seed = 0; seed < NR; seed++) {
            if (layer_of_row[seed] != INT32_MIN) continue;
            layer_of_row[seed] = 0;
            stack.push_back(seed);
            while (!stack.empty()) {
                const int32_t r = stack.back();
                stack.pop_back();
                for (int64_t a = ra.adj_ptr[r]; a < ra.adj_ptr[r + 1]; a++) {
                    const int32_t rr = ra.adj_row[a];
                    if (layer_of_row[rr] != INT32_MIN) continue;
                    layer_of_row[rr] = layer_of_row[r] + (ra.adj_pair[a] ? (rr > r ? 1 : -1) : 0);
                    stack.push_back(rr);
                }
            }
        }
        for (int32_t r = 0; r < NR; r++) min_layer = std::min(min_layer, layer_of_row[r]);
    }
    const bool balance_residues = !getenv("EUT_SEG_NO_BALANCE");
    // Residue step per strip (pad_rows 3): position - u == step * krow (mod 16).  Full-width pieces
    // (8 segments) pack two rows per half warp with step 8 and need no gaps; narrow strips (polygon
    // border) need more distinct row residues to fill 16 lanes with distinct bank pairs: step 4 / 2 / 1.
    std::vector<int8_t> strip_step(P, 8);
    if (sp.pad_rows == 3) {
        std::vector<std::vector<int32_t>> widths(P);
        for (int32_t a = 0; a < P; a++) widths[pieces[a].strip].push_back((pieces[a].count + R - 1) / R);
        for (int32_t st = 0; st < P; st++) {
            if (widths[st].empty()) continue;
            std::nth_element(widths[st].begin(), widths[st].begin() + widths[st].size() / 2, widths[st].end());
            const int32_t med = widths[st][widths[st].size() / 2];
            strip_step[st] = med >= 8 ? 8 : (med >= 4 ? 4 : (med >= 2 ? 2 : 1));
        }
    }
    // Storage order inside a tile: level by level (as the strips were cut) or reference order.  Level
    // major keeps the rows a tile reads from the tiles above / below it in its strip contiguous -- two
    // halo runs instead of one per layer and side, i.e. 3 instead of 7 bulk copies per tile-column,
    // which the single issuing thread of the producer warp feels.  The price: rows of one layer are a
    // whole level group apart, (pieces per level) x 56 cells, which is 8 (mod 16) only for an odd
    // number of layers; with an even number the two rows of a half warp would collide, so reference
    // order is kept there.
    bool level_major = false;
    {
        std::vector<int32_t> per_level;
        for (size_t i = 0; i < ord.size();) {
            size_t e = i;
            while (e < ord.size() && pieces[ord[e]].strip == pieces[ord[i]].strip && pieces[ord[e]].level == pieces[ord[i]].level) e++;
            per_level.push_back((int32_t)(e - i));
            i = e;
        }
        if (!per_level.empty()) {
            std::nth_element(per_level.begin(), per_level.begin() + per_level.size() / 2, per_level.end());
            level_major = (per_level[per_level.size() / 2] % 2) == 1;
        }
        if (const char *e = getenv("EUT_SEG_ORDER")) level_major = atoi(e) != 0;
        if (sp.pad_rows == 3) level_major = false;      // the global rule is written for reference order
    }
    S.tile_start.assign(1, 0);
    std::vector<int32_t> tile_t0;        // first cell of every tile (tile t ends at tile_start[t + 1])
    {
        int32_t pos = 0;
        std::vector<int32_t> gap;
        int64_t total_gap = 0;
        for (int32_t t = 0; t < n_tiles; t++) {
            if (!level_major)
                std::sort(ord.begin() + tile_piece0[t], ord.begin() + tile_piece0[t + 1],
                          [&](int32_t a, int32_t b) { return pieces[a].first < pieces[b].first; });
            const int32_t k0 = tile_piece0[t], k1 = tile_piece0[t + 1];
            gap.assign(k1 - k0, 0);
            if (sp.pad_rows == 1 && k1 - k0 > 1) {
                // coordinate rule: rho = position - u of a piece is a function of WHERE its row is,
                //     rho == a * (row index inside its layer) + b * layer + (odd layer ? c : 0)   (mod 16),
                // so every stream of every segment of the tile's own block keeps the same residue shift against
                // the segment's own cells -- however many pieces a level has and in whatever order they are
                // stored (the chain rule below loses that in ragged tiles: rows cut into two pieces, layers
                // missing at the border) -- and a half warp is conflict free in all seven streams as soon as
                // its 16 segments start on 16 different residues.  (a, b, c) per tile: the triple that costs the
                // fewest gap cells plus the extra half-warp passes an unbalanced residue histogram of some layer
                // would force (16 x R cells each); a = 0 puts all rows of a layer on the same 8 residues and
                // doubles the tile's time (tools/tile_times.py: 75-86 us against 42 us per 32 columns).
                const int32_t np_ = k1 - k0;
                std::vector<int32_t> kr(np_), ly(np_), uu(np_), nseg(np_), dfix(np_, 0);
                int nl = 0;
                for (int32_t k = 0; k < np_; k++) {
                    const Piece &pc = pieces[ord[k0 + k]];
                    const int32_t r = ra.row_of[pc.first];
                    kr[k] = krow[r];
                    ly[k] = std::min(kMaxSegLayers - 1, std::max(0, layer_of_row[r] - min_layer));
                    nl = std::max(nl, ly[k] + 1);
                    uu[k] = u_of(pc.first);
                    nseg[k] = (pc.count + R - 1) / R;
                    if (k > 0) dfix[k] = uu[k] - uu[k - 1] - pieces[ord[k0 + k - 1]].count;
                }
                int64_t best_cost = INT64_MAX;
                int best_a = 8, best_b = 0, best_c = 0;
                std::vector<int32_t> want(np_);
                for (int a = 0; a < 16; a++)
                    for (int b = 0; b < 16; b++)
                        for (int c = 0; c < 16; c++) {
                            int64_t total = 0;
                            for (int32_t k = 0; k < np_; k++) {
                                want[k] = a * kr[k] + b * ly[k] + ((ly[k] & 1) ? c : 0);
                                if (k > 0) total += ((want[k] - want[k - 1] + dfix[k]) % 16 + 16) % 16;
                            }
                            if (total >= best_cost) continue;                 // the penalty only adds
                            int32_t hist[kMaxSegLayers][16] = {};
                            int32_t n_layer[kMaxSegLayers] = {};
                            for (int32_t k = 0; k < np_; k++) {
                                const int32_t r0 = ((want[k] + uu[k]) % 16 + 16) % 16;
                                for (int32_t j = 0; j < nseg[k]; j++) hist[ly[k]][(r0 + R * j) & 15]++;
                                n_layer[ly[k]] += nseg[k];
                            }
                            int64_t extra = 0;
                            for (int l = 0; l < nl; l++)
                                if (n_layer[l]) extra += std::max(0, *std::max_element(hist[l], hist[l] + 16) - (n_layer[l] + 15) / 16);
                            const int64_t cost = total + extra * 16 * R;
                            if (cost < best_cost) { best_cost = cost; best_a = a; best_b = b; best_c = c; }
                        }
                if (const char *e = getenv("EUT_TRACE_TILE"))
                    if (atoi(e) == t) {
                        fprintf(stderr, "[eut] tile %d: a %d b %d c %d cost %lld\n", t, best_a, best_b, best_c, (long long)best_cost);
                        for (int32_t k = 0; k < np_; k++)
                            fprintf(stderr, "[eut]   piece %3d: row %5d level %4d layer %d krow %4d u %5d count %3d\n", k, ra.row_of[pieces[ord[k0 + k]].first],
                                    pieces[ord[k0 + k]].level, ly[k], kr[k], uu[k], pieces[ord[k0 + k]].count);
                    }
                int64_t tile_pts = 0, tile_gap = 0;
                for (int32_t k = 0; k < np_; k++) tile_pts += pieces[ord[k0 + k]].count;
                for (int32_t k = 1; k < np_; k++) {
                    const int32_t w1 = best_a * kr[k] + best_b * ly[k] + ((ly[k] & 1) ? best_c : 0);
                    const int32_t w0 = best_a * kr[k - 1] + best_b * ly[k - 1] + ((ly[k - 1] & 1) ? best_c : 0);
                    gap[k - 1] = ((w1 - w0 + dfix[k]) % 16 + 16) % 16;
                    tile_gap += gap[k - 1];
                }
                // not a lattice (renumbered ids: rows of one or two points): residues cannot be arranged, and the
                // gap cells would outnumber the points
                if (tile_gap * 4 > tile_pts) gap.assign(k1 - k0, 0);
            }
            if (sp.pad_rows == 2 && k1 - k0 > 1) {
                // chain rule: rho advances by the same step s from piece to piece.  Which s?  The one that costs
                // the fewest gap cells -- as long as the segments' start residues stay BALANCED inside every
                // layer: a half warp is conflict free only if its 16 segments start on 16 different residues, so
                // a residue that n segments of a layer share forces n half warps however the lanes are dealt.
                // (A tile whose rows come in an even number of pieces per level has a same-layer step of 0 with
                // s = 8: every row on the same 8 residues, twice the shared-memory wavefronts and twice the tile
                // time -- 75-86 us against 42 us per 32 columns on the bench mesh, tools/tile_times.py.)  Cost of
                // an s = gap cells + the extra half-warp passes its worst residue forces, at 16 x R cells each.
                int64_t best_cost = INT64_MAX;
                int best_s = 8;
                for (int s = 0; s < 16; s++) {
                    int64_t total = 0;
                    int32_t hist[kMaxSegLayers][16] = {};
                    int32_t n_layer[kMaxSegLayers] = {};
                    int32_t at = 0;                              // position relative to the tile's first cell
                    for (int32_t k = k0; k < k1; k++) {
                        const Piece &a = pieces[ord[k]];
                        const int l = std::min(kMaxSegLayers - 1, std::max(0, layer_of_row[ra.row_of[a.first]] - min_layer));
                        for (int32_t q = 0; q < a.count; q += R) { hist[l][(at + q) & 15]++; n_layer[l]++; }
                        at += a.count;
                        if (k + 1 < k1) {
                            const Piece &b = pieces[ord[k + 1]];
                            const int32_t g = (((s + u_of(b.first) - u_of(a.first) - a.count) % 16) + 16) % 16;
                            total += g; at += g;
                        }
                    }
                    int64_t extra = 0;
                    for (int l = 0; l < kMaxSegLayers; l++) {
                        if (n_layer[l] == 0) continue;
                        const int32_t need = (n_layer[l] + 15) / 16;
                        extra += std::max(0, *std::max_element(hist[l], hist[l] + 16) - need);
                    }
                    const int64_t cost = total + (balance_residues ? extra * 16 * R : 0);
                    if (cost < best_cost || (cost == best_cost && s == 8)) { best_cost = cost; best_s = s; }
                }
                for (int32_t k = k0; k + 1 < k1; k++) {
                    const Piece &a = pieces[ord[k]], &b = pieces[ord[k + 1]];
                    gap[k - k0] = (((best_s + u_of(b.first) - u_of(a.first) - a.count) % 16) + 16) % 16;
                }
            }
            if (sp.pad_rows == 3) {
                // global rule: position - u == step * krow (mod 16) for every piece of the strip, so that
                // the rows of OTHER tiles (halo) keep the same relation; the image keeps
                // cell == position (mod 16).  The gap sits in front of a piece.
                for (int32_t k = k0; k < k1; k++) {
                    const Piece &pc = pieces[ord[k]];
                    const int32_t want = (int32_t)((((int64_t)u_of(pc.first) + (int64_t)strip_step[pc.strip] * krow[ra.row_of[pc.first]]) % 16 + 16) % 16);
                    const int32_t g = (((want - pos) % 16) + 16) % 16;
                    pos += g; total_gap += g;
                    if (k == k0) tile_t0.push_back(pos);
                    for (int32_t q = 0; q < pc.count; q++) p->perm[pc.first + q] = pos + q;
                    pos += pc.count;
                }
                S.tile_start.push_back(pos);
                continue;
            }
            tile_t0.push_back(pos);
            for (int32_t k = k0; k < k1; k++) {
                const Piece &pc = pieces[ord[k]];
                for (int32_t q = 0; q < pc.count; q++) p->perm[pc.first + q] = pos + q;
                pos += pc.count + gap[k - k0];
                total_gap += gap[k - k0];
            }
            S.tile_start.push_back(pos);
        }
        if (getenv("EUT_TRACE"))
            fprintf(stderr, "[eut] segment planner: %lld gap cells (%.2f %%), %lld row index clashes\n", (long long)total_gap,
                    100.0 * (double)total_gap / (double)N, (long long)krow_clash);
    }
    p->Np = (((int64_t)S.tile_start.back() + kPadPoints - 1) / kPadPoints) * kPadPoints;
    p->iperm.assign(p->Np, -1);
    for (int64_t i = 0; i < N; i++) p->iperm[p->perm[i]] = (int32_t)i;

    // ---- per tile: image, copy lists, segments ----------------------------------------------------
    // image = [zero header | own points, dense, in internal order (ONE bulk copy) | halo runs: maximal
    // runs of consecutive internal ids among the in-neighbours outside the tile, small gaps filled
    // (bulk copies) | lone halo points (8-byte copies)].  Everything a stream reads must be contiguous
    // in this image except the two end cells of the O and pair streams, which get explicit offsets
    // (the row ends: their neighbours belong to the next tile column, or do not exist -> zero cell).
    const int ZH = (R + 3 + 1) & ~1;
    S.R = R; S.zero_cells = ZH;
    S.meta.assign((size_t)n_tiles * 12, 0);
    int64_t halo_total = 0;
    std::vector<int32_t> halo, halo_off;
    std::vector<int2> bulk, single;
    struct Rec { int cls; uint16_t w[24]; };
    std::vector<Rec> recs;
    std::vector<int32_t> tgens;
    // dead points (ghosts of a slab) keep their cell -- neighbours read it -- but get neither a segment nor a
    // generic record: their result cell stays +0.0 like a gap cell, and the halo exchange fills it in
    auto is_dead = [&](int32_t i) { return !p->dead.empty() && p->dead[i] != 0; };
    auto emit = [&](int32_t g, int32_t s, int32_t len) {
        if ((g ^ s) & 1) {      // parities differ: no 16-byte aligned pairing exists
            for (int32_t k = 0; k < len; k++) single.push_back({g + k, s + k});
            return;
        }
        if (len > 0 && (g & 1)) { single.push_back({g, s}); g++; s++; len--; }
        const int32_t body = len & ~1;
        if (body > 0) bulk.push_back({g, s | (body << 16)});
        if (len & 1) single.push_back({g + body, s + body});
    };
    const int kGap = 6;          // halo runs closer than this are merged (the cells between come along)
    for (int32_t t = 0; t < n_tiles; t++) {
        const int32_t t0 = tile_t0[t], t1 = S.tile_start[t + 1];
        halo.clear(); bulk.clear(); single.clear(); recs.clear(); tgens.clear();
        for (int32_t q = t0; q < t1; q++) {
            const int32_t i = p->iperm[q];
            if (i < 0) continue;                       // gap cell
            for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                const int32_t nq = p->perm[p->col[k]];
                if (nq < t0 || nq >= t1) halo.push_back(nq);
            }
        }
        std::sort(halo.begin(), halo.end());
        halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
        // pad_rows 3: image cell == internal position (mod 16) everywhere, which carries the global
        // residue rule into the image (and implies the 16-byte parity rule of the bulk copies)
        const bool keep_res = sp.pad_rows == 3;
        const int32_t own_off = keep_res ? ZH + (((t0 - ZH) % 16) + 16) % 16 : ZH + ((t0 ^ ZH) & 1);
        emit(t0, own_off, t1 - t0);
        int32_t next = own_off + (t1 - t0);
        halo_off.assign(halo.size(), -1);
        for (size_t h = 0; h < halo.size();) {
            size_t e = h + 1;
            while (e < halo.size() && halo[e] - halo[e - 1] <= kGap) e++;
            const int32_t span = halo[e - 1] - halo[h] + 1;
            if (span >= 3) {
                if (keep_res) next += (((halo[h] - next) % 16) + 16) % 16;
                else if ((next & 1) != (halo[h] & 1)) next++;
                for (size_t k = h; k < e; k++) halo_off[k] = next + (halo[k] - halo[h]);
                emit(halo[h], next, span);
                halo_total += span;
                next += span;
            }
            h = e;
        }
        {
            // Lone points: mostly the left / right neighbour of a row piece, which lives in the next tile
            // column.  A lane reads it in place of the cell before / after its row, so it is placed at
            // the residue (mod 16 cells) that cell would have had -- a 16-column grid, column = residue --
            // and the end-cell loads of a half warp stay on distinct bank pairs like the in-line ones.
            int32_t fill[16] = {0};
            next = (next + 15) / 16 * 16;
            int32_t rr = 0;
            for (size_t h = 0; h < halo.size(); h++) {
                if (halo_off[h] >= 0) continue;
                int32_t r = -1;
                if (keep_res) r = halo[h] & 15;
                else {
                    const int32_t i = p->iperm[halo[h]];
                    if (i + 1 < (int32_t)N && ra.row_of[i + 1] == ra.row_of[i] && p->perm[i + 1] >= t0 && p->perm[i + 1] < t1)
                        r = (own_off + (p->perm[i + 1] - t0) - 1) & 15;
                    else if (i > 0 && ra.row_of[i - 1] == ra.row_of[i] && p->perm[i - 1] >= t0 && p->perm[i - 1] < t1)
                        r = (own_off + (p->perm[i - 1] - t0) + 1) & 15;
                    else r = (rr++) & 15;
                }
                halo_off[h] = next + 16 * fill[r] + r;
                fill[r]++;
                single.push_back({halo[h], halo_off[h]});
                halo_total++;
            }
            next += 16 * *std::max_element(fill, fill + 16);
        }
        next += R + 2;                              // masked lanes of short segments read past their stream
        if (next > sp.max_img_points || next * 8 > 65535) {
            if (getenv("EUT_TRACE"))
                fprintf(stderr, "[eut] segment planner: tile %d (%d points) needs an image of %d cells > %d\n",
                        t, t1 - t0, next, sp.max_img_points);
            return EUT_ERR_UNSUPPORTED;
        }
        // image cell of a point (reference id); -1 if the tile does not hold it
        auto cell_of = [&](int32_t i) -> int32_t {
            const int32_t nq = p->perm[i];
            if (nq >= t0 && nq < t1) return own_off + (nq - t0);
            const auto it = std::lower_bound(halo.begin(), halo.end(), nq);
            if (it == halo.end() || *it != nq) return -1;
            return halo_off[it - halo.begin()];
        };
        // cell of the canonical neighbour (row rr, column u): 0 = hole (reads the zero header)
        auto cell_at = [&](int32_t rr, int32_t u) -> int32_t {
            if (rr < 0) return 0;
            const int32_t k = u - ra.xs[rr];
            if (k < 0 || k >= row_len(rr)) return 0;
            return cell_of(ra.row_start[rr] + k);
        };
        // the layout side of a segment: fills the record, or fails
        auto layout = [&](int32_t first, int32_t len, const Pattern &pat, int32_t piece_end, Rec &rec) -> bool {
            const int32_t r = ra.row_of[first], u0 = u_of(first);
            const bool full = (len == R);
            for (int k = 0; k < 24; k++) rec.w[k] = 0;
            const int32_t base_o = cell_of(first);
            if (!full && first + len >= piece_end) return false;    // o[len + 1] is read in line
            const int32_t c_left = cell_at(r, u0 - 1), c_right = full ? cell_at(r, u0 + R) : 0;
            if (base_o < 0 || c_left < 0 || c_right < 0) return false;
            rec.w[0] = (uint16_t)(base_o * 8);
            rec.w[8] = (uint16_t)(c_left * 8);
            rec.w[9] = (uint16_t)(c_right * 8);
            rec.cls = 0;
            static const int slot_first[NSTREAM] = {10, 12, -1, -1, 14, 16};
            for (int x = 0; x < NSTREAM; x++) {
                const int32_t rr = pat.row[x];
                if (rr < 0) continue;
                const int32_t ub = u0 + pat.sh[x];
                if (x == SC || x == SD) {                            // single stream: all cells in line
                    const int32_t b = cell_at(rr, ub);
                    if (b <= 0) return false;
                    for (int32_t i = 1; i < len; i++)
                        if (cell_at(rr, ub + i) != b + i) return false;
                    rec.w[1 + x] = (uint16_t)(b * 8);
                } else {                                             // pair stream: cells 0 .. len
                    const int32_t c1 = cell_at(rr, ub + 1);
                    if (c1 <= 1) return false;
                    const int32_t last_inline = full ? R - 1 : len;
                    for (int32_t i = 2; i <= last_inline; i++)
                        if (cell_at(rr, ub + i) != c1 + (i - 1)) return false;
                    rec.w[1 + x] = (uint16_t)((c1 - 1) * 8);
                    const int32_t c0 = cell_at(rr, ub), cl = full ? cell_at(rr, ub + R) : 0;
                    if (c0 < 0 || cl < 0) return false;
                    rec.w[slot_first[x]] = (uint16_t)(c0 * 8);
                    rec.w[slot_first[x] + 1] = (uint16_t)(cl * 8);
                    rec.cls |= (x == SA || x == SB) ? 1 : 2;
                }
            }
            rec.w[7] = (uint16_t)((p->perm[first] - t0 + (t0 & 1)) | (len << 12));
            return true;
        };
        for (int32_t k = tile_piece0[t]; k < tile_piece0[t + 1]; k++) {
            const Piece &pc = pieces[ord[k]];
            int32_t i = pc.first;
            const int32_t end = pc.first + pc.count;
            while (i < end) {
                // longest run the graph check accepts ...
                Pattern pat, best;
                int32_t len = 0;
                while (len < R && i + len < end) {
                    if (len > 0 && check_point(i + len, best)) { len++; continue; }
                    if (!derive(i, len + 1, pat)) break;
                    bool ok = true;
                    for (int32_t j = 0; j <= len && ok; j++) ok = check_point(i + j, pat);
                    if (!ok) break;
                    best = pat;
                    len++;
                }
                // ... shortened until the image layout accepts it as well
                Rec rec;
                bool placed = false;
                while (len > 0 && !placed) {
                    if (derive(i, len, pat)) {
                        bool ok = true;
                        for (int32_t j = 0; j < len && ok; j++) ok = check_point(i + j, pat);
                        if (ok && layout(i, len, pat, end, rec)) placed = true;
                    }
                    if (!placed) len--;
                }
                if (placed) { recs.push_back(rec); S.seg_points += len; i += len; }
                else { if (!is_dead(i)) tgens.push_back(i); i++; }
            }
        }
        // Lane assignment.  Lanes of a warp should run the same stream set (class), and the 16 lanes of
        // a half warp should read 16 different 8-byte bank pairs: an LDS.64 is conflict free iff the
        // lanes' cells differ mod 16.  All streams of a segment move with its own cell (the rows of a
        // regular tile are a constant distance apart); in ragged tiles they do not, so half warps are
        // filled greedily with the segment that adds the fewest residue clashes over all seven streams
        // (ties: the fullest own-cell residue bucket, which keeps the remainder balanced).
        std::stable_sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) { return a.cls < b.cls; });
        {
            std::vector<Rec> out;
            out.reserve(recs.size());
            std::vector<uint8_t> taken(recs.size(), 0);
            int bucket[16];
            uint32_t used[7] = {0, 0, 0, 0, 0, 0, 0};
            size_t c0 = 0;
            while (c0 < recs.size()) {
                size_t c1 = c0;
                while (c1 < recs.size() && recs[c1].cls == recs[c0].cls) c1++;
                for (int r = 0; r < 16; r++) bucket[r] = 0;
                for (size_t k = c0; k < c1; k++) bucket[(recs[k].w[0] / 8) & 15]++;
                for (size_t n = c0; n < c1; n++) {
                    if (out.size() % 16 == 0) for (int x = 0; x < 7; x++) used[x] = 0;
                    int64_t pick = -1;
                    int pick_cost = 99, pick_bucket = -1;
                    for (size_t k = c0; k < c1; k++) {
                        if (taken[k]) continue;
                        int cost = 0;
                        for (int x = 0; x < 7; x++)
                            if (recs[k].w[x]) cost += (used[x] >> ((recs[k].w[x] / 8) & 15)) & 1u;
                        const int bsz = bucket[(recs[k].w[0] / 8) & 15];
                        if (cost < pick_cost || (cost == pick_cost && bsz > pick_bucket)) { pick = (int64_t)k; pick_cost = cost; pick_bucket = bsz; }
                    }
                    taken[pick] = 1;
                    bucket[(recs[pick].w[0] / 8) & 15]--;
                    for (int x = 0; x < 7; x++)
                        if (recs[pick].w[x]) used[x] |= 1u << ((recs[pick].w[x] / 8) & 15);
                    out.push_back(recs[pick]);
                }
                c0 = c1;
            }
            recs.swap(out);
        }
        // more segments than lanes: the excess is computed point by point
        while ((int32_t)recs.size() > sp.max_tile_segs) {
            const Rec &rc = recs.back();
            const int32_t q0 = t0 + (rc.w[7] & 0xfff) - (t0 & 1), len = rc.w[7] >> 12;
            for (int32_t j = 0; j < len; j++) if (!is_dead(p->iperm[q0 + j])) tgens.push_back(p->iperm[q0 + j]);
            S.seg_points -= len;
            recs.pop_back();
        }
        int32_t *m = &S.meta[(size_t)t * 12];
        m[0] = t0; m[1] = t1 - t0;
        m[2] = (int32_t)S.copies.size(); m[3] = (int32_t)bulk.size();
        S.copies.insert(S.copies.end(), bulk.begin(), bulk.end());
        m[4] = (int32_t)S.copies.size(); m[5] = (int32_t)single.size();
        S.copies.insert(S.copies.end(), single.begin(), single.end());
        int32_t tx = 0;
        for (const int2 &b : bulk) tx += (b.y >> 16) * 8;
        m[6] = tx; m[7] = next;
        m[8] = (int32_t)(S.segs.size() / 24); m[9] = (int32_t)recs.size();
        m[10] = (int32_t)(S.gens.size() / 16); m[11] = (int32_t)tgens.size();
        for (const Rec &rc : recs) S.segs.insert(S.segs.end(), rc.w, rc.w + 24);
        for (const int32_t i : tgens) {
            const int d = (int)(p->row_ptr[i + 1] - p->row_ptr[i]);
            for (int k = 0; k < 12; k++) S.gens.push_back(k < d ? (uint16_t)cell_of(p->col[p->row_ptr[i] + k]) : (uint16_t)0);
            S.gens.push_back((uint16_t)cell_of(i));
            S.gens.push_back((uint16_t)(p->perm[i] - t0 + (t0 & 1)));
            S.gens.push_back(0); S.gens.push_back(0);
        }
        S.img_points = std::max(S.img_points, next);
        S.max_tile_points = std::max(S.max_tile_points, t1 - t0);
        S.max_tile_segs = std::max(S.max_tile_segs, (int)recs.size());
        S.max_gen = std::max(S.max_gen, (int)tgens.size());
        S.max_bulk = std::max(S.max_bulk, (int)bulk.size());
        S.max_single = std::max(S.max_single, (int)single.size());
    }
    S.halo_ratio = (double)halo_total / (double)N;
    S.n_copies = (int64_t)S.copies.size();
    S.n_segs = (int64_t)(S.segs.size() / 24);
    S.n_gens = (int64_t)(S.gens.size() / 16);
    if (S.max_bulk > sp.max_bulk || S.max_single > sp.max_single) {
        if (getenv("EUT_TRACE"))
            fprintf(stderr, "[eut] segment planner: %d bulk / %d single copies in one tile exceed %d / %d\n",
                    S.max_bulk, S.max_single, sp.max_bulk, sp.max_single);
        return EUT_ERR_UNSUPPORTED;
    }
    p->n_tiles = n_tiles;
    return EUT_OK;
}

}  // namespace eut
