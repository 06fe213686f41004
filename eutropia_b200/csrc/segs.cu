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
// Cells that do not exist (polygon boundary, ragged row ends) are holes of the shared-memory image
// and stay +0.0; adding +0.0 to a chain that started at +0.0 is exact, so border points are regular
// too.  Points that fail the check alone are "generic" points and keep twelve explicit slots.
//
// Per tile the image is [zero header | one image row per touched graph row, covering every cell a
// canonical read can touch]; the copy lists move the existing cells (bulk copies for 16-byte
// aligned runs of consecutive internal ids, 8-byte copies for odd ends and apron cells).
#include <algorithm>
#include <numeric>

#include "common.h"

namespace eut {

namespace {
struct Piece {
    int32_t ucol, level, first, count, strip;
    int32_t seg0, nseg, gen0, ngen;     // ranges in the global segment / generic lists
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
struct Seg {
    int32_t first, len;       // reference id of the first point
    Pattern pat;
};
struct RowWin { int32_t row, lo, hi, base; };
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
            pieces.push_back({ucol, ra.level[r], ra.row_start[r] + k, kend - k, 0, 0, 0, 0, 0});
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
    std::vector<Seg> segs;
    std::vector<int32_t> gens;
    segs.reserve((size_t)(N / R + P));
    for (int32_t a = 0; a < P; a++) {
        Piece &pc = pieces[a];
        pc.seg0 = (int32_t)segs.size(); pc.gen0 = (int32_t)gens.size();
        int32_t i = pc.first;
        const int32_t end = pc.first + pc.count;
        while (i < end) {
            Pattern pat, best;
            int32_t len = 0;
            while (len < R && i + len < end) {
                if (len > 0 && check_point(i + len, best)) { len++; continue; }
                // first point, or a point that brings a row the pattern does not have yet: refine
                if (!derive(i, len + 1, pat)) break;
                bool ok = true;
                for (int32_t j = 0; j <= len && ok; j++) ok = check_point(i + j, pat);
                if (!ok) break;
                best = pat;
                len++;
            }
            if (len == 0) { gens.push_back(i); i++; }
            else { segs.push_back({i, len, best}); i += len; }
        }
        pc.nseg = (int32_t)segs.size() - pc.seg0; pc.ngen = (int32_t)gens.size() - pc.gen0;
    }

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
    auto units = [](const Piece &pc) { return pc.nseg + (pc.ngen + 1) / 2; };
    p->perm.assign(N, 0);
    std::vector<int32_t> tile_piece0;    // index into ord of every tile's first piece
    S.tile_start.clear();
    {
        int32_t pos = 0, in_units = 0, in_pts = 0, cur_strip = -1;
        size_t i = 0;
        auto open_tile = [&](size_t at) { S.tile_start.push_back(pos); tile_piece0.push_back((int32_t)at); in_units = 0; in_pts = 0; };
        while (i < ord.size()) {
            size_t e = i;
            int32_t gunits = 0, gpts = 0;
            while (e < ord.size() && pieces[ord[e]].strip == pieces[ord[i]].strip &&
                   pieces[ord[e]].level == pieces[ord[i]].level) { gunits += units(pieces[ord[e]]); gpts += pieces[ord[e]].count; e++; }
            if (pieces[ord[i]].strip != cur_strip || in_units + gunits > sp.max_tile_segs ||
                in_pts + gpts > sp.max_tile_points) {
                open_tile(i);
                cur_strip = pieces[ord[i]].strip;
            }
            for (size_t k = i; k < e; k++) {
                const Piece &pc = pieces[ord[k]];
                if (in_pts > 0 && (in_units + units(pc) > sp.max_tile_segs || in_pts + pc.count > sp.max_tile_points))
                    open_tile(k);
                if (units(pc) > sp.max_tile_segs || pc.count > sp.max_tile_points) return EUT_ERR_UNSUPPORTED;
                pos += pc.count; in_pts += pc.count; in_units += units(pc);
            }
            i = e;
        }
        S.tile_start.push_back(pos);
        tile_piece0.push_back((int32_t)ord.size());
    }
    const int32_t n_tiles = (int32_t)S.tile_start.size() - 1;
    // Inside a tile the pieces -- and with them the internal ids, the segments and the lanes that
    // run them -- follow the reference id: layer by layer, row by row.  Consecutive lanes then hold
    // consecutive segments of a row and consecutive rows of a layer (the bank rule below relies on
    // it), and warps are uniform in their stream set except where a layer ends.
    for (int32_t t = 0; t < n_tiles; t++) {
        std::sort(ord.begin() + tile_piece0[t], ord.begin() + tile_piece0[t + 1],
                  [&](int32_t a, int32_t b) { return pieces[a].first < pieces[b].first; });
        int32_t pos = S.tile_start[t];
        for (int32_t k = tile_piece0[t]; k < tile_piece0[t + 1]; k++) {
            const Piece &pc = pieces[ord[k]];
            for (int32_t q = 0; q < pc.count; q++) p->perm[pc.first + q] = pos + q;
            pos += pc.count;
        }
    }
    p->Np = ((N + kPadPoints - 1) / kPadPoints) * kPadPoints;
    p->iperm.assign(p->Np, -1);
    for (int64_t i = 0; i < N; i++) p->iperm[p->perm[i]] = (int32_t)i;

    // ---- per tile: image rows, copy lists, segment and generic records ---------------------------
    const int ZH = (R + 3 + 1) & ~1;           // zero header: absent streams and idle lanes read it
    S.R = R; S.zero_cells = ZH;
    S.meta.assign((size_t)n_tiles * 12, 0);
    S.segs.reserve(segs.size() * 8);
    S.gens.reserve(gens.size() * 16);
    int64_t halo_total = 0;
    std::vector<RowWin> wins;
    std::vector<int2> bulk, single;
    std::vector<std::pair<int32_t, int32_t>> tsegs;   // (class, index into segs)
    std::vector<int32_t> tgens;
    std::vector<std::vector<int32_t>> refs;           // per image row: (half warp * 7 + stream) * 16 + (u mod 16)
    std::vector<int32_t> order_rows, bank_cnt;
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
    for (int32_t t = 0; t < n_tiles; t++) {
        const int32_t t0 = S.tile_start[t], t1 = S.tile_start[t + 1];
        wins.clear(); bulk.clear(); single.clear(); tsegs.clear(); tgens.clear();
        auto touch = [&](int32_t row, int32_t lo, int32_t hi) { wins.push_back({row, lo, hi, 0}); };
        for (int32_t k = tile_piece0[t]; k < tile_piece0[t + 1]; k++) {
            const Piece &pc = pieces[ord[k]];
            const int32_t r = ra.row_of[pc.first];
            touch(r, u_of(pc.first), u_of(pc.first) + pc.count - 1);
            for (int32_t s = pc.seg0; s < pc.seg0 + pc.nseg; s++) {
                const Seg &sg = segs[s];
                const int32_t u0 = u_of(sg.first);
                touch(r, u0 - 1, u0 + sg.len);
                int cls = 0;
                for (int x = 0; x < NSTREAM; x++) {
                    if (sg.pat.row[x] < 0) continue;
                    const bool pair = (x == SA || x == SB || x == SE_ || x == SF);
                    touch(sg.pat.row[x], u0 + sg.pat.sh[x], u0 + sg.pat.sh[x] + sg.len - 1 + (pair ? 1 : 0));
                    if (x == SA || x == SB) cls |= 1;
                    if (x == SE_ || x == SF) cls |= 2;
                }
                tsegs.emplace_back(cls, s);
            }
            for (int32_t g = pc.gen0; g < pc.gen0 + pc.ngen; g++) {
                const int32_t i = gens[g];
                for (int64_t e = p->row_ptr[i]; e < p->row_ptr[i + 1]; e++) {
                    const int32_t n = p->col[e];
                    touch(ra.row_of[n], u_of(n), u_of(n));
                }
                tgens.push_back(i);
            }
        }
        // merge the windows per row
        std::sort(wins.begin(), wins.end(), [](const RowWin &a, const RowWin &b) { return a.row < b.row; });
        {
            size_t w = 0;
            for (size_t k = 0; k < wins.size(); k++) {
                if (w > 0 && wins[w - 1].row == wins[k].row) {
                    wins[w - 1].lo = std::min(wins[w - 1].lo, wins[k].lo);
                    wins[w - 1].hi = std::max(wins[w - 1].hi, wins[k].hi);
                } else wins[w++] = wins[k];
            }
            wins.resize(w);
        }
        // Placement.  What matters for the shared-memory pipe is cell(row, u) mod 16 (an 8-byte
        // bank pair): an LDS.64 of a stream is conflict free iff the 16 lanes of a half warp read 16
        // different residues.  Every image row gets a residue rho = (base - lo) mod 16, chosen
        // greedily -- own rows in lane order, then the halo rows -- to minimise the conflicts with
        // the rows placed before it, counted over all (half warp, stream) groups that read it.
        // On the regular interior this reproduces "consecutive rows of a layer differ by 8 cells";
        // ragged tiles get the best the greedy finds.  The parity of rho is not free: the row's
        // longest run of consecutive internal ids must keep 16-byte aligned bulk copies.
        const size_t nw = wins.size();
        auto win_of = [&](int32_t row) -> int32_t {
            return (int32_t)(std::lower_bound(wins.begin(), wins.end(), row,
                                              [](const RowWin &w, int32_t r) { return w.row < r; }) - wins.begin());
        };
        std::vector<int32_t> best_of(nw, -1);
        for (size_t k = 0; k < nw; k++) {
            const RowWin &w = wins[k];
            const int32_t k0 = std::max(0, w.lo - ra.xs[w.row]);
            const int32_t k1 = std::min(row_len(w.row), w.hi - ra.xs[w.row] + 1);
            if (k1 <= k0) continue;                                           // nothing but holes
            const int32_t e0 = ra.row_start[w.row] + k0, e1 = ra.row_start[w.row] + k1;
            int32_t best_len = 0;
            for (int32_t i = e0; i < e1;) {
                int32_t j = i + 1;
                while (j < e1 && p->perm[j] == p->perm[j - 1] + 1) j++;
                if (j - i > best_len) { best_len = j - i; best_of[k] = i; }
                i = j;
            }
        }
        refs.assign(nw, std::vector<int32_t>());
        order_rows.clear();
        for (size_t g = 0; g < tsegs.size(); g++) {
            const Seg &sg = segs[tsegs[g].second];
            const int32_t r = ra.row_of[sg.first], u0 = u_of(sg.first);
            const int32_t h = (int32_t)(g / 16);
            const int32_t wo = win_of(r);
            if (order_rows.empty() || order_rows.back() != wo) order_rows.push_back(wo);
            refs[wo].push_back((h * 7 + 0) * 16 + (u0 & 15));
            for (int x = 0; x < NSTREAM; x++)
                if (sg.pat.row[x] >= 0)
                    refs[win_of(sg.pat.row[x])].push_back((h * 7 + 1 + x) * 16 + ((u0 + sg.pat.sh[x]) & 15));
        }
        {
            std::vector<uint8_t> seen(nw, 0);
            for (const int32_t k : order_rows) seen[k] = 1;
            for (size_t k = 0; k < nw; k++) if (!seen[k]) order_rows.push_back((int32_t)k);
        }
        bank_cnt.assign(((tsegs.size() + 15) / 16 + 1) * 7 * 16, 0);
        std::vector<int32_t> rho(nw, 0);
        for (const int32_t k : order_rows) {
            const RowWin &w = wins[k];
            int want_par = -1;
            if (best_of[k] >= 0) want_par = (p->perm[best_of[k]] - u_of(best_of[k])) & 1;
            int32_t best_rho = want_par < 0 ? 0 : want_par, best_cost = INT32_MAX;
            for (int32_t c = 0; c < 16; c++) {
                if (want_par >= 0 && (c & 1) != want_par) continue;
                int32_t cost = 0;
                for (const int32_t rf : refs[k]) {
                    const int32_t grp = rf >> 4, res = (rf + c) & 15;
                    cost += bank_cnt[grp * 16 + res];
                    bank_cnt[grp * 16 + res]++;                               // conflicts inside the row count too
                }
                for (const int32_t rf : refs[k]) bank_cnt[(rf >> 4) * 16 + ((rf + c) & 15)]--;
                if (cost < best_cost) { best_cost = cost; best_rho = c; }
            }
            rho[k] = best_rho;
            for (const int32_t rf : refs[k]) bank_cnt[(rf >> 4) * 16 + ((rf + best_rho) & 15)]++;
            (void)w;
        }
        // chain the rows so that each starts where the previous one ended (mod 16): little padding
        int32_t next = ZH;
        {
            std::vector<uint8_t> placed(nw, 0);
            for (size_t n_placed = 0; n_placed < nw; n_placed++) {
                int32_t pick = -1, pick_pad = 99;
                for (size_t k = 0; k < nw; k++) {
                    if (placed[k]) continue;
                    const int32_t pad = (((rho[k] + wins[k].lo - next) % 16) + 16) % 16;
                    if (pad < pick_pad) { pick_pad = pad; pick = (int32_t)k; }
                    if (pad == 0) break;
                }
                RowWin &w = wins[pick];
                placed[pick] = 1;
                next += pick_pad;
                w.base = next;
                if (best_of[pick] >= 0) {
                    const int32_t k0 = std::max(0, w.lo - ra.xs[w.row]);
                    const int32_t k1 = std::min(row_len(w.row), w.hi - ra.xs[w.row] + 1);
                    const int32_t e0 = ra.row_start[w.row] + k0, e1 = ra.row_start[w.row] + k1;
                    for (int32_t i = e0; i < e1;) {
                        int32_t j = i + 1;
                        while (j < e1 && p->perm[j] == p->perm[j - 1] + 1) j++;
                        const int32_t q = p->perm[i];
                        emit(q, w.base + (u_of(i) - w.lo), j - i);
                        if (q < t0 || q >= t1) halo_total += j - i;
                        i = j;
                    }
                }
                next += w.hi - w.lo + 1;
            }
        }
        next += R + 2;                              // masked lanes of short segments read past their row
        if (next > sp.max_img_points || next > 65535) {
            if (getenv("EUT_TRACE"))
                fprintf(stderr, "[eut] segment planner: tile %d (%d points) needs an image of %d cells > %d\n",
                        t, t1 - t0, next, sp.max_img_points);
            return EUT_ERR_UNSUPPORTED;
        }
        auto cell_of = [&](int32_t row, int32_t u) -> int32_t {
            const auto it = std::lower_bound(wins.begin(), wins.end(), row,
                                             [](const RowWin &w, int32_t r) { return w.row < r; });
            return it->base + (u - it->lo);
        };
        const int32_t out0 = t0 & 1;                // result-buffer cell parity = global parity
        int32_t *m = &S.meta[(size_t)t * 12];
        m[0] = t0; m[1] = t1 - t0;
        m[2] = (int32_t)S.copies.size(); m[3] = (int32_t)bulk.size();
        S.copies.insert(S.copies.end(), bulk.begin(), bulk.end());
        m[4] = (int32_t)S.copies.size(); m[5] = (int32_t)single.size();
        S.copies.insert(S.copies.end(), single.begin(), single.end());
        int32_t tx = 0;
        for (const int2 &b : bulk) tx += (b.y >> 16) * 8;
        m[6] = tx; m[7] = next;
        m[8] = (int32_t)(S.segs.size() / 8); m[9] = (int32_t)tsegs.size();
        m[10] = (int32_t)(S.gens.size() / 16); m[11] = (int32_t)tgens.size();
        for (const auto &ts : tsegs) {
            const Seg &sg = segs[ts.second];
            const int32_t r = ra.row_of[sg.first], u0 = u_of(sg.first);
            S.segs.push_back((uint16_t)cell_of(r, u0));
            for (int x = 0; x < NSTREAM; x++)
                S.segs.push_back(sg.pat.row[x] < 0 ? (uint16_t)0 : (uint16_t)cell_of(sg.pat.row[x], u0 + sg.pat.sh[x]));
            S.segs.push_back((uint16_t)((p->perm[sg.first] - t0 + out0) | (sg.len << 12)));
            S.seg_points += sg.len;
        }
        for (const int32_t i : tgens) {
            const int d = (int)(p->row_ptr[i + 1] - p->row_ptr[i]);
            for (int k = 0; k < 12; k++) {
                uint16_t c = 0;
                if (k < d) { const int32_t n = p->col[p->row_ptr[i] + k]; c = (uint16_t)cell_of(ra.row_of[n], u_of(n)); }
                S.gens.push_back(c);
            }
            S.gens.push_back((uint16_t)cell_of(ra.row_of[i], u_of(i)));
            S.gens.push_back((uint16_t)(p->perm[i] - t0 + out0));
            S.gens.push_back(0); S.gens.push_back(0);
        }
        S.img_points = std::max(S.img_points, next);
        S.max_tile_points = std::max(S.max_tile_points, t1 - t0);
        S.max_tile_segs = std::max(S.max_tile_segs, (int)tsegs.size());
        S.max_gen = std::max(S.max_gen, (int)tgens.size());
        S.max_bulk = std::max(S.max_bulk, (int)bulk.size());
        S.max_single = std::max(S.max_single, (int)single.size());
    }
    S.halo_ratio = (double)halo_total / (double)N;
    S.n_copies = (int64_t)S.copies.size();
    S.n_segs = (int64_t)(S.segs.size() / 8);
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
