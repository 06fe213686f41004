/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see the header of diffchange_oracle.c: same rules).
 *
 * PARITY UNPINNED: no golden vectors exist in the reference for these steps, and three of the
 * arithmetic providers are un-vendored, un-pinned R packages:
 *   sf/GEOS  st_distance        -> restated as planar xy distance sqrt(dx*dx+dy*dy), Z ignored
 *                                  (GEOS geom::Coordinate::distance); dist_mode=1 gives xyz
 *   base R   sum / colSums      -> long double accumulator, rounded once (R's LDOUBLE rsum)
 *   data.table  setkeyv / by-group sum -> stable ordering; groups summed in table (row) order
 *
 * What is restated (citations relative to /root/reference):
 *   eut_oracle_cellmap   <- R/run_simulation.R:165-168 (grid keyed z,x,y), :217-230 (box prefilter),
 *                           :591-624 (get_cellFieldEnv_ex: distance, filter, acc.prop)
 *   eut_oracle_claim     <- R/run_simulation.R:236-241 (oversubscribed fields: acc^e rescale)
 *   eut_oracle_gather    <- R/run_simulation.R:249 + R/scavenge_compounds.R:17-22
 *   eut_oracle_scatter   <- R/run_simulation.R:689-740 (triplets), :309-319 (sum, add, clamp)
 *
 * Conventions: column-major matrices, 1-based field ids and compound columns, ragged per-cell
 * lists in CSR form (cell_ptr[M+1], 0-based offsets into the entry arrays).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- stable order of field points by (z, x, y): setkeyv(gridDT, c("z","x","y")) :168 ---- */
typedef struct { double z, x, y; int field; } grid_row;

static int grid_cmp(const void *a, const void *b)
{
    const grid_row *p = (const grid_row *)a, *q = (const grid_row *)b;
    if (p->z != q->z) return p->z < q->z ? -1 : 1;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    return p->field < q->field ? -1 : (p->field > q->field); /* stable: original row order */
}

/* Cell -> field lists.
 *   pts        n_pts x 3 column-major (x | y | z)   (st_coordinates(field.pts), :165-166)
 *   cx,cy,csize,cscv  per cell: position, diameter, scavenge distance (:219-222)
 *   dist_mode  0 = planar xy distance (GEOS; default), 1 = xyz distance
 * Output (entries of cell i at [cell_ptr[i], cell_ptr[i+1])), in gridDT key order (z,x,y):
 *   field_id (1-based), field_dist, acc_prop = min(1, (-1/scv) * (d - (size/2+scv)))  (:618-619)
 * cell_ptr is always filled; entry arrays only if the total fits in `capacity`.
 * Returns the total number of entries, or -1 on allocation failure.
 */
long eut_oracle_cellmap(const double *pts, size_t n_pts,
                        const double *cx, const double *cy, const double *csize,
                        const double *cscv, size_t n_cells, int dist_mode,
                        long *cell_ptr, int *field_id, double *field_dist, double *acc_prop,
                        long capacity)
{
    grid_row *g = (grid_row *)malloc((n_pts ? n_pts : 1) * sizeof(grid_row));
    if (!g) return -1;
    for (size_t i = 0; i < n_pts; i++) {
        g[i].x = pts[i]; g[i].y = pts[n_pts + i]; g[i].z = pts[2 * n_pts + i];
        g[i].field = (int)i + 1;
    }
    qsort(g, n_pts, sizeof(grid_row), grid_cmp);

    long total = 0;
    for (size_t c = 0; c < n_cells; c++) {
        cell_ptr[c] = total;
        const double R = csize[c] / 2 + cscv[c];
        const double inv = -1 / cscv[c];
        for (size_t i = 0; i < n_pts; i++) {
            /* box prefilter :225-227 */
            if (!(fabs(g[i].x - cx[c]) <= R)) continue;
            if (!(fabs(g[i].y - cy[c]) <= R)) continue;
            if (!(fabs(g[i].z - 0) <= R)) continue;
            double dx = g[i].x - cx[c], dy = g[i].y - cy[c];
            double d2 = dx * dx + dy * dy;
            if (dist_mode == 1) { double dz = g[i].z - 0; d2 = d2 + dz * dz; }
            double d = sqrt(d2);
            if (!(d <= R)) continue;                     /* :612 */
            double acc = inv * (d - R);                  /* :618 */
            if (acc > 1) acc = 1;                        /* :619 */
            if (total < capacity) {
                field_id[total] = g[i].field;
                field_dist[total] = d;
                acc_prop[total] = acc;
            }
            total++;
        }
    }
    cell_ptr[n_cells] = total;
    free(g);
    return total;
}

/* Oversubscribed fields (R/run_simulation.R:236-241), in place on acc_prop:
 *   claim[f] = sum(acc) over all entries of field f            (table order)
 *   where claim[f] > 1:  tmp = acc^exp(1);  tmp = tmp / sum_f(tmp) * claim[f];  acc = tmp / claim[f]
 */
int eut_oracle_claim(const int *field_id, double *acc_prop, long n_entries, size_t n_pts)
{
    long double *claim = (long double *)calloc(n_pts ? n_pts : 1, sizeof(long double));
    long double *tsum = (long double *)calloc(n_pts ? n_pts : 1, sizeof(long double));
    if (!claim || !tsum) { free(claim); free(tsum); return 2; }
    const double e = exp(1.0);
    for (long k = 0; k < n_entries; k++) claim[field_id[k] - 1] += acc_prop[k];
    for (long k = 0; k < n_entries; k++) {
        int f = field_id[k] - 1;
        if ((double)claim[f] > 1) tsum[f] += pow(acc_prop[k], e);
    }
    for (long k = 0; k < n_entries; k++) {
        int f = field_id[k] - 1;
        double cl = (double)claim[f];
        if (cl > 1) {
            double tmp = pow(acc_prop[k], e);
            tmp = tmp / (double)tsum[f] * cl;
            acc_prop[k] = tmp / cl;
        }
    }
    free(claim); free(tsum);
    return 0;
}

/* Gather (R/run_simulation.R:249, R/scavenge_compounds.R:17-22).
 *   conc   n_pts x n_cpd column-major (mM)
 * For cell i, entry k (field f), compound c:
 *   fmolPerField = ((conc[f,c] / 1000) * fieldVol) * acc_prop[k]          (:17-18)
 *   fmol[i,c]    = sum_k fmolPerField  (R sum: long double accumulator)    (:22)
 * Outputs: fmol  n_cells x n_cpd column-major;  fpf (optional, may be NULL) n_entries x n_cpd
 * column-major (the per-entry values, reused by the scatter).
 */
int eut_oracle_gather(const double *conc, size_t n_pts, size_t n_cpd,
                      const long *cell_ptr, const int *field_id, const double *acc_prop,
                      size_t n_cells, double field_vol, double *fmol, double *fpf)
{
    long n_entries = cell_ptr[n_cells];
    for (size_t c = 0; c < n_cpd; c++) {
        const double *col = conc + c * n_pts;
        for (size_t i = 0; i < n_cells; i++) {
            long double s = 0;
            for (long k = cell_ptr[i]; k < cell_ptr[i + 1]; k++) {
                int f = field_id[k] - 1;
                if (f < 0 || (size_t)f >= n_pts) return 1;
                double v = ((col[f] / 1000) * field_vol) * acc_prop[k];
                if (fpf) fpf[c * (size_t)n_entries + k] = v;
                s += v;
            }
            fmol[c * n_cells + i] = (double)s;
        }
    }
    return 0;
}

/* Scatter + clamp (R/run_simulation.R:689-740 per cell, :309-319 aggregate/apply).
 * Per cell i the exchange list ex_ptr[i]..ex_ptr[i+1]: compound column ex_cpd (1-based) and
 * ex_flx = flux * cMass * deltaTime in fmol, non-zero (:682-687).
 *   uptake (ex_flx < 0), every entry k of the cell (:693-708):
 *       accmat = fmolPerField[k,c] / colSum_c ;  d = ((accmat * ex) / 1e12) / (fieldVol / 1e15)
 *   production (ex_flx > 0) (:729-734):
 *       frac_k = (max_k dist - dist_k) / sum_k(max dist - dist) ;  d = frac_k * ((ex/1e12)/(fieldVol/1e15))
 * Aggregate: rows concatenated cell by cell (I.up then I.pd), summed by (field, compound) in row
 * order (:309-310); compounds flagged constant are dropped (:311-312); NaN sums -> 0 (:315);
 * conc[f,c] += sum (:318); touched entries < 1e-12 -> 0 (:319).  conc is updated in place.
 */
int eut_oracle_scatter(double *conc, size_t n_pts, size_t n_cpd, const unsigned char *is_constant,
                       const long *cell_ptr, const int *field_id, const double *field_dist,
                       const double *acc_prop, size_t n_cells, double field_vol,
                       const long *ex_ptr, const int *ex_cpd, const double *ex_flx)
{
    /* dense accumulators are fine for an oracle: delta (NaN-propagating sum) + touched flags */
    double *delta = (double *)calloc((n_pts * n_cpd > 0) ? n_pts * n_cpd : 1, sizeof(double));
    unsigned char *touched = (unsigned char *)calloc((n_pts * n_cpd > 0) ? n_pts * n_cpd : 1, 1);
    if (!delta || !touched) { free(delta); free(touched); return 2; }
    const double vol_l = field_vol / 1e15;

    for (size_t i = 0; i < n_cells; i++) {
        long k0 = cell_ptr[i], k1 = cell_ptr[i + 1];
        /* field.frac (:729-730) */
        double dmax = -INFINITY;
        for (long k = k0; k < k1; k++) if (field_dist[k] > dmax) dmax = field_dist[k];
        long double fsum = 0;
        for (long k = k0; k < k1; k++) fsum += (dmax - field_dist[k]);
        double fsum_d = (double)fsum;

        for (int pass = 0; pass < 2; pass++) {          /* 0: I.up rows, 1: I.pd rows */
            for (long e = ex_ptr[i]; e < ex_ptr[i + 1]; e++) {
                double ex = ex_flx[e];
                size_t c = (size_t)ex_cpd[e] - 1;
                if (c >= n_cpd) { free(delta); free(touched); return 1; }
                if (pass == 0 && !(ex < 0)) continue;
                if (pass == 1 && !(ex > 0)) continue;
                const double *col = conc + c * n_pts;
                long double cs = 0;
                if (pass == 0) {                         /* colSums(accmat) :696 */
                    for (long k = k0; k < k1; k++)
                        cs += ((col[field_id[k] - 1] / 1000) * field_vol) * acc_prop[k];
                }
                double cs_d = (double)cs;
                double ex_mM = (ex / 1e12) / vol_l;      /* :732 */
                for (long k = k0; k < k1; k++) {
                    size_t f = (size_t)field_id[k] - 1;
                    double d;
                    if (pass == 0) {
                        double fpf = ((col[f] / 1000) * field_vol) * acc_prop[k];
                        double acc = fpf / cs_d;         /* :696 */
                        d = ((acc * ex) / 1e12) / vol_l; /* :707-708 */
                    } else {
                        double frac = (dmax - field_dist[k]) / fsum_d;
                        d = frac * ex_mM;                /* :734 */
                    }
                    delta[c * n_pts + f] += d;
                    touched[c * n_pts + f] = 1;
                }
            }
        }
    }
    for (size_t c = 0; c < n_cpd; c++) {
        if (is_constant && is_constant[c]) continue;     /* :311-312 */
        for (size_t f = 0; f < n_pts; f++) {
            size_t o = c * n_pts + f;
            if (!touched[o]) continue;
            double d = delta[o];
            if (isnan(d)) d = 0;                         /* :315 */
            double v = conc[o] + d;                      /* :318 */
            if (v < 1e-12) v = 0;                        /* :319 */
            conc[o] = v;
        }
    }
    free(delta); free(touched);
    return 0;
}
