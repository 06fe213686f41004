// extern "C" surface of libeutropia_b200.so (include/eutropia_b200.h) and the one-shot tier that
// reproduces the two Rcpp entry points call for call.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <atomic>
#include <mutex>
#include <thread>

#include <functional>
#include <future>

#include "common.h"

using namespace eut;

// Replays the tables the segment kernel indexes with (tile meta records, copy lists, segment and
// generic-point records): the image of every tile (image cell -> internal position) from its copy
// lists, then per point the cells its summation chain reads, in the kernel's order
// A A B B C O- O+ D E E F F (stencil_seg.cu), translated to reference ids.
static int decode_seg_tables(const eut_plan *p, const std::vector<int32_t> &meta, const std::vector<int2> &copies,
                             const std::vector<uint16_t> &segs, const std::vector<uint16_t> &gens,
                             int64_t *row_ptr, int32_t *col)
{
    const int64_t N = p->N;
    const int R = p->segs.R;
    const int nt = p->n_tiles;
    std::vector<int32_t> nbr((size_t)12 * p->Np, -1);          // [position][slot], reference ids (0-based)
    std::vector<uint8_t> deg(p->Np, 0), done(p->Np, 0);
    std::vector<int32_t> img;
    for (int t = 0; t < nt; t++) {
        const int32_t *m = meta.data() + (size_t)12 * t;
        const int32_t t0 = m[0], cnt = m[1], b0 = m[2], nb = m[3], s0 = m[4], ns = m[5], cells = m[7];
        const int32_t g0 = m[8], ng = m[9], x0 = m[10], nx = m[11];
        img.assign((size_t)cells + 1, -1);                     // -1: the copies never write the cell (zero header, holes)
        for (int32_t k = b0; k < b0 + nb; k++) {
            const int32_t g = copies[k].x, so = copies[k].y & 0xffff, n = (int32_t)((uint32_t)copies[k].y >> 16);
            for (int32_t j = 0; j < n; j++)
                if (so + j <= cells && g + j < p->Np) img[so + j] = g + j;
        }
        for (int32_t k = s0; k < s0 + ns; k++)
            if (copies[k].y >= 0 && copies[k].y <= cells) img[copies[k].y] = copies[k].x;
        const int head = t0 & 1;
        auto add = [&](int64_t q, int64_t cell) {
            if (cell < 0 || cell > cells) return;
            const int32_t src = img[cell];
            if (src < 0 || p->iperm[src] < 0) return;          // absent neighbour: the kernel adds +0.0
            if (deg[q] < 12) nbr[(size_t)q * 12 + deg[q]] = p->iperm[src];
            deg[q]++;
        };
        for (int32_t s = g0; s < g0 + ng; s++) {
            const uint16_t *r = segs.data() + (size_t)24 * s;
            const int len = r[7] >> 12, out = r[7] & 0xfff;
            auto pair_cell = [&](int col, int red, int k) -> int64_t {   // cell k (0..R) of a pair stream
                if (k == 0) return r[red] / 8;
                if (k == R) return r[red + 1] / 8;
                return r[col] / 8 + k;
            };
            for (int j = 0; j < len; j++) {
                const int64_t q = (int64_t)t0 + out + j - head;
                if (q < 0 || q >= p->Np || done[q]) { set_error("eut_plan_get_csr: segment tables produce position %lld twice or out of range", (long long)q); return EUT_ERR_UNSUPPORTED; }
                done[q] = 1;
                add(q, pair_cell(1, 10, j)); add(q, pair_cell(1, 10, j + 1));       // A A
                add(q, pair_cell(2, 12, j)); add(q, pair_cell(2, 12, j + 1));       // B B
                add(q, r[3] / 8 + j);                                                  // C
                add(q, j == 0 ? r[8] / 8 : r[0] / 8 + j - 1);                          // O-
                add(q, j == R - 1 ? r[9] / 8 : r[0] / 8 + j + 1);                      // O+
                add(q, r[4] / 8 + j);                                                  // D
                add(q, pair_cell(5, 14, j)); add(q, pair_cell(5, 14, j + 1));       // E E
                add(q, pair_cell(6, 16, j)); add(q, pair_cell(6, 16, j + 1));       // F F
                if (img[r[0] / 8 + j] != q) { set_error("eut_plan_get_csr: segment %d does not read its own point", s); return EUT_ERR_UNSUPPORTED; }
            }
        }
        for (int32_t x = x0; x < x0 + nx; x++) {
            const uint16_t *r = gens.data() + (size_t)16 * x;
            const int64_t q = (int64_t)t0 + r[13] - head;
            if (q < 0 || q >= p->Np || done[q]) { set_error("eut_plan_get_csr: generic tables produce position %lld twice or out of range", (long long)q); return EUT_ERR_UNSUPPORTED; }
            done[q] = 1;
            for (int k = 0; k < 12; k++)
                if (r[k]) add(q, r[k]);
            if (img[r[12]] != q) { set_error("eut_plan_get_csr: generic point %d does not read its own cell", x); return EUT_ERR_UNSUPPORTED; }
        }
        (void)cnt;
    }
    row_ptr[0] = 0;
    for (int64_t i = 0; i < N; i++) {
        const int64_t q = p->perm[i];
        if (!done[q] && !p->dead.empty() && p->dead[i]) { row_ptr[i + 1] = row_ptr[i]; deg[q] = 0; continue; }   // a ghost: no work by design
        if (!done[q] || deg[q] > 12) { set_error("eut_plan_get_csr: point %lld is not produced by any segment", (long long)i + 1); return EUT_ERR_UNSUPPORTED; }
        row_ptr[i + 1] = row_ptr[i] + deg[q];
    }
    if (row_ptr[N] != p->E) { set_error("eut_plan_get_csr: the segment tables hold %lld edges, mat.in has %lld", (long long)row_ptr[N], (long long)p->E); return EUT_ERR_UNSUPPORTED; }
    for (int64_t i = 0; i < N; i++) {
        const int64_t q = p->perm[i];
        for (int k = 0; k < deg[q]; k++) col[row_ptr[i] + k] = nbr[(size_t)q * 12 + k] + 1;
    }
    return EUT_OK;
}

extern "C" {

const char *eut_last_error(void) { return last_error_cstr(); }
int eut_abi_version(void) { return 1000; }

int eut_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

// ---- plan ------------------------------------------------------------------------------------
int eut_plan_create(eut_plan **out, const int32_t *from, const int32_t *to, int64_t n_edges,
                    const int32_t *nn_id, const int32_t *nn_cnt, int64_t n_nn, int64_t n_points,
                    int device, uint32_t flags)
{
    clear_error();
    if (!out) { set_error("eut_plan_create: out is NULL"); return EUT_ERR_ARG; }
    *out = nullptr;
    eut_plan *p = new eut_plan();
    int rc = plan_build(p, from, to, n_edges, nn_id, nn_cnt, n_nn, n_points, device, flags);
    if (rc != EUT_OK) {
        plan_free(p);
        delete p;
        return rc;
    }
    *out = p;
    return EUT_OK;
}

int eut_build_dcm(const double *pts, int64_t ld, int64_t n_points, double field_size, int device,
                  int32_t *mat_in, int64_t cap, int64_t *n_edges, int32_t *mat_out)
{
    clear_error();
    if (n_points < 0 || cap < 0 || !n_edges || (n_points > 0 && (!pts || !mat_out || ld < n_points)) || (cap > 0 && !mat_in) ||
        !(field_size > 0)) {
        set_error("eut_build_dcm: NULL pointer, negative size or non-positive field size");
        return EUT_ERR_ARG;
    }
    return mesh_build_dcm(pts, ld, n_points, field_size, device, mat_in, cap, n_edges, mat_out);
}

int eut_plan_destroy(eut_plan *plan)
{
    if (!plan) return EUT_OK;
    plan_free(plan);
    delete plan;
    return EUT_OK;
}

int eut_plan_get_info(const eut_plan *p, eut_plan_info *info)
{
    if (!p || !info) { set_error("eut_plan_get_info: NULL argument"); return EUT_ERR_ARG; }
    info->n_points = p->N; info->n_edges = p->E; info->n_padded = p->Np;
    info->max_in_degree = p->max_in; info->ell_width = p->ell_w; info->order = p->order;
    info->regular_outflow = p->regular_out; info->device = p->device; info->n_tiles = p->n_tiles;
    info->device_bytes = p->device_bytes; info->build_ms = p->build_ms;
    info->tiled = p->tiled ? 1 : 0; info->img_points = p->tiles.img_points;
    info->max_tile_points = p->tiles.max_tile_points; info->max_bulk = p->tiles.max_bulk;
    info->max_single = p->tiles.max_single; info->n_copies = (int32_t)p->tiles.n_copies;
    info->halo_ratio = p->tiles.halo_ratio;
    const auto &S = p->segs;
    info->segmented = p->segmented ? 1 : 0; info->seg_points = S.R;
    info->max_tile_segs = S.max_tile_segs; info->max_tile_generic = S.max_gen;
    info->n_segments = S.n_segs; info->n_generic = S.n_gens; info->n_seg_covered = S.seg_points;
    if (p->segmented) {
        info->img_points = S.img_points; info->max_tile_points = S.max_tile_points;
        info->max_bulk = S.max_bulk; info->max_single = S.max_single;
        info->n_copies = (int32_t)S.n_copies; info->halo_ratio = S.halo_ratio;
    }
    return EUT_OK;
}

// The map is read back FROM THE DEVICE tables and translated to reference ids, so the comparison in
// tests/ covers what the kernels actually index with.
int eut_plan_get_csr(const eut_plan *p, int64_t *row_ptr, int32_t *col)
{
    clear_error();
    if (!p || !row_ptr || (!col && p->E > 0)) { set_error("eut_plan_get_csr: NULL argument"); return EUT_ERR_ARG; }
    if (p->host_only && !p->segmented) { set_error("eut_plan_get_csr: host-only plan has no device tables"); return EUT_ERR_UNSUPPORTED; }
    if (!p->host_only) EUT_TRY(ensure_device(p->device));
    const int64_t N = p->N;
    if (p->segmented) {
        // The segment kernel does not index with the ELL tables: read ITS tables back from the device
        // (a host-only plan: the planner's host copies) and replay them.
        const SegPlanHost &S = p->segs;
        if (p->host_only) return decode_seg_tables(p, S.meta, S.copies, S.segs, S.gens, row_ptr, col);
        const int nt = p->n_tiles;
        std::vector<int32_t> meta((size_t)12 * nt);
        std::vector<int2> copies((size_t)S.n_copies);
        std::vector<uint16_t> segs((size_t)24 * S.n_segs), gens((size_t)16 * S.n_gens);
        if (nt) EUT_CUDA(cudaMemcpy(meta.data(), p->d_seg_meta, meta.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (!copies.empty()) EUT_CUDA(cudaMemcpy(copies.data(), p->d_copies, copies.size() * sizeof(int2), cudaMemcpyDeviceToHost));
        if (!segs.empty()) EUT_CUDA(cudaMemcpy(segs.data(), p->d_segs, segs.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
        if (!gens.empty()) EUT_CUDA(cudaMemcpy(gens.data(), p->d_gens, gens.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
        return decode_seg_tables(p, meta, copies, segs, gens, row_ptr, col);
    }
    if (p->ell_w > 0) {
        std::vector<int32_t> nbr((size_t)p->ell_w * p->Np);
        std::vector<uint8_t> deg(p->Np);
        EUT_CUDA(cudaMemcpy(nbr.data(), p->d_nbr, nbr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        EUT_CUDA(cudaMemcpy(deg.data(), p->d_deg, deg.size(), cudaMemcpyDeviceToHost));
        row_ptr[0] = 0;
        for (int64_t i = 0; i < N; i++) row_ptr[i + 1] = row_ptr[i] + deg[p->perm[i]];
        for (int64_t i = 0; i < N; i++) {
            const int64_t q = p->perm[i];
            for (int k = 0; k < deg[q]; k++)
                col[row_ptr[i] + k] = p->iperm[nbr[(size_t)k * p->Np + q]] + 1;
        }
    } else {
        std::vector<int64_t> rp(N + 1);
        std::vector<int32_t> cl(p->E);
        EUT_CUDA(cudaMemcpy(rp.data(), p->d_row_ptr, rp.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
        if (p->E) EUT_CUDA(cudaMemcpy(cl.data(), p->d_col, cl.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        row_ptr[0] = 0;
        for (int64_t i = 0; i < N; i++) {
            const int64_t q = p->perm[i];
            row_ptr[i + 1] = row_ptr[i] + (rp[q + 1] - rp[q]);
        }
        for (int64_t i = 0; i < N; i++) {
            const int64_t q = p->perm[i];
            for (int64_t k = rp[q], o = row_ptr[i]; k < rp[q + 1]; k++, o++) col[o] = p->iperm[cl[k]] + 1;
        }
    }
    return EUT_OK;
}

int eut_plan_get_tiles(const eut_plan *p, int32_t *tile_meta, int32_t *copies, uint16_t *loc)
{
    if (!p || !p->tiled) { set_error("eut_plan_get_tiles: plan has no tile tables"); return EUT_ERR_ARG; }
    if (!p->host_only) { set_error("eut_plan_get_tiles: only for EUT_PLAN_HOST_ONLY plans"); return EUT_ERR_UNSUPPORTED; }
    const auto &T = p->tiles;
    if (tile_meta) std::copy(T.meta.begin(), T.meta.end(), tile_meta);
    if (copies) memcpy(copies, T.copies.data(), T.copies.size() * sizeof(int2));
    if (loc) std::copy(T.loc.begin(), T.loc.end(), loc);
    return EUT_OK;
}

int eut_plan_get_segs(const eut_plan *p, int32_t *seg_meta, int32_t *copies, uint16_t *segs, uint16_t *gens)
{
    if (!p || !p->segmented) { set_error("eut_plan_get_segs: plan has no segment tables"); return EUT_ERR_ARG; }
    if (!p->host_only) { set_error("eut_plan_get_segs: only for EUT_PLAN_HOST_ONLY plans"); return EUT_ERR_UNSUPPORTED; }
    const auto &S = p->segs;
    if (seg_meta) std::copy(S.meta.begin(), S.meta.end(), seg_meta);
    if (copies) memcpy(copies, S.copies.data(), S.copies.size() * sizeof(int2));
    if (segs) std::copy(S.segs.begin(), S.segs.end(), segs);
    if (gens) std::copy(S.gens.begin(), S.gens.end(), gens);
    return EUT_OK;
}

int eut_plan_get_perm(const eut_plan *p, int32_t *perm)
{
    if (!p || (!perm && p->N > 0)) { set_error("eut_plan_get_perm: NULL argument"); return EUT_ERR_ARG; }
    std::copy(p->perm.begin(), p->perm.end(), perm);
    return EUT_OK;
}

// ---- field -----------------------------------------------------------------------------------
int eut_field_create(eut_field **out, eut_plan *plan, int64_t n_cols)
{
    clear_error();
    if (out) *out = nullptr;
    return field_create(out, plan, n_cols);
}
int eut_field_destroy(eut_field *f) { return field_destroy(f); }

int eut_field_upload(eut_field *f, const double *host, int64_t ld, const int32_t *cols, int64_t n_cols)
{
    clear_error();
    if (!f) { set_error("eut_field_upload: field is NULL"); return EUT_ERR_ARG; }
    EUT_TRY(field_upload(f, host, ld, cols, n_cols));
    // the caller may free / overwrite `host` right after this returns
    EUT_CUDA(cudaStreamSynchronize(f->copy_stream));
    return EUT_OK;
}

int eut_field_download(eut_field *f, double *host, int64_t ld, const int32_t *cols, int64_t n_cols)
{
    clear_error();
    if (!f) { set_error("eut_field_download: field is NULL"); return EUT_ERR_ARG; }
    EUT_TRY(field_download_async(f, host, ld, cols, n_cols));
    return field_sync_down(f);
}

int eut_field_record(eut_field *f, double *host, int64_t ld, const int32_t *cols, int64_t n_cols, int digits)
{
    clear_error();
    if (!f) { set_error("eut_field_record: field is NULL"); return EUT_ERR_ARG; }
    if (digits < 1 || digits > 15) { set_error("eut_field_record: digits = %d, expected 1..15", digits); return EUT_ERR_ARG; }
    EUT_TRY(field_sync_down(f));                       // no download of this field in flight with the other setting
    double p10 = 1.0;
    for (int k = 0; k < digits; k++) p10 *= 10.0;      // exact up to 1e22
    f->round_p10 = p10;
    int rc = field_download_async(f, host, ld, cols, n_cols);
    const int rc2 = field_sync_down(f);
    f->round_p10 = 0.0;
    return rc != EUT_OK ? rc : rc2;
}

int eut_field_diffuse(eut_field *f, const int32_t *niter, const int32_t *indvar, int64_t n_indvar)
{
    clear_error();
    if (!f) { set_error("eut_field_diffuse: field is NULL"); return EUT_ERR_ARG; }
    return field_diffuse(f, niter, indvar, n_indvar);
}

int eut_field_sync(eut_field *f)
{
    if (!f) return EUT_ERR_ARG;
    EUT_TRY(ensure_device(f->plan->device));
    EUT_CUDA(cudaStreamSynchronize(f->copy_stream));
    EUT_CUDA(cudaStreamSynchronize(f->stream));
    EUT_TRY(field_sync_down(f));
    return EUT_OK;
}

void *eut_field_stream(eut_field *f) { return f ? (void *)f->stream : nullptr; }

int eut_field_set_stream(eut_field *f, void *cuda_stream)
{
    if (!f) return EUT_ERR_ARG;
    EUT_TRY(eut_field_sync(f));
    if (f->own_stream && f->stream) cudaStreamDestroy(f->stream);
    f->stream = (cudaStream_t)cuda_stream;
    f->own_stream = false;
    return EUT_OK;
}

int64_t eut_field_launch_count(const eut_field *f) { return f ? f->launches.load() : (int64_t)0; }

int eut_field_set_option(eut_field *f, const char *name, int64_t value)
{
    if (!f || !name) return EUT_ERR_ARG;
    std::string n(name);
    if (n == "cpt") f->opt_cpt = (int)value;
    else if (n == "variant") f->opt_variant = (int)value;
    else if (n == "group_cols") f->opt_group_cols = (int)value;
    else if (n == "nbuf") f->opt_nbuf = (int)value;
    else if (n == "debug") f->opt_debug = (int)value;
    else if (n == "shape") f->opt_shape = (int)value;
    else if (n == "phase") f->opt_phase = (int)value;
    else { set_error("eut_field_set_option: unknown option '%s'", name); return EUT_ERR_ARG; }
    return EUT_OK;
}

int eut_field_column_stats(eut_field *f, double *col_min, double *col_max, double *col_mean)
{
    clear_error();
    if (!f) return EUT_ERR_ARG;
    EUT_TRY(ensure_device(f->plan->device));
    EUT_TRY(launch_column_stats(f, f->d_cur));
    std::vector<double> h(3 * std::max<int64_t>(f->C, 1));
    EUT_CUDA(cudaMemcpyAsync(h.data(), f->d_stats, 3 * f->C * sizeof(double), cudaMemcpyDeviceToHost,
                             f->stream));
    EUT_CUDA(cudaStreamSynchronize(f->stream));
    for (int64_t c = 0; c < f->C; c++) {
        if (col_min) col_min[c] = h[c];
        if (col_max) col_max[c] = h[f->C + c];
        if (col_mean) col_mean[c] = h[2 * f->C + c];
    }
    return EUT_OK;
}

int eut_field_fill_column(eut_field *f, int32_t col, double value)
{
    clear_error();
    if (!f) return EUT_ERR_ARG;
    if (col < 1 || col > f->C) { set_error("eut_field_fill_column: column %d out of bounds", col); return EUT_ERR_INDEX; }
    EUT_TRY(ensure_device(f->plan->device));
    return launch_fill_column(f, col - 1, value);
}

int eut_field_exoenzyme_catalysis(eut_field *conc, eut_field *exo, int32_t exo_col, double kcat, double km,
                                  const int32_t *met_cols, const double *stoich, int32_t n_mets,
                                  const uint8_t *is_constant, double delta_time)
{
    clear_error();
    if (!conc || !exo || conc->plan != exo->plan || !met_cols || !stoich) {
        set_error("eut_field_exoenzyme_catalysis: NULL argument, or the two fields do not share a plan");
        return EUT_ERR_ARG;
    }
    if (exo_col < 1 || exo_col > exo->C) { set_error("index out of bounds: exoenzyme column %d of %lld", exo_col, (long long)exo->C); return EUT_ERR_INDEX; }
    if (n_mets < 1 || n_mets > max_exo_mets()) { set_error("eut_field_exoenzyme_catalysis: %d metabolites (1..%d supported)", n_mets, max_exo_mets()); return EUT_ERR_UNSUPPORTED; }
    for (int32_t i = 0; i < n_mets; i++)
        if (met_cols[i] < 1 || met_cols[i] > conc->C) {
            // match(exec@mets, compounds) = NA in the reference -> subscript error
            set_error("index out of bounds: metabolite column %d of %lld", met_cols[i], (long long)conc->C);
            return EUT_ERR_INDEX;
        }
    EUT_TRY(ensure_device(conc->plan->device));
    return field_exoenzyme_catalysis(conc, exo, exo_col, kcat, km, met_cols, stoich, n_mets, is_constant, delta_time);
}

int eut_field_scale_column(eut_field *f, int32_t col, double factor)
{
    clear_error();
    if (!f) return EUT_ERR_ARG;
    if (col < 1 || col > f->C) { set_error("eut_field_scale_column: column %d out of bounds", col); return EUT_ERR_INDEX; }
    EUT_TRY(ensure_device(f->plan->device));
    return field_scale_column(f, col - 1, factor);
}

}  // extern "C"
