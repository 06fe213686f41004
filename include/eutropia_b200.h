/*
 * eutropia_b200 -- C ABI of the B200 (sm_100a) implementation of Eutropia's per-timestep compound
 * diffusion + cell<->field exchange path.
 *
 * This header is the drop-in boundary.  Every entry point takes plain pointers and sizes (no R,
 * Rcpp, Armadillo or torch types) and cites the reference interface it replaces; paths are relative
 * to the reference tree (Waschina/Eutropia 0.1.7).  The Rcpp glue a maintainer adds on the R side
 * is shown in INTEGRATION.md and shipped under rcpp/.
 *
 * Conventions (identical to the reference):
 *   - matrices are column-major, field / compound indices are 1-based;
 *   - concentrations are IEEE binary64; all arithmetic on the device is binary64 with
 *     round-to-nearest and NO fused multiply-add, in the reference's operation order.
 *
 * Two tiers:
 *   one-shot : eut_diffchange / eut_diffchange_par -- exact `diffChange` / `diffChangePar`
 *              semantics, host buffers in, host buffer out (H2D + substeps + D2H per call);
 *   session  : eut_plan_* (graph resident on a GPU), eut_field_* (field resident across rounds),
 *              eut_cells_* (cell lists, gather, scatter) -- nothing crosses PCIe per round except
 *              what the caller asks for.
 *
 * There is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 * EUT_ERR_CUDA / EUT_ERR_NO_DEVICE and says so in eut_last_error().
 */
#ifndef EUTROPIA_B200_H
#define EUTROPIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EUT_API __attribute__((visibility("default")))

/* ---- status codes -------------------------------------------------------------------------- */
enum {
    EUT_OK             = 0,
    EUT_ERR_INDEX      = 1,  /* an index is outside 1..N / 1..C: where Armadillo's bounds check
                                (no ARMA_NO_DEBUG in src/Makevars) makes the reference raise an R
                                error instead of reading out of range */
    EUT_ERR_ARG        = 2,  /* NULL pointer, negative size, shape mismatch */
    EUT_ERR_ALLOC      = 3,  /* host or device allocation failed */
    EUT_ERR_CUDA       = 4,  /* a CUDA runtime call failed; text in eut_last_error() */
    EUT_ERR_NO_DEVICE  = 5,  /* no CUDA device: this library has no CPU path */
    EUT_ERR_UNSUPPORTED= 6,
    EUT_ERR_NCCL       = 7
};

/* Message describing the last failure on the calling thread ("" if none). */
EUT_API const char *eut_last_error(void);
/* ABI version of this header (major*1000 + minor). */
EUT_API int eut_abi_version(void);
/* Number of usable CUDA devices (0 when there is no driver / no GPU). */
EUT_API int eut_device_count(void);

/* ---- one-shot tier: the two Rcpp entry points ------------------------------------------------
 *
 * eut_diffchange  replaces  diffChange(adjmat, nneighbors, conc, niter)
 *     R closure            R/RcppExports.R:4-6
 *     .Call symbol         _Eutropia_diffChange, src/RcppExports.cpp:16-27
 *     implementation       src/diffChange.cpp:6-35
 *   adjmat      n_edges x 2 doubles (from | to), cast with (int) like src/diffChange.cpp:20
 *   nneighbors  n_nn x 2 doubles (id | n_neighbors), src/diffChange.cpp:26
 *   conc_in     n_rows x n_cols; never written (R value semantics, SURVEY 8b "Ownership")
 *   niter       n_niter doubles; column i runs while (int)k < niter[i] (src/diffChange.cpp:12-14);
 *               columns >= n_niter are returned unchanged; n_niter > n_cols with a positive entry
 *               beyond n_cols -> EUT_ERR_INDEX
 *   conc_out    n_rows x n_cols, distinct from conc_in; receives the whole matrix (:34)
 *   device      CUDA ordinal
 */
EUT_API int eut_diffchange(const double *adjmat, int64_t n_edges,
                           const double *nneighbors, int64_t n_nn,
                           const double *conc_in, int64_t n_rows, int64_t n_cols,
                           const double *niter, int64_t n_niter,
                           double *conc_out, int device);

/* eut_diffchange_par  replaces  diffChangePar(adjmat, nneighbors, conc, niter, indvar, ncores)
 *     R closure            R/RcppExports.R:8-10
 *     .Call symbol         _Eutropia_diffChangePar, src/RcppExports.cpp:30-43
 *     implementation       src/diffChange.cpp:38-69
 *     live caller          diffuse_compoundsPar, R/growthEnvironment.R:385-390 and :417-422
 *   adjmat / nneighbors    int32, column-major, 1-based
 *   niter[i]               substeps of column indvar[i] (by position, src/diffChange.cpp:48)
 *   indvar                 1-based column numbers, unique (the reference's OpenMP loop races
 *                          otherwise, :44-45/:63); duplicates -> EUT_ERR_ARG here
 *   ncores                 accepted for signature compatibility; the GPU path ignores it
 */
EUT_API int eut_diffchange_par(const int32_t *adjmat, int64_t n_edges,
                               const int32_t *nneighbors, int64_t n_nn,
                               const double *conc_in, int64_t n_rows, int64_t n_cols,
                               const int32_t *niter, const int32_t *indvar, int64_t n_indvar,
                               int ncores, double *conc_out, int device);

/* eut_diffchange_par_multi: the same call on SEVERAL GPUs of the box from ONE process -- the
 * counterpart of the reference's `#pragma omp parallel for num_threads(ncores)` over the compounds
 * of one call (src/diffChange.cpp:44-45), which diffuse_compoundsPar issues once per round from the
 * main R process (R/growthEnvironment.R:385-390).  The active columns are dealt out to the devices
 * in contiguous blocks of about equal substep counts; one host thread per device drives that
 * device's column pipeline (H2D | substeps | D2H); the graph is analysed and uploaded once per
 * device and recognised afterwards by an exact byte comparison with the copy kept on the host; all
 * results land in the caller's single conc_out (columns that do not diffuse are copied through).
 *   devices / n_devices    CUDA ordinals, unique; NULL / 0 = every visible device
 * Host matrices may be pageable (what R hands over) or pinned: pageable memory is staged through
 * pinned slot rings next to each GPU by copier threads (csrc/hoststage.h), pinned memory is read
 * and written by the DMA engines directly. */
EUT_API int eut_diffchange_par_multi(const int32_t *adjmat, int64_t n_edges,
                                     const int32_t *nneighbors, int64_t n_nn,
                                     const double *conc_in, int64_t n_rows, int64_t n_cols,
                                     const int32_t *niter, const int32_t *indvar, int64_t n_indvar,
                                     int ncores, double *conc_out, const int *devices, int n_devices);

/* How eut_diffchange_par_multi deals the columns of a call out to n_devices devices (no CUDA call;
 * for tests and for callers that want to place data): owner[i] = index into the device list that
 * diffuses column indvar[i], -1 when niter[i] <= 0.  Blocks are contiguous in ascending column
 * number, hold about equal substep counts, and none is empty while columns remain. */
EUT_API int eut_deal_columns(const int32_t *niter, const int32_t *indvar, int64_t n_indvar,
                             int n_devices, int32_t *owner);

/* Drop the cached plans / staging buffers the one-shot tier keeps between calls. */
EUT_API void eut_oneshot_reset(void);

/* ---- session tier: plan = neighbour graph resident on one GPU --------------------------------
 *
 * Consumes the `mat.in` / `mat.out` slots of growthEnvironment as given
 * (R/growthEnvironment.R:50-51, built at :138-149): `from`/`to` are the two columns of mat.in,
 * `nn_id`/`nn_cnt` the two columns of mat.out.  The graph is never regenerated from geometry.
 * Per destination the inflow is summed in mat.in row order (stable sort by `to`).
 */
typedef struct eut_plan eut_plan;

enum {                       /* eut_plan_create flags */
    EUT_ORDER_AUTO      = 0, /* library picks the internal point ordering */
    EUT_ORDER_IDENTITY  = 1, /* keep the reference's (layer-major) order */
    EUT_ORDER_ROWBLOCK  = 2, /* whole raster rows reordered breadth-first over the row graph */
    EUT_ORDER_TILED     = 3, /* compact 3-D tiles derived from the graph + shared-memory kernel */
    EUT_ORDER_SEGMENTS  = 4, /* tiles of row segments: register reuse along the raster row (the
                                default when the graph qualifies; falls back to TILED, ROWBLOCK) */
    EUT_ORDER_MASK      = 0xf,
    EUT_PLAN_FORCE_GENERIC = 0x10, /* use the generic CSR kernel even when the fast path applies */
    EUT_PLAN_HOST_ONLY = 0x20      /* run the host analysis only (no CUDA calls): for inspecting the
                                      ordering / tile tables; such a plan cannot compute */
};

typedef struct eut_plan_info {
    int64_t n_points;        /* N */
    int64_t n_edges;         /* E */
    int64_t n_padded;        /* internal column stride (points) */
    int32_t max_in_degree;
    int32_t ell_width;       /* 0 when the generic CSR kernel is used */
    int32_t order;           /* EUT_ORDER_* actually used */
    int32_t regular_outflow; /* 1: exactly one mat.out row per point (the fast path) */
    int32_t device;
    int32_t n_tiles;
    int64_t device_bytes;    /* graph bytes resident in HBM */
    double  build_ms;        /* host analysis + upload time */
    int32_t tiled;           /* 1: the tiled shared-memory kernel is available */
    int32_t img_points;      /* points per shared-memory image (tiled) */
    int32_t max_tile_points;
    int32_t max_bulk;        /* most bulk (16-byte aligned run) copies of any tile */
    int32_t max_single;      /* most single 8-byte copies of any tile */
    int32_t n_copies;        /* entries in the copy table */
    double  halo_ratio;      /* halo points / own points, summed over tiles (tiled) */
    int32_t segmented;       /* 1: the row-segment kernel is available (EUT_ORDER_SEGMENTS) */
    int32_t seg_points;      /* R: points per full segment */
    int32_t max_tile_segs;   /* most segments of any tile */
    int32_t max_tile_generic;/* most generic (twelve explicit slots) points of any tile */
    int64_t n_segments;
    int64_t n_generic;       /* points that failed the canonical-order check */
    int64_t n_seg_covered;   /* points covered by segments (= N - n_generic) */
} eut_plan_info;

EUT_API int eut_plan_create(eut_plan **out,
                            const int32_t *from, const int32_t *to, int64_t n_edges,
                            const int32_t *nn_id, const int32_t *nn_cnt, int64_t n_nn,
                            int64_t n_points, int device, uint32_t flags);
EUT_API int eut_plan_destroy(eut_plan *plan);
EUT_API int eut_plan_get_info(const eut_plan *plan, eut_plan_info *info);
/* The in-neighbour index map the device uses, translated back to reference ids (1-based), for
 * bit-exact comparison with the CSR derived from mat.in: row_ptr[N+1], col[E].  Read back from the
 * tables the selected kernel indexes with: ELL / CSR arrays, or -- for a segment plan -- the tile
 * meta records, copy lists, segment and generic-point records, replayed in the kernel's summation
 * order (a EUT_PLAN_HOST_ONLY segment plan replays the planner's host copies of the same tables). */
EUT_API int eut_plan_get_csr(const eut_plan *plan, int64_t *row_ptr, int32_t *col);
/* Tile tables of a EUT_PLAN_HOST_ONLY plan (sizes from eut_plan_get_info), for inspection and for
 * the CPU tests of the planner:
 *   tile_meta[8*n_tiles]  per tile: first internal id, points, first bulk copy, #bulk, first single
 *                         copy, #single, bytes moved by the bulk copies, image points
 *   copies[2*n_copies]    bulk: (first internal id, image offset | n_points<<16); single: (id, offset)
 *   loc[12*n_padded]      byte offset inside the image of every neighbour slot, slot-major
 * Any pointer may be NULL. */
EUT_API int eut_plan_get_tiles(const eut_plan *plan, int32_t *tile_meta, int32_t *copies,
                               uint16_t *loc);
/* Segment tables of a EUT_PLAN_HOST_ONLY plan built with EUT_ORDER_SEGMENTS (segs.cu):
 *   seg_meta[12*n_tiles]  per tile: first internal id, points, first bulk copy, #bulk, first single
 *                         copy, #single, bulk bytes, image cells, first segment, #segments, first
 *                         generic point, #generic
 *   copies[2*n_copies]    as eut_plan_get_tiles
 *   segs[24*n_segments]   byte offsets inside the image (0 = the zero header): [0..6] first in-line
 *                         cell of the streams O, A, B, C, D, E, F; [7] result cell | length << 12;
 *                         [8,9] left / right neighbour of the segment in its own row; [10..17]
 *                         first / last cell of the pair streams A, B, E, F; rest 0
 *   gens[16*n_generic]    image cells of the 12 neighbour slots (0 = unused), own cell, result
 *                         cell, 0, 0
 * Any pointer may be NULL. */
EUT_API int eut_plan_get_segs(const eut_plan *plan, int32_t *seg_meta, int32_t *copies,
                              uint16_t *segs, uint16_t *gens);
/* Internal position (0-based) of every reference point id: perm[N]. */
EUT_API int eut_plan_get_perm(const eut_plan *plan, int32_t *perm);

/* Neighbour index maps from field points on the device: `build.DCM` and the emit step of
 * `initialize` (R/growthEnvironment.R:164-273, :138-149).  pts: n x 3 column-major (x, y, z), leading
 * dimension ld.  mat_in: caller-provided cap x 2 column-major int32 (from | to), cap >= number of
 * edges (12 n always suffices); rows [0, *n_edges) are filled, ordered by (to, from).  mat_out: n x 2
 * column-major (id | n_neighbors = out-degree + 1).  1-based ids like the R slots. */
EUT_API int eut_build_dcm(const double *pts, int64_t ld, int64_t n_points, double field_size, int device,
                          int32_t *mat_in, int64_t cap, int64_t *n_edges, int32_t *mat_out);

/* ---- session tier: field = N x C concentrations resident in HBM -------------------------------
 * Mirrors growthEnvironment@concentrations (R/growthEnvironment.R:44) / @exoenzymes.conc (:56). */
typedef struct eut_field eut_field;

EUT_API int eut_field_create(eut_field **out, eut_plan *plan, int64_t n_cols);
EUT_API int eut_field_destroy(eut_field *field);
/* host -> device; `host` is column-major with leading dimension ld (>= N).  cols = NULL copies
 * columns 1..n_cols of the field from host columns 0..n_cols-1; otherwise cols[i] (1-based field
 * column) receives host column i. */
EUT_API int eut_field_upload(eut_field *field, const double *host, int64_t ld,
                             const int32_t *cols, int64_t n_cols);
EUT_API int eut_field_download(eut_field *field, double *host, int64_t ld,
                               const int32_t *cols, int64_t n_cols);
/* `niter[i]` Jacobi substeps on column `indvar[i]` (the diffChangePar contract), asynchronous on
 * the field's stream. */
EUT_API int eut_field_diffuse(eut_field *field, const int32_t *niter, const int32_t *indvar,
                              int64_t n_indvar);
EUT_API int eut_field_sync(eut_field *field);
/* cudaStream_t the field's work is enqueued on (for CUDA-event timing by the caller), and a
 * setter to make the field use a caller-owned stream. */
EUT_API void *eut_field_stream(eut_field *field);
EUT_API int eut_field_set_stream(eut_field *field, void *cuda_stream);
/* Number of kernels this field has launched since creation (bench.py's gpu_launches). */
EUT_API int64_t eut_field_launch_count(const eut_field *field);
/* Kernel variant selection for experiments: see DESIGN.md ("kernel variants"). */
EUT_API int eut_field_set_option(eut_field *field, const char *name, int64_t value);
/* Diagnostics: with option "debug" bit 2048 set, every CTA of the segment kernel records how long it lived;
 * this copies the nanoseconds of the last launch, [2][4096] by (column slice 0 / 1, tile), from the current
 * device (tools/tile_times.py relates them to the tiles' properties).  Returns 0 on success. */
EUT_API int eut_debug_tile_ns(unsigned long long *out, int n);
/* Concentration recording (R/run_simulation.R:511-514, `round(COI.rec, digits = 7)`): like
 * eut_field_download, with every value rounded to `digits` (1..15) decimals on the device on its way out, by
 * base R's rule (the decimal neighbour closer to the represented value, ties to even).  The RDS file itself
 * (`saveRDS`, :517) is written on the host: eutropia_b200/recording.py. */
EUT_API int eut_field_record(eut_field *field, double *host, int64_t ld, const int32_t *cols, int64_t n_cols,
                             int digits);
/* Per-column reductions on the device (diffuse_compoundsPar's column selection and D = Inf rule,
 * R/growthEnvironment.R:372, :393-401): writes min, max, mean of every column (arrays of n_cols). */
EUT_API int eut_field_column_stats(eut_field *field, double *col_min, double *col_max,
                                   double *col_mean);
/* Exoenzyme catalysis on the device (R/run_simulation.R:343-364), point by point:
 *   vmax = kcat * E / 1e6;  S_t = km * W0( S_0/km * exp((S_0 - vmax * delta_time * 60 * 60) / km) );
 *   conc[, met_cols[i]] += stoich[i] * (S_0 - S_t)   for every i whose column is not constant,
 * E = column `exo_col` of `exo` (exoenzymes.conc, nM), S_0 = column met_cols[0] of `conc` (read
 * before any update).  Both fields must share a plan.  is_constant: n_cols(conc) flags or NULL. */
EUT_API int eut_field_exoenzyme_catalysis(eut_field *conc, eut_field *exo, int32_t exo_col, double kcat,
                                          double km, const int32_t *met_cols, const double *stoich,
                                          int32_t n_mets, const uint8_t *is_constant, double delta_time);
/* column *= factor (exoenzyme decay, E * exp(-lambda * deltaTime), R/run_simulation.R:367-371; the
 * factor is evaluated by the caller in binary64 like R does). */
EUT_API int eut_field_scale_column(eut_field *field, int32_t col, double factor);
/* Overwrite a column with one value (D = Inf -> column mean, R/growthEnvironment.R:393-401). */
EUT_API int eut_field_fill_column(eut_field *field, int32_t col, double value);

/* Run eut_field_halo_pack / _unpack on `stream` (a cudaStream_t; NULL = the field's stream), so that
 * the exchange of the boundary tiles' results overlaps the interior tiles of the same substep:
 *   eut_field_set_option(f, "phase", 1); eut_field_diffuse(one substep)   -- boundary tiles, flip
 *   [halo stream waits; pack -> send/recv -> unpack on the halo stream]
 *   eut_field_set_option(f, "phase", 2); eut_field_diffuse(same call)      -- interior tiles, no flip
 *   [field stream waits for the halo stream]; eut_field_set_option(f, "phase", 0) when done. */
EUT_API int eut_field_set_halo_stream(eut_field *field, void *stream);
/* Halo support for spatial slab sharding (one process per GPU, each with the plan of its slab: own
 * points + ghost copies of the remote in-neighbours; eutropia_b200/slab.py builds those).  The
 * reference has no counterpart (it runs on one host); the exchange itself is done by the caller
 * with NCCL send/recv on the device buffers named here, stream-ordered on eut_field_stream().
 *   eut_field_halo_set     point ids (1-based, this plan's numbering) whose values are sent to /
 *                          received from peers, each list in the order of the exchange buffer
 *   eut_field_halo_pack    dev_out[s * n_cols + j] = current value of column cols[j] at send_ids[s]
 *   eut_field_halo_unpack  current value of column cols[j] at recv_ids[r] = dev_in[r * n_cols + j]
 *                          (point-major, so the points of one peer are one contiguous slice)
 * dev_out / dev_in are DEVICE pointers (e.g. torch tensors). */
EUT_API int eut_field_halo_set(eut_field *field, const int32_t *send_ids, int64_t n_send,
                               const int32_t *recv_ids, int64_t n_recv);
EUT_API int eut_field_halo_pack(eut_field *field, const int32_t *cols, int64_t n_cols, void *dev_out);
EUT_API int eut_field_halo_unpack(eut_field *field, const int32_t *cols, int64_t n_cols, const void *dev_in);

/* ---- session tier: one field sharded by spatial slab over several GPUs, one process -----------
 * For fields larger than one GPU (BASELINE.json configs[4]).  The reference has no counterpart (one
 * host, one matrix); the contract is diffChangePar's: the sharded run is bit-identical to the
 * unsharded one.  The slabs are derived from the graph alone (contiguous ranges of the tile planner's
 * internal order); every slab holds its own points plus ghost copies of the remote in-neighbours and
 * its own plan over that local graph.  After every substep the owner of a boundary point stores the
 * new value straight into the neighbour's ghost cell through a peer-mapped pointer (NVLink P2P
 * within the process, one small kernel per ordered pair of neighbouring slabs) -- no pack / send /
 * recv / unpack.  From three slabs on, every slab keeps four rings of ghost points, the inner three diffused
 * like its own points, and the slabs exchange (and wait for each other) every fourth substep only
 * (EUT_SLAB_DEPTH overrides).  `devices` may name a device more than once (several slabs on one GPU).
 * (eutropia_b200/slab.py drives the same partition with one process per GPU and NCCL send / recv
 * through eut_field_halo_*.) */
typedef struct eut_slab eut_slab;
typedef struct eut_slab_info {
    int64_t n_points, n_cols;
    int32_t n_slabs, n_links;   /* links: ordered pairs of slabs that exchange points */
    int64_t max_local;          /* most points (own + ghost) of any slab */
    int64_t halo_points;        /* points that travel per exchange and column, summed over all links */
    int64_t max_link;           /* most points of any link */
    int64_t pushes;             /* halo kernels launched so far */
    int32_t segmented;          /* 1: every slab runs the segment kernel */
    int32_t depth;              /* ghost rings per slab = substeps between two halo exchanges */
    double  build_ms;
} eut_slab_info;
EUT_API int eut_slab_create(eut_slab **out, const int32_t *from, const int32_t *to, int64_t n_edges,
                            const int32_t *nn_id, const int32_t *nn_cnt, int64_t n_nn, int64_t n_points,
                            int64_t n_cols, const int *devices, int n_devices);
EUT_API int eut_slab_destroy(eut_slab *slab);
EUT_API int eut_slab_get_info(const eut_slab *slab, eut_slab_info *info);
/* Device, point counts and cudaStream_t of one slab (any pointer may be NULL). */
EUT_API int eut_slab_part(const eut_slab *slab, int index, int32_t *device, int64_t *n_local, int64_t *n_own,
                          void **stream);
/* host <-> slabs: the global N x C matrix, column-major, leading dimension ld. */
EUT_API int eut_slab_upload(eut_slab *slab, const double *host, int64_t ld);
EUT_API int eut_slab_download(eut_slab *slab, double *host, int64_t ld);
/* niter[i] substeps on column indvar[i] (diffChangePar contract), asynchronous on the slabs' streams. */
EUT_API int eut_slab_diffuse(eut_slab *slab, const int32_t *niter, const int32_t *indvar, int64_t n_indvar);
EUT_API int eut_slab_sync(eut_slab *slab);
EUT_API int64_t eut_slab_launch_count(const eut_slab *slab);

/* Cells on a slab-sharded field.  Every slab keeps, of every cell, the entries on its own points;
 * per-field work (claim rescale, scatter sums in cell order, clamp) runs on the owner exactly as on an
 * unsharded field, per-cell quantities (production shares, accessible amounts = colSums of the uptake
 * shares) are combined over the slabs a cell touches.  Same argument meaning as eut_cells_*; `pts` are the
 * global field points, field ids in eut_slab_cells_get_lists are global.  Destroy the cells before the slab they
 * were created on. */
typedef struct eut_slab_cells eut_slab_cells;
EUT_API int eut_slab_cells_create(eut_slab_cells **out, eut_slab *slab);
EUT_API int eut_slab_cells_destroy(eut_slab_cells *cells);
EUT_API int eut_slab_cells_build_lists(eut_slab_cells *cells, const double *pts, int64_t ld, int64_t n_pts,
                                       const double *cx, const double *cy, const double *csize, const double *cscv,
                                       int64_t n_cells, int dist_mode);
EUT_API int eut_slab_cells_claim_rescale(eut_slab_cells *cells);
EUT_API int64_t eut_slab_cells_num_entries(const eut_slab_cells *cells);
EUT_API int eut_slab_cells_get_lists(const eut_slab_cells *cells, int slab_index, int64_t *n_entries, int64_t *cell_ptr,
                                     int32_t *field_id, double *field_dist, double *acc_prop);
EUT_API int eut_slab_cells_gather(eut_slab_cells *cells, double field_vol, double *fmol_host);
EUT_API int eut_slab_cells_scatter(eut_slab_cells *cells, double field_vol, const uint8_t *is_constant,
                                   const int64_t *ex_ptr, const int32_t *ex_cpd, const double *ex_flx);

/* ---- session tier: cells = ragged cell -> field lists + gather / scatter ----------------------
 * Lists follow get_cellFieldEnv_ex's output (R/run_simulation.R:591-624): per cell the field ids
 * (1-based), the cell-field distances and the accessible proportions, in CSR form. */
typedef struct eut_cells eut_cells;

EUT_API int eut_cells_create(eut_cells **out, eut_plan *plan);
EUT_API int eut_cells_destroy(eut_cells *cells);
/* Install lists computed elsewhere (e.g. by R): cell_ptr[n_cells+1] 0-based offsets. */
EUT_API int eut_cells_set_lists(eut_cells *cells, int64_t n_cells, const int64_t *cell_ptr,
                                const int32_t *field_id, const double *field_dist,
                                const double *acc_prop);
/* Build the lists on the device from geometry: replaces the per-cell box prefilter +
 * get_cellFieldEnv_ex (R/run_simulation.R:217-232, :591-624).  pts is N x 3 column-major.
 * dist_mode 0: planar xy distance (sf/GEOS), 1: xyz. */
EUT_API int eut_cells_build_lists(eut_cells *cells, const double *pts, int64_t n_pts,
                                  const double *cx, const double *cy, const double *csize,
                                  const double *cscv, int64_t n_cells, int dist_mode);
/* Oversubscribed-field rescale of acc_prop (R/run_simulation.R:236-241). */
EUT_API int eut_cells_claim_rescale(eut_cells *cells);
EUT_API int64_t eut_cells_num_entries(const eut_cells *cells);
EUT_API int eut_cells_get_lists(const eut_cells *cells, int64_t *cell_ptr, int32_t *field_id,
                                double *field_dist, double *acc_prop);
/* Gather (R/run_simulation.R:249, R/scavenge_compounds.R:17-22): fmol[n_cells x n_cols]
 * column-major = accessible fmol per cell and compound. */
EUT_API int eut_cells_gather(eut_cells *cells, eut_field *field, double field_vol, double *fmol_host);
/* Chemotaxis (R/run_simulation.R:260-291): for compound column `col`, per cell over its field list the
 * sums the reference's weighted means are made of.  pts: N x 3 column-major field points (x | y | z);
 * cell_x, cell_y: cell positions; cell_r = size / 2 + scavengeDist.  out: M x 7 column-major:
 * sum c | sum c^2 | sum dx c | sum dy c | sum dx | sum dy | n   with dx = (x - cell_x) / cell_r.
 * centr_x = sum dx c / sum c (sum dx / n when sum c = 0), weighted.mean(conc) = sum c^2 / sum c. */
EUT_API int eut_cells_chemotaxis_moments(eut_cells *cells, eut_field *field, int32_t col, const double *pts,
                                         int64_t n_pts, const double *cell_x, const double *cell_y,
                                         const double *cell_r, double *out);
/* Scatter + clamp (R/run_simulation.R:689-740, :309-319).  Per cell i the exchanges
 * ex_ptr[i]..ex_ptr[i+1]: ex_cpd (1-based field column), ex_flx (fmol; <0 uptake, >0 production).
 * is_constant[n_cols] (may be NULL) marks buffered compounds, which are left untouched. */
EUT_API int eut_cells_scatter(eut_cells *cells, eut_field *field, double field_vol,
                              const uint8_t *is_constant, const int64_t *ex_ptr,
                              const int32_t *ex_cpd, const double *ex_flx);

#ifdef __cplusplus
}
#endif
#endif /* EUTROPIA_B200_H */
