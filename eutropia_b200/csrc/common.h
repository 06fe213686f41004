// Internal declarations shared by the translation units of libeutropia_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/eutropia_b200.h"

namespace eut {

void set_error(const char *fmt, ...);
void clear_error();

#define EUT_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            eut::set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,      \
                           __LINE__);                                                       \
            return EUT_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define EUT_TRY(call)                \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != EUT_OK) return rc__; \
    } while (0)

// 1/13: the share of a field's content that moves to each of its (at most 12) neighbours per
// substep, src/diffChange.cpp:8.
constexpr double kBDiv = 1.0 / 13;

constexpr int kEllMax = 12;       // widest in-neighbour list the ELL fast path handles
constexpr int kPadPoints = 64;    // column stride is a multiple of this many points (512 B)
constexpr int kTileMaxCols = 64;  // compound columns one CTA of the tiled kernel loops over

int ensure_device(int device);
double now_ms();

}  // namespace eut

// ---------------------------------------------------------------------------------------------
// tiled path (tiles.cu): host-side result of the tile planner
// ---------------------------------------------------------------------------------------------
namespace eut {
struct TileParams {
    int tile_cols = 64;          // columns of the pseudo x coordinate per tile
    int band_levels = 16;        // tiles are ordered by coarse level band, then by u column
    int max_tile_points = 1024;  // = threads per CTA x points per thread of the kernel
    int max_img_points = 4096;   // shared-memory image limit (points) per buffer
    int max_bulk = 64;           // bulk copies per tile the kernel can hold (2 per lane of warp 0)
    int max_single = 512;        // single 8-byte copies per tile (2 per thread)
};
struct TilePlanHost {
    std::vector<int32_t> tile_start;  // [n_tiles+1] internal ids
    // per tile 8 ints: first id, points, first bulk copy, #bulk, first single copy, #single,
    // bytes moved by the bulk copies, image points
    std::vector<int32_t> meta;
    // bulk:   x = first internal id (even), y = image offset (even) | n_points (even) << 16
    // single: x = internal id,              y = image offset
    std::vector<int2> copies;
    std::vector<uint16_t> loc;        // [12][Np] byte offset of every in-neighbour inside the image
    int img_points = 0, max_tile_points = 0, max_bulk = 0, max_single = 0;
    int64_t n_copies = 0;
    double halo_ratio = 0;
};
}  // namespace eut

// ---------------------------------------------------------------------------------------------
// row analysis (rows.cu) and the segment planner (segs.cu)
// ---------------------------------------------------------------------------------------------
namespace eut {
struct RowAnalysis {
    int32_t n_rows = 0, xmin = 0;
    std::vector<int32_t> row_of;      // [N]   row of every reference id
    std::vector<int32_t> row_start;   // [R+1] first reference id of every row
    std::vector<int64_t> adj_ptr;     // [R+1] adjacent rows of every row, ascending
    std::vector<int32_t> adj_row, adj_delta;
    std::vector<uint8_t> adj_pair;    // some point of the row has two neighbours in the other row
    std::vector<int32_t> level, xs;   // [R]   BFS level, column coordinate of the row's first point
};
struct SegParams {
    int R = 7;                   // points per segment (odd: conflict-free 8-byte strides)
    int tile_cols = 56;          // columns of the pseudo x coordinate per tile (8 segments)
    int band_levels = 16;
    int max_tile_segs = 224;     // consumer threads per CTA x segments per thread of the kernel
    int max_tile_points = 4000;  // 12-bit offsets into the tile's result buffer
    int max_img_points = 8190;   // shared-memory image limit (cells) per buffer: 16-bit byte offsets
    int max_bulk = 128;          // bulk copies per tile the kernel can hold (4 per lane of warp 0)
    int max_single = 512;        // single 8-byte copies per tile
    int pad_rows = 1;            // gap cells after ragged row pieces: 0 none, 1 per-tile residue step (default),
                                 // 3 global rule position - u == 8 * colour (mod 16): conflict free, ~6 % gaps
};
struct SegPlanHost {
    std::vector<int32_t> tile_start;  // [n_tiles+1] internal ids
    // per tile 12 ints: first id, points, first bulk copy, #bulk, first single copy, #single, bytes
    // moved by the bulk copies, image cells, first segment, #segments, first generic point, #generic
    std::vector<int32_t> meta;
    std::vector<int2> copies;         // as TilePlanHost::copies
    // per segment 24 x uint16 (byte offsets inside the image; 0 = zero header):
    //   [0..6]  first in-line cell of the streams O, A, B, C, D, E, F (pair streams: cell 0's slot)
    //   [7]     result-buffer cell | length << 12
    //   [8,9]   left / right neighbour of the segment in its own row
    //   [10..17] first / last cell of the pair streams A, B, E, F
    std::vector<uint16_t> segs;
    // per generic point 16 x uint16: image cells of the 12 neighbour slots (0 = unused), own image
    // cell, result-buffer cell, 0, 0
    std::vector<uint16_t> gens;
    int R = 0, zero_cells = 0, img_points = 0, max_tile_points = 0, max_tile_segs = 0, max_bulk = 0,
        max_single = 0, max_gen = 0;
    int64_t n_copies = 0, n_segs = 0, n_gens = 0, seg_points = 0;
    double halo_ratio = 0;
};
}  // namespace eut

// ---------------------------------------------------------------------------------------------
// plan: neighbour graph resident on one GPU
// ---------------------------------------------------------------------------------------------
struct eut_plan {
    eut::TilePlanHost tiles;
    eut::SegPlanHost segs;
    bool tiled = false;
    bool segmented = false;      // segs (not tiles) describes the tiles; kernel: stencil_seg.cu
    int32_t *d_seg_meta = nullptr;
    uint4 *d_segs = nullptr;     // three per segment
    uint4 *d_gens = nullptr;     // two per generic point
    bool host_only = false;
    int32_t *d_tile_meta = nullptr;
    int2 *d_copies = nullptr;
    uint16_t *d_loc16 = nullptr;
    int device = 0;
    int64_t N = 0, E = 0, Np = 0;
    int max_in = 0, ell_w = 0, order = 0, regular_out = 0, n_tiles = 0;
    int64_t device_bytes = 0;
    double build_ms = 0;
    uint64_t graph_hash = 0;

    // host, reference ids, 0-based: in-neighbours per destination in mat.in row order
    std::vector<int64_t> row_ptr;
    std::vector<int32_t> col;
    std::vector<int32_t> perm;   // reference id (0-based) -> internal position
    std::vector<int32_t> iperm;  // internal position -> reference id (0-based), -1 for padding
    // points nobody computes (set before plan_build by slab.cu: ghost copies of a neighbour slab's points,
    // overwritten after every substep): the segment planner gives them a cell but no work
    std::vector<uint8_t> dead;

    // device, internal order
    int32_t *d_nbr = nullptr;    // [ell_w][Np] internal ids of in-neighbours, mat.in order
    uint8_t *d_deg = nullptr;    // [Np] in-degree
    double *d_w = nullptr;       // [Np] (n_neighbors - 1) * (1/13)   (regular outflow only)
    int32_t *d_perm = nullptr;   // [N]
    int32_t *d_iperm = nullptr;  // [Np]
    // generic path (any in-degree, any number of mat.out rows per point)
    int64_t *d_row_ptr = nullptr;  // [N+1] internal order
    int32_t *d_col = nullptr;      // [E] internal ids
    int64_t *d_out_ptr = nullptr;  // [N+1]
    double *d_out_w = nullptr;     // [n_nn]  (nn_j - 1) * (1/13), in mat.out row order per point
};

// ---------------------------------------------------------------------------------------------
// field: N x C concentrations resident in HBM (two buffers, Jacobi ping-pong per column)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxLanes = 4;
namespace eut { struct HostStager; }

struct eut_field {
    eut_plan *plan = nullptr;
    int64_t C = 0;
    double *d_buf = nullptr;           // [2][C][Np]
    std::vector<uint8_t> cur;          // which buffer holds column c now
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;
    // staging for host <-> device column transfers (reference order)
    double *d_stage = nullptr;
    int64_t stage_cols = 0;
    // active-column tables of the current diffuse call
    int32_t *d_act = nullptr;          // [3][cap]: column, starting buffer, final buffer
    int64_t act_cap = 0;
    int64_t act_base = 0;              // first table entry the next diffuse call uses (one-shot pipeline: per chunk)
    // extra compute streams (created on demand): wide diffuse calls run their columns in up to
    // kMaxLanes parts side by side; the one-shot pipeline alternates its chunks between stream and lane[0]
    cudaStream_t lane[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    bool in_lanes = false;             // a lane launch is being issued (stencil_seg.cu slices the columns coarser)
    bool no_split = false;             // set by the one-shot pipeline, which uses lane[0] itself
    bool lane_shape = false;           // the caller runs column groups side by side itself (slab.cu): slice the columns as inside lanes
    // the tables the substep launch reads (default: d_act; the column-streaming pipeline points them
    // at a ring of per-launch tables)
    const int32_t *act_col = nullptr, *act_par = nullptr;
    int32_t *d_ring = nullptr, *h_ring = nullptr;   // [kRing][2][cap] device / pinned host
    int32_t *d_stream_cols = nullptr;                // [3][C]: column list, zeros, final buffer of each column
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int32_t> h_act;        // host mirror (kept alive for async copies)
    int32_t *d_cur = nullptr;          // [C] device mirror of cur
    std::vector<int32_t> h_cur32;
    int32_t *d_cols = nullptr;         // [2][C] scratch: column lists of transfers in flight
    std::vector<int32_t> h_cols;
    double *d_stage2 = nullptr;        // second staging buffer (download direction)
    cudaEvent_t ev_stage_free[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage_ready[2] = {nullptr, nullptr};
    cudaStream_t down_stream = nullptr;
    // pageable caller memory: pinned slot rings + copier threads (hoststage.h), created on first use
    eut::HostStager *stager = nullptr;
    unsigned copy_threads = 0;         // threads per copier pool (0: from the host's thread budget)
    cudaEvent_t ev_dstage_free[2] = {nullptr, nullptr};
    cudaEvent_t ev_dstage_ready[2] = {nullptr, nullptr};
    int64_t up_chunks = 0, down_chunks = 0;
    double *d_stats = nullptr;         // [3][C]
    uint64_t version = 0;              // bumped by everything that changes field values (cells.cu: is a gather still current?)
    std::atomic<int64_t> launches{0};  // also bumped by the staged downloader thread (hoststage.cu)
    // substep loops seen before, captured as CUDA graphs (key: substep counts + options)
    std::map<uint64_t, cudaGraphExec_t> graphs;
    std::set<uint64_t> graph_seen;
    // options
    int opt_cpt = 0;      // compounds per thread (0 = auto)
    int opt_block = 256;
    int opt_variant = 0;  // 0 auto, 1 ell, 2 generic
    int opt_group_cols = 0;  // 0: substep-major over all active columns; >0: column groups
    int opt_nbuf = 3;        // staging depth of the tiled kernel (3 or 4 columns)
    int opt_debug = 0;
    int opt_shape = 0;       // thread shape of the lean tiled kernel (stencil_tile.cu)
    // halo lists of a slab field (halo.cu)
    int32_t *d_halo_send = nullptr, *d_halo_recv = nullptr, *d_halo_cols = nullptr;
    std::vector<int32_t> h_halo_cols;
    int64_t n_halo_send = 0, n_halo_recv = 0;
    cudaStream_t halo_stream = nullptr;   // pack / unpack run here when set (overlap with interior tiles)
    // slab overlap: tiles that hold points a neighbour reads (boundary) and the rest (interior)
    int32_t *d_tiles_b = nullptr, *d_tiles_i = nullptr;
    int n_tiles_b = 0, n_tiles_i = 0;
    double round_p10 = 0.0;  // > 0: downloads round to log10(round_p10) decimals on the way out (eut_field_record)
    int opt_phase = 0;       // 0: all tiles; 1: boundary tiles, then flip; 2: interior tiles of the substep just flipped
};

namespace eut {
int plan_build(eut_plan *p, const int32_t *from, const int32_t *to, int64_t n_edges,
               const int32_t *nn_id, const int32_t *nn_cnt, int64_t n_nn, int64_t n_points,
               int device, uint32_t flags);
void plan_free(eut_plan *p);
uint64_t hash_bytes(const void *data, size_t n, uint64_t seed);
int build_tiles(eut_plan *p, const TileParams &tp);
void analyse_rows(const eut_plan *p, RowAnalysis &ra);
int build_segs(eut_plan *p, const SegParams &sp);

// field.cu
int field_create(eut_field **out, eut_plan *plan, int64_t n_cols);
int field_destroy(eut_field *f);
int field_upload(eut_field *f, const double *host, int64_t ld, const int32_t *cols, int64_t n_cols);
int field_download_async(eut_field *f, double *host, int64_t ld, const int32_t *cols, int64_t n_cols);
int field_sync_down(eut_field *f);
int field_diffuse(eut_field *f, const int32_t *niter, const int32_t *indvar, int64_t n_indvar);
// one-shot tier, column streaming: upload, diffuse and download `n` columns (1-based `cols`) with the
// substeps of a column starting as soon as it has arrived; `verify` is called before the first byte is
// written to h_out (false -> returns -1000 after draining the streams)
int field_stream_columns(eut_field *f, const double *h_in, double *h_out, int64_t ld, const int32_t *cols,
                         const int32_t *niter, int64_t n, const std::function<bool()> &verify);

// stencil.cu
int launch_substep(eut_field *f, int n_act, int substep);
int launch_tile_u(eut_field *f, int n_act, int substep);   // stencil_tile.cu
int launch_seg(eut_field *f, int n_act, int substep);      // stencil_seg.cu
int launch_permute_in(eut_field *f, const double *d_stage, int64_t ld, const int32_t *d_cols,
                      const int32_t *d_cur, int64_t n_cols, cudaStream_t s);
int launch_permute_out(eut_field *f, double *d_stage, int64_t ld, const int32_t *d_cols,
                       const int32_t *d_cur, int64_t n_cols, cudaStream_t s);
int launch_column_stats(eut_field *f, const int32_t *d_cur);
int launch_set_cur(eut_field *f, const int32_t *d_col, const int32_t *d_value, int n);
const char *last_error_cstr();
int launch_fill_column(eut_field *f, int32_t col0, double value);
// round.cu
int field_exoenzyme_catalysis(eut_field *cf, eut_field *ef, int32_t exo_col, double kcat, double km,
                              const int32_t *met_cols, const double *stoich, int32_t n_mets,
                              const uint8_t *is_constant, double delta_time);
int field_scale_column(eut_field *f, int32_t col0, double factor);
int max_exo_mets();
// cells.cu: hooks for cells on a slab-sharded field (slab.cu)
int64_t cells_count(const eut_cells *c);
int cells_get_dmax(eut_cells *c, double *host);
int cells_partial_fsum(eut_cells *c, const double *dmax_host, double *fsum_host);
int cells_set_fsum(eut_cells *c, const double *fsum_host);
double *cells_fmol(eut_cells *c);
void cells_mark_gathered(eut_cells *c, const eut_field *f, double field_vol);
bool cells_gather_is_current(const eut_cells *c, const eut_field *f, double field_vol);
// meshgen.cu
int mesh_build_dcm(const double *pts, int64_t ld, int64_t n, double fs, int device, int32_t *mat_in, int64_t cap,
                   int64_t *n_edges_out, int32_t *mat_out);
}  // namespace eut
