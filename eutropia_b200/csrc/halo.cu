// Halo pack / unpack for spatial slab sharding (see include/eutropia_b200.h, eutropia_b200/slab.py).
// The exchange buffers are plain [point][column] arrays in the order of the id lists, so that the
// points of one peer are a contiguous slice and one NCCL send / recv per peer moves it.
#include <algorithm>

#include "common.h"

namespace eut {

__global__ void __launch_bounds__(256)
k_halo_pack(const double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
            const int32_t *__restrict__ cur, const int32_t *__restrict__ cols,
            const int32_t *__restrict__ pos, const int n, double *__restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int c = cols[blockIdx.y];
    out[(int64_t)s * gridDim.y + blockIdx.y] = buf[(int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + pos[s]];
}

__global__ void __launch_bounds__(256)
k_halo_unpack(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
              const int32_t *__restrict__ cur, const int32_t *__restrict__ cols,
              const int32_t *__restrict__ pos, const int n, const double *__restrict__ in)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int c = cols[blockIdx.y];
    buf[(int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + pos[r]] = in[(int64_t)r * gridDim.y + blockIdx.y];
}

static int to_positions(const eut_plan *p, const int32_t *ids, int64_t n, std::vector<int32_t> &pos)
{
    pos.resize(n);
    for (int64_t k = 0; k < n; k++) {
        if (ids[k] < 1 || ids[k] > p->N) {
            set_error("index out of bounds: halo id %d, slab has %lld points", ids[k], (long long)p->N);
            return EUT_ERR_INDEX;
        }
        pos[k] = p->perm[ids[k] - 1];
    }
    return EUT_OK;
}

static int upload_cols(eut_field *f, const int32_t *cols, int64_t n_cols, int32_t **d_cols)
{
    if (n_cols < 0 || n_cols > f->C || (n_cols > 0 && !cols)) { set_error("halo: bad column list"); return EUT_ERR_ARG; }
    // the halo calls have their own column list (h_halo_cols / d_halo_cols); field.cu's d_cols slots belong to the transfers
    int32_t *hc = f->h_halo_cols.data();
    for (int64_t j = 0; j < n_cols; j++) {
        if (cols[j] < 1 || cols[j] > f->C) { set_error("halo: column %d out of bounds", cols[j]); return EUT_ERR_INDEX; }
        hc[j] = cols[j] - 1;
    }
    EUT_CUDA(cudaMemcpyAsync(f->d_halo_cols, hc, n_cols * sizeof(int32_t), cudaMemcpyHostToDevice, f->halo_stream ? f->halo_stream : f->stream));
    *d_cols = f->d_halo_cols;
    return EUT_OK;
}

}  // namespace eut

using namespace eut;

extern "C" {

int eut_field_halo_set(eut_field *f, const int32_t *send_ids, int64_t n_send, const int32_t *recv_ids,
                       int64_t n_recv)
{
    clear_error();
    if (!f || n_send < 0 || n_recv < 0 || (n_send > 0 && !send_ids) || (n_recv > 0 && !recv_ids)) {
        set_error("eut_field_halo_set: bad argument");
        return EUT_ERR_ARG;
    }
    EUT_TRY(ensure_device(f->plan->device));
    std::vector<int32_t> ps, pr;
    EUT_TRY(to_positions(f->plan, send_ids, n_send, ps));
    EUT_TRY(to_positions(f->plan, recv_ids, n_recv, pr));
    EUT_CUDA(cudaStreamSynchronize(f->stream));
    cudaFree(f->d_halo_send); cudaFree(f->d_halo_recv); cudaFree(f->d_halo_cols);
    f->d_halo_send = f->d_halo_recv = f->d_halo_cols = nullptr;
    EUT_CUDA(cudaMalloc((void **)&f->d_halo_send, std::max<size_t>(n_send, 1) * sizeof(int32_t)));
    EUT_CUDA(cudaMalloc((void **)&f->d_halo_recv, std::max<size_t>(n_recv, 1) * sizeof(int32_t)));
    EUT_CUDA(cudaMalloc((void **)&f->d_halo_cols, std::max<size_t>(f->C, 1) * sizeof(int32_t)));
    if (n_send) EUT_CUDA(cudaMemcpy(f->d_halo_send, ps.data(), n_send * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (n_recv) EUT_CUDA(cudaMemcpy(f->d_halo_recv, pr.data(), n_recv * sizeof(int32_t), cudaMemcpyHostToDevice));
    f->h_halo_cols.assign(std::max<size_t>(f->C, 1), 0);
    f->n_halo_send = n_send;
    f->n_halo_recv = n_recv;
    // boundary / interior tile lists of the segment plan: a substep can then run the boundary tiles
    // first, and the exchange of their results overlaps the interior tiles (eut_field_set_option
    // "phase", eutropia_b200/slab.py)
    cudaFree(f->d_tiles_b); cudaFree(f->d_tiles_i);
    f->d_tiles_b = f->d_tiles_i = nullptr;
    f->n_tiles_b = f->n_tiles_i = 0;
    if (f->plan->segmented) {
        const auto &meta = f->plan->segs.meta;
        const int nt = f->plan->n_tiles;
        std::vector<int32_t> t0(nt);
        for (int t = 0; t < nt; t++) t0[t] = meta[(size_t)t * 12];
        std::vector<uint8_t> is_b(nt, 0);
        // tiles with points a neighbour reads, and tiles with ghost points: a substep copies a ghost's
        // old value forward, which must happen before -- not while -- the unpack overwrites it
        for (const std::vector<int32_t> *lst : {&ps, &pr})
            for (int32_t q : *lst) {
                const int t = (int)(std::upper_bound(t0.begin(), t0.end(), q) - t0.begin()) - 1;
                if (t >= 0 && q < meta[(size_t)t * 12] + meta[(size_t)t * 12 + 1]) is_b[t] = 1;
            }
        std::vector<int32_t> tb, ti;
        for (int t = 0; t < nt; t++) (is_b[t] ? tb : ti).push_back(t);
        EUT_CUDA(cudaMalloc((void **)&f->d_tiles_b, std::max<size_t>(tb.size(), 1) * sizeof(int32_t)));
        EUT_CUDA(cudaMalloc((void **)&f->d_tiles_i, std::max<size_t>(ti.size(), 1) * sizeof(int32_t)));
        if (!tb.empty()) EUT_CUDA(cudaMemcpy(f->d_tiles_b, tb.data(), tb.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        if (!ti.empty()) EUT_CUDA(cudaMemcpy(f->d_tiles_i, ti.data(), ti.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        f->n_tiles_b = (int)tb.size(); f->n_tiles_i = (int)ti.size();
    }
    return EUT_OK;
}

int eut_field_set_halo_stream(eut_field *f, void *stream)
{
    clear_error();
    if (!f) return EUT_ERR_ARG;
    f->halo_stream = (cudaStream_t)stream;
    return EUT_OK;
}

int eut_field_halo_pack(eut_field *f, const int32_t *cols, int64_t n_cols, void *dev_out)
{
    clear_error();
    if (!f || !f->d_halo_cols) { set_error("eut_field_halo_pack: call eut_field_halo_set first"); return EUT_ERR_ARG; }
    if (n_cols == 0 || f->n_halo_send == 0) return EUT_OK;
    if (!dev_out) { set_error("eut_field_halo_pack: NULL buffer"); return EUT_ERR_ARG; }
    EUT_TRY(ensure_device(f->plan->device));
    int32_t *d_cols = nullptr;
    EUT_TRY(upload_cols(f, cols, n_cols, &d_cols));
    dim3 grid((unsigned)((f->n_halo_send + 255) / 256), (unsigned)n_cols, 1);
    k_halo_pack<<<grid, 256, 0, f->halo_stream ? f->halo_stream : f->stream>>>(f->d_buf, f->plan->Np, f->C * f->plan->Np, f->d_cur, d_cols,
                                             f->d_halo_send, (int)f->n_halo_send, (double *)dev_out);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

int eut_field_halo_unpack(eut_field *f, const int32_t *cols, int64_t n_cols, const void *dev_in)
{
    clear_error();
    if (!f || !f->d_halo_cols) { set_error("eut_field_halo_unpack: call eut_field_halo_set first"); return EUT_ERR_ARG; }
    if (n_cols == 0 || f->n_halo_recv == 0) return EUT_OK;
    if (!dev_in) { set_error("eut_field_halo_unpack: NULL buffer"); return EUT_ERR_ARG; }
    EUT_TRY(ensure_device(f->plan->device));
    int32_t *d_cols = nullptr;
    EUT_TRY(upload_cols(f, cols, n_cols, &d_cols));
    f->version++;
    dim3 grid((unsigned)((f->n_halo_recv + 255) / 256), (unsigned)n_cols, 1);
    k_halo_unpack<<<grid, 256, 0, f->halo_stream ? f->halo_stream : f->stream>>>(f->d_buf, f->plan->Np, f->C * f->plan->Np, f->d_cur, d_cols,
                                               f->d_halo_recv, (int)f->n_halo_recv, (const double *)dev_in);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

}  // extern "C"
