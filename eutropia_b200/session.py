"""Session tier: graph, field and cell lists resident on one B200 (thin wrappers over the C ABI).

`Plan` <- `mat.in` / `mat.out` (R/growthEnvironment.R:50-51), `Field` <- `concentrations` /
`exoenzymes.conc` (:44, :56), `Cells` <- the per-cell lists of `get_cellFieldEnv_ex`
(R/run_simulation.R:591-624).  All compute happens in libeutropia_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr

__all__ = ["Plan", "Field", "Cells", "SlabField"]


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(a):
    a = np.asarray(a)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.require(a, dtype=np.float64, requirements=["F", "A"])


class Plan:
    """The neighbour graph resident on a GPU, consumed exactly as given."""

    def __init__(self, mat_in, mat_out, n_points=None, device=0, order=_lib.ORDER_AUTO,
                 force_generic=False, host_only=False):
        mat_in = np.asarray(mat_in).reshape(-1, 2)
        mat_out = np.asarray(mat_out).reshape(-1, 2)
        self._from, self._to = _i32(mat_in[:, 0]), _i32(mat_in[:, 1])
        self._id, self._cnt = _i32(mat_out[:, 0]), _i32(mat_out[:, 1])
        if n_points is None:
            n_points = int(mat_out.shape[0])
        flags = int(order) | (_lib.PLAN_FORCE_GENERIC if force_generic else 0)
        flags |= _lib.PLAN_HOST_ONLY if host_only else 0
        h = C.c_void_p()
        check(_lib.lib().eut_plan_create(C.byref(h), ptr(self._from), ptr(self._to),
                                         self._from.shape[0], ptr(self._id), ptr(self._cnt),
                                         self._id.shape[0], int(n_points), int(device), flags))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and getattr(_lib, "lib", None):     # module globals are gone at interpreter exit
            _lib.lib().eut_plan_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def info(self) -> _lib.PlanInfo:
        inf = _lib.PlanInfo()
        check(_lib.lib().eut_plan_get_info(self._h, C.byref(inf)))
        return inf

    @property
    def n_points(self) -> int:
        return int(self.info.n_points)

    def csr(self):
        """(row_ptr int64[N+1], col int32[E]) as the device tables hold them, in reference ids."""
        inf = self.info
        rp = np.zeros(inf.n_points + 1, dtype=np.int64)
        col = np.zeros(max(inf.n_edges, 1), dtype=np.int32)
        check(_lib.lib().eut_plan_get_csr(self._h, ptr(rp), ptr(col)))
        return rp, col[:inf.n_edges]

    def tiles(self):
        """Host tile tables of a `host_only` plan: (meta[n_tiles,8], copies[n,2], loc[12,Np]); see
        eut_plan_get_tiles in include/eutropia_b200.h for the layout."""
        inf = self.info
        meta = np.zeros((max(inf.n_tiles, 1), 8), dtype=np.int32)
        copies = np.zeros((max(inf.n_copies, 1), 2), dtype=np.int32)
        loc = np.zeros((12, inf.n_padded), dtype=np.uint16)
        check(_lib.lib().eut_plan_get_tiles(self._h, ptr(meta), ptr(copies), ptr(loc)))
        return meta[:inf.n_tiles], copies[:inf.n_copies], loc

    def segs(self):
        """Host segment tables of a `host_only` plan built with ORDER_SEGMENTS: (meta[n_tiles,12],
        copies[n,2], segs[n_segments,24], gens[n_generic,16]); see eut_plan_get_segs."""
        inf = self.info
        meta = np.zeros((max(inf.n_tiles, 1), 12), dtype=np.int32)
        copies = np.zeros((max(inf.n_copies, 1), 2), dtype=np.int32)
        segs = np.zeros((max(inf.n_segments, 1), 24), dtype=np.uint16)
        gens = np.zeros((max(inf.n_generic, 1), 16), dtype=np.uint16)
        check(_lib.lib().eut_plan_get_segs(self._h, ptr(meta), ptr(copies), ptr(segs), ptr(gens)))
        return meta[:inf.n_tiles], copies[:inf.n_copies], segs[:inf.n_segments], gens[:inf.n_generic]

    def perm(self):
        p = np.zeros(max(self.n_points, 1), dtype=np.int32)
        check(_lib.lib().eut_plan_get_perm(self._h, ptr(p)))
        return p[:self.n_points]


class Field:
    """N x C concentrations resident in HBM across rounds."""

    def __init__(self, plan: Plan, n_cols: int):
        self.plan = plan
        self.n_cols = int(n_cols)
        h = C.c_void_p()
        check(_lib.lib().eut_field_create(C.byref(h), plan._h, self.n_cols))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(_lib, "lib", None):
            _lib.lib().eut_field_destroy(self._h)
            self._h = None

    __del__ = close

    def upload(self, conc, cols=None):
        """`conc`: N x len(cols) (or N x C) column-major host matrix; `cols` 1-based."""
        a = _colmajor(conc)
        c = None if cols is None else _i32(cols)
        n = a.shape[1] if c is None else c.shape[0]
        ld = a.strides[1] // 8 if a.shape[1] > 1 else max(a.shape[0], 1)
        check(_lib.lib().eut_field_upload(self._h, ptr(a), ld, ptr(c), n))

    def upload_ptr(self, host_ptr: int, ld: int, cols=None, n_cols=None):
        c = None if cols is None else _i32(cols)
        n = int(n_cols) if c is None else c.shape[0]
        check(_lib.lib().eut_field_upload(self._h, C.c_void_p(host_ptr), int(ld), ptr(c), n))

    def download(self, cols=None, out=None):
        c = None if cols is None else _i32(cols)
        n = self.n_cols if c is None else c.shape[0]
        npts = self.plan.n_points
        if out is None:
            out = np.empty((npts, n), dtype=np.float64, order="F")
        check(_lib.lib().eut_field_download(self._h, ptr(out), max(npts, 1), ptr(c), n))
        return out

    def record(self, cols=None, digits: int = 7):
        """Columns rounded to `digits` decimals on the device on their way out (`round(COI.rec, digits = 7)`,
        R/run_simulation.R:514); base R's rounding rule, see eut_field_record."""
        c = None if cols is None else _i32(cols)
        n = self.n_cols if c is None else c.shape[0]
        npts = self.plan.n_points
        out = np.empty((npts, n), dtype=np.float64, order="F")
        check(_lib.lib().eut_field_record(self._h, ptr(out), max(npts, 1), ptr(c), n, int(digits)))
        return out

    def download_ptr(self, host_ptr: int, ld: int, cols=None, n_cols=None):
        c = None if cols is None else _i32(cols)
        n = int(n_cols) if c is None else c.shape[0]
        check(_lib.lib().eut_field_download(self._h, C.c_void_p(host_ptr), int(ld), ptr(c), n))

    def diffuse(self, niter, indvar):
        """`niter[i]` substeps on column `indvar[i]` (1-based): the diffChangePar contract."""
        ni = _i32(np.trunc(np.asarray(niter, dtype=np.float64)))
        iv = _i32(np.trunc(np.asarray(indvar, dtype=np.float64)))
        if ni.shape[0] < iv.shape[0]:
            raise _lib.EutropiaIndexError(_lib.ERR_INDEX, "niter is shorter than indvar")
        check(_lib.lib().eut_field_diffuse(self._h, ptr(ni), ptr(iv), iv.shape[0]))

    def sync(self):
        check(_lib.lib().eut_field_sync(self._h))

    @property
    def stream(self) -> int:
        return int(_lib.lib().eut_field_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(_lib.lib().eut_field_launch_count(self._h))

    def set_option(self, name: str, value: int):
        check(_lib.lib().eut_field_set_option(self._h, name.encode(), int(value)))

    def column_stats(self):
        mn = np.zeros(max(self.n_cols, 1)); mx = np.zeros_like(mn); me = np.zeros_like(mn)
        check(_lib.lib().eut_field_column_stats(self._h, ptr(mn), ptr(mx), ptr(me)))
        return mn[:self.n_cols], mx[:self.n_cols], me[:self.n_cols]

    def fill_column(self, col: int, value: float):
        check(_lib.lib().eut_field_fill_column(self._h, int(col), float(value)))

    # -- halo support for slab sharding (eutropia_b200/slab.py) ---------------------------------
    def scale_column(self, col: int, factor: float):
        """column *= factor (exoenzyme decay, R/run_simulation.R:367-371)."""
        check(_lib.lib().eut_field_scale_column(self._h, int(col), float(factor)))

    def exoenzyme_catalysis(self, exo: "Field", exo_col: int, kcat: float, km: float, met_cols, stoich,
                            is_constant, delta_time: float):
        """R/run_simulation.R:343-364 on the device; `self` is the concentration field."""
        mets = _i32(met_cols)
        st = _f64c(stoich)
        isc = None if is_constant is None else np.ascontiguousarray(is_constant, dtype=np.uint8)
        check(_lib.lib().eut_field_exoenzyme_catalysis(self._h, exo._h, int(exo_col), float(kcat), float(km),
                                                       ptr(mets), ptr(st), mets.shape[0], ptr(isc),
                                                       float(delta_time)))

    def halo_set(self, send_ids, recv_ids):
        s, r = _i32(send_ids), _i32(recv_ids)
        check(_lib.lib().eut_field_halo_set(self._h, ptr(s), s.shape[0], ptr(r), r.shape[0]))

    def set_halo_stream(self, stream_ptr):
        check(_lib.lib().eut_field_set_halo_stream(self._h, C.c_void_p(stream_ptr) if stream_ptr else None))

    def halo_pack(self, cols, dev_ptr: int):
        c = _i32(cols)
        check(_lib.lib().eut_field_halo_pack(self._h, ptr(c), c.shape[0], C.c_void_p(dev_ptr)))

    def halo_unpack(self, cols, dev_ptr: int):
        c = _i32(cols)
        check(_lib.lib().eut_field_halo_unpack(self._h, ptr(c), c.shape[0], C.c_void_p(dev_ptr)))


class Cells:
    """Ragged cell -> field lists on the device, with gather and scatter against a `Field`."""

    def __init__(self, plan: Plan):
        self.plan = plan
        h = C.c_void_p()
        check(_lib.lib().eut_cells_create(C.byref(h), plan._h))
        self._h = h
        self.n_cells = 0

    def close(self):
        if getattr(self, "_h", None) and getattr(_lib, "lib", None):
            _lib.lib().eut_cells_destroy(self._h)
            self._h = None

    __del__ = close

    def set_lists(self, cell_ptr, field_id, field_dist, acc_prop):
        cp = np.ascontiguousarray(cell_ptr, dtype=np.int64)
        fid, fd, acc = _i32(field_id), _f64c(field_dist), _f64c(acc_prop)
        self.n_cells = cp.shape[0] - 1
        check(_lib.lib().eut_cells_set_lists(self._h, self.n_cells, ptr(cp), ptr(fid), ptr(fd), ptr(acc)))

    def build_lists(self, pts, cx, cy, csize, cscv, dist_mode=0):
        p = _colmajor(pts)
        cx, cy, csize, cscv = (_f64c(a) for a in (cx, cy, csize, cscv))
        self.n_cells = cx.shape[0]
        check(_lib.lib().eut_cells_build_lists(self._h, ptr(p), p.shape[0], ptr(cx), ptr(cy), ptr(csize),
                                               ptr(cscv), self.n_cells, int(dist_mode)))

    def claim_rescale(self):
        check(_lib.lib().eut_cells_claim_rescale(self._h))

    @property
    def n_entries(self) -> int:
        return int(_lib.lib().eut_cells_num_entries(self._h))

    def get_lists(self):
        t = self.n_entries
        cp = np.zeros(self.n_cells + 1, dtype=np.int64)
        fid = np.zeros(max(t, 1), dtype=np.int32)
        fd = np.zeros(max(t, 1)); acc = np.zeros(max(t, 1))
        check(_lib.lib().eut_cells_get_lists(self._h, ptr(cp), ptr(fid), ptr(fd), ptr(acc)))
        return cp, fid[:t], fd[:t], acc[:t]

    def gather(self, field: Field, field_vol: float, want_host=True):
        out = np.zeros((self.n_cells, field.n_cols), dtype=np.float64, order="F") if want_host else None
        check(_lib.lib().eut_cells_gather(self._h, field._h, float(field_vol), ptr(out)))
        return out

    def chemotaxis_moments(self, field: Field, col: int, pts, cell_x, cell_y, cell_r):
        """M x 7 sums per cell for compound column `col` (see eut_cells_chemotaxis_moments)."""
        p = np.require(np.asarray(pts, dtype=np.float64), requirements=["F", "A"])
        cx, cy, cr = _f64c(cell_x), _f64c(cell_y), _f64c(cell_r)
        out = np.zeros((max(cx.shape[0], 1), 7), dtype=np.float64, order="F")
        check(_lib.lib().eut_cells_chemotaxis_moments(self._h, field._h, int(col), ptr(p), p.shape[0], ptr(cx), ptr(cy),
                                                      ptr(cr), ptr(out)))
        return out[:cx.shape[0]]

    def scatter(self, field: Field, field_vol: float, is_constant, ex_ptr, ex_cpd, ex_flx):
        isc = None if is_constant is None else np.ascontiguousarray(is_constant, dtype=np.uint8)
        ep = np.ascontiguousarray(ex_ptr, dtype=np.int64)
        ec, ef = _i32(ex_cpd), _f64c(ex_flx)
        check(_lib.lib().eut_cells_scatter(self._h, field._h, float(field_vol), ptr(isc), ptr(ep),
                                           ptr(ec), ptr(ef)))


class SlabField:
    """One N x C field sharded by spatial slab over `devices` (one process, NVLink peer stores for the
    halo): `eut_slab_*`.  Bit-identical to the unsharded `Field`."""

    def __init__(self, mat_in, mat_out, n_cols: int, devices, n_points=None):
        mat_in = np.asarray(mat_in).reshape(-1, 2)
        mat_out = np.asarray(mat_out).reshape(-1, 2)
        frm, to = _i32(mat_in[:, 0]), _i32(mat_in[:, 1])
        nid, cnt = _i32(mat_out[:, 0]), _i32(mat_out[:, 1])
        self.n_points = int(mat_out.shape[0]) if n_points is None else int(n_points)
        self.n_cols = int(n_cols)
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        check(_lib.lib().eut_slab_create(C.byref(h), ptr(frm), ptr(to), frm.shape[0], ptr(nid), ptr(cnt), nid.shape[0],
                                         self.n_points, self.n_cols, ptr(dv), dv.shape[0]))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(_lib, "lib", None):
            _lib.lib().eut_slab_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def info(self) -> _lib.SlabInfo:
        inf = _lib.SlabInfo()
        check(_lib.lib().eut_slab_get_info(self._h, C.byref(inf)))
        return inf

    def part(self, index: int):
        """(device, points held, points owned, cudaStream_t) of one slab."""
        dev, nl, no, st = C.c_int32(), C.c_int64(), C.c_int64(), C.c_void_p()
        check(_lib.lib().eut_slab_part(self._h, int(index), C.byref(dev), C.byref(nl), C.byref(no), C.byref(st)))
        return dev.value, nl.value, no.value, int(st.value or 0)

    def upload(self, conc):
        a = _colmajor(conc)
        if a.shape != (self.n_points, self.n_cols):
            raise ValueError("conc must be N x C")
        check(_lib.lib().eut_slab_upload(self._h, ptr(a), max(self.n_points, 1)))

    def download(self):
        out = np.empty((self.n_points, self.n_cols), dtype=np.float64, order="F")
        check(_lib.lib().eut_slab_download(self._h, ptr(out), max(self.n_points, 1)))
        return out

    def diffuse(self, niter, indvar):
        ni = _i32(np.trunc(np.asarray(niter, dtype=np.float64)))
        iv = _i32(np.trunc(np.asarray(indvar, dtype=np.float64)))
        if ni.shape[0] < iv.shape[0]:
            raise _lib.EutropiaIndexError(_lib.ERR_INDEX, "niter is shorter than indvar")
        check(_lib.lib().eut_slab_diffuse(self._h, ptr(ni), ptr(iv), iv.shape[0]))

    def sync(self):
        check(_lib.lib().eut_slab_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(_lib.lib().eut_slab_launch_count(self._h))


class SlabCells:
    """Cells on a slab-sharded field (`eut_slab_cells_*`): same calls as `Cells`, the field is the `SlabField`
    given at construction.  Per-field work runs on the slab that owns the point, per-cell quantities are
    combined over the slabs a cell touches."""

    def __init__(self, slab: SlabField):
        self.slab = slab
        self.n_cells = 0
        h = C.c_void_p()
        check(_lib.lib().eut_slab_cells_create(C.byref(h), slab._h))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(_lib, "lib", None) and getattr(self.slab, "_h", None):
            _lib.lib().eut_slab_cells_destroy(self._h)
        self._h = None

    __del__ = close

    def build_lists(self, pts, cx, cy, csize, cscv, dist_mode=0):
        p = _colmajor(pts)
        cx, cy, csize, cscv = (_f64c(a) for a in (cx, cy, csize, cscv))
        self.n_cells = cx.shape[0]
        check(_lib.lib().eut_slab_cells_build_lists(self._h, ptr(p), p.shape[0], p.shape[0], ptr(cx), ptr(cy), ptr(csize),
                                                    ptr(cscv), self.n_cells, int(dist_mode)))

    def claim_rescale(self):
        check(_lib.lib().eut_slab_cells_claim_rescale(self._h))

    @property
    def n_entries(self) -> int:
        return int(_lib.lib().eut_slab_cells_num_entries(self._h))

    def get_lists(self, slab_index: int):
        """(cell_ptr, global field ids, distances, acc.prop) of the entries one slab holds."""
        n = C.c_int64()
        check(_lib.lib().eut_slab_cells_get_lists(self._h, int(slab_index), C.byref(n), None, None, None, None))
        t = int(n.value)
        cp = np.zeros(self.n_cells + 1, dtype=np.int64)
        fid = np.zeros(max(t, 1), dtype=np.int32)
        fd = np.zeros(max(t, 1)); acc = np.zeros(max(t, 1))
        check(_lib.lib().eut_slab_cells_get_lists(self._h, int(slab_index), C.byref(n), ptr(cp), ptr(fid), ptr(fd), ptr(acc)))
        return cp, fid[:t], fd[:t], acc[:t]

    def gather(self, field_vol: float, want_host=True):
        out = np.zeros((self.n_cells, self.slab.n_cols), dtype=np.float64, order="F") if want_host else None
        check(_lib.lib().eut_slab_cells_gather(self._h, float(field_vol), ptr(out)))
        return out

    def scatter(self, field_vol: float, is_constant, ex_ptr, ex_cpd, ex_flx):
        isc = None if is_constant is None else np.ascontiguousarray(is_constant, dtype=np.uint8)
        ep = np.ascontiguousarray(ex_ptr, dtype=np.int64)
        ec, ef = _i32(ex_cpd), _f64c(ex_flx)
        check(_lib.lib().eut_slab_cells_scatter(self._h, float(field_vol), ptr(isc), ptr(ep), ptr(ec), ptr(ef)))
