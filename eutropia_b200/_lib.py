"""ctypes binding of `libeutropia_b200.so` (the C ABI declared in include/eutropia_b200.h).

There is no Python or CPU implementation of the compute path behind this module: if the shared
library is missing, or there is no CUDA device, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeutropia_b200.so")

OK, ERR_INDEX, ERR_ARG, ERR_ALLOC, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_NCCL = range(8)

ORDER_AUTO, ORDER_IDENTITY, ORDER_ROWBLOCK, ORDER_TILED, ORDER_SEGMENTS = 0, 1, 2, 3, 4
PLAN_FORCE_GENERIC = 0x10
PLAN_HOST_ONLY = 0x20


class EutropiaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"eutropia_b200 error {code}: {message}")
        self.code = code


class EutropiaIndexError(EutropiaError, IndexError):
    """An index outside 1..N / 1..C -- where the reference raises through Armadillo's bounds check."""


class PlanInfo(C.Structure):
    _fields_ = [
        ("n_points", C.c_int64), ("n_edges", C.c_int64), ("n_padded", C.c_int64),
        ("max_in_degree", C.c_int32), ("ell_width", C.c_int32), ("order", C.c_int32),
        ("regular_outflow", C.c_int32), ("device", C.c_int32), ("n_tiles", C.c_int32),
        ("device_bytes", C.c_int64), ("build_ms", C.c_double),
        ("tiled", C.c_int32), ("img_points", C.c_int32), ("max_tile_points", C.c_int32),
        ("max_bulk", C.c_int32), ("max_single", C.c_int32), ("n_copies", C.c_int32),
        ("halo_ratio", C.c_double),
        ("segmented", C.c_int32), ("seg_points", C.c_int32), ("max_tile_segs", C.c_int32),
        ("max_tile_generic", C.c_int32), ("n_segments", C.c_int64), ("n_generic", C.c_int64),
        ("n_seg_covered", C.c_int64),
    ]


class SlabInfo(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_cols", C.c_int64), ("n_slabs", C.c_int32), ("n_links", C.c_int32),
                ("max_local", C.c_int64), ("halo_points", C.c_int64), ("max_link", C.c_int64), ("pushes", C.c_int64),
                ("segmented", C.c_int32), ("depth", C.c_int32), ("build_ms", C.c_double)]


_lib = None


def lib() -> C.CDLL:
    """Load the shared library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C eutropia_b200/csrc`. eutropia_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    sig = {
        "eut_last_error": (C.c_char_p, []),
        "eut_abi_version": (C.c_int, []),
        "eut_device_count": (C.c_int, []),
        "eut_diffchange": (C.c_int, [vp, i64, vp, i64, vp, i64, i64, vp, i64, vp, C.c_int]),
        "eut_diffchange_par": (C.c_int, [vp, i64, vp, i64, vp, i64, i64, vp, vp, i64, C.c_int, vp, C.c_int]),
        "eut_diffchange_par_multi": (C.c_int, [vp, i64, vp, i64, vp, i64, i64, vp, vp, i64, C.c_int, vp, vp, C.c_int]),
        "eut_deal_columns": (C.c_int, [vp, vp, i64, C.c_int, vp]),
        "eut_oneshot_reset": (None, []),
        "eut_build_dcm": (C.c_int, [vp, i64, i64, dbl, C.c_int, vp, i64, C.POINTER(i64), vp]),
        "eut_plan_create": (C.c_int, [C.POINTER(vp), vp, vp, i64, vp, vp, i64, i64, C.c_int, C.c_uint32]),
        "eut_plan_destroy": (C.c_int, [vp]),
        "eut_plan_get_info": (C.c_int, [vp, C.POINTER(PlanInfo)]),
        "eut_plan_get_csr": (C.c_int, [vp, vp, vp]),
        "eut_plan_get_perm": (C.c_int, [vp, vp]),
        "eut_plan_get_tiles": (C.c_int, [vp, vp, vp, vp]),
        "eut_plan_get_segs": (C.c_int, [vp, vp, vp, vp, vp]),
        "eut_field_create": (C.c_int, [C.POINTER(vp), vp, i64]),
        "eut_field_destroy": (C.c_int, [vp]),
        "eut_field_upload": (C.c_int, [vp, vp, i64, vp, i64]),
        "eut_field_download": (C.c_int, [vp, vp, i64, vp, i64]),
        "eut_field_record": (C.c_int, [vp, vp, i64, vp, i64, C.c_int]),
        "eut_debug_tile_ns": (C.c_int, [vp, C.c_int]),
        "eut_field_diffuse": (C.c_int, [vp, vp, vp, i64]),
        "eut_field_sync": (C.c_int, [vp]),
        "eut_field_stream": (vp, [vp]),
        "eut_field_set_stream": (C.c_int, [vp, vp]),
        "eut_field_launch_count": (i64, [vp]),
        "eut_field_set_option": (C.c_int, [vp, C.c_char_p, i64]),
        "eut_field_column_stats": (C.c_int, [vp, vp, vp, vp]),
        "eut_field_fill_column": (C.c_int, [vp, i32, dbl]),
        "eut_field_exoenzyme_catalysis": (C.c_int, [vp, vp, i32, dbl, dbl, vp, vp, i32, vp, dbl]),
        "eut_field_scale_column": (C.c_int, [vp, i32, dbl]),
        "eut_field_halo_set": (C.c_int, [vp, vp, i64, vp, i64]),
        "eut_field_set_halo_stream": (C.c_int, [vp, vp]),
        "eut_field_halo_pack": (C.c_int, [vp, vp, i64, vp]),
        "eut_field_halo_unpack": (C.c_int, [vp, vp, i64, vp]),
        "eut_slab_create": (C.c_int, [C.POINTER(vp), vp, vp, i64, vp, vp, i64, i64, i64, vp, C.c_int]),
        "eut_slab_destroy": (C.c_int, [vp]),
        "eut_slab_get_info": (C.c_int, [vp, C.POINTER(SlabInfo)]),
        "eut_slab_part": (C.c_int, [vp, C.c_int, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), C.POINTER(vp)]),
        "eut_slab_upload": (C.c_int, [vp, vp, i64]),
        "eut_slab_download": (C.c_int, [vp, vp, i64]),
        "eut_slab_diffuse": (C.c_int, [vp, vp, vp, i64]),
        "eut_slab_sync": (C.c_int, [vp]),
        "eut_slab_launch_count": (i64, [vp]),
        "eut_slab_cells_create": (C.c_int, [C.POINTER(vp), vp]),
        "eut_slab_cells_destroy": (C.c_int, [vp]),
        "eut_slab_cells_build_lists": (C.c_int, [vp, vp, i64, i64, vp, vp, vp, vp, i64, C.c_int]),
        "eut_slab_cells_claim_rescale": (C.c_int, [vp]),
        "eut_slab_cells_num_entries": (i64, [vp]),
        "eut_slab_cells_get_lists": (C.c_int, [vp, C.c_int, C.POINTER(i64), vp, vp, vp, vp]),
        "eut_slab_cells_gather": (C.c_int, [vp, C.c_double, vp]),
        "eut_slab_cells_scatter": (C.c_int, [vp, C.c_double, vp, vp, vp, vp]),
        "eut_cells_create": (C.c_int, [C.POINTER(vp), vp]),
        "eut_cells_destroy": (C.c_int, [vp]),
        "eut_cells_set_lists": (C.c_int, [vp, i64, vp, vp, vp, vp]),
        "eut_cells_build_lists": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, i64, C.c_int]),
        "eut_cells_claim_rescale": (C.c_int, [vp]),
        "eut_cells_num_entries": (i64, [vp]),
        "eut_cells_get_lists": (C.c_int, [vp, vp, vp, vp, vp]),
        "eut_cells_gather": (C.c_int, [vp, vp, dbl, vp]),
        "eut_cells_scatter": (C.c_int, [vp, vp, dbl, vp, vp, vp, vp]),
        "eut_cells_chemotaxis_moments": (C.c_int, [vp, vp, i32, vp, i64, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTS = (
    "eut_last_error eut_abi_version eut_device_count eut_diffchange eut_diffchange_par eut_diffchange_par_multi eut_deal_columns "
    "eut_oneshot_reset eut_build_dcm eut_plan_create eut_plan_destroy eut_plan_get_info eut_plan_get_csr "
    "eut_plan_get_perm eut_plan_get_tiles eut_plan_get_segs eut_field_create eut_field_destroy eut_field_upload eut_field_download eut_field_record "
    "eut_field_diffuse eut_field_sync eut_field_stream eut_field_set_stream "
    "eut_field_launch_count eut_field_set_option eut_debug_tile_ns eut_field_column_stats eut_field_fill_column "
    "eut_field_exoenzyme_catalysis eut_field_scale_column "
    "eut_field_halo_set eut_field_set_halo_stream eut_field_halo_pack eut_field_halo_unpack "
    "eut_slab_create eut_slab_destroy eut_slab_get_info eut_slab_part eut_slab_upload eut_slab_download "
    "eut_slab_diffuse eut_slab_sync eut_slab_launch_count "
    "eut_slab_cells_create eut_slab_cells_destroy eut_slab_cells_build_lists eut_slab_cells_claim_rescale "
    "eut_slab_cells_num_entries eut_slab_cells_get_lists eut_slab_cells_gather eut_slab_cells_scatter "
    "eut_cells_create eut_cells_destroy eut_cells_set_lists eut_cells_build_lists "
    "eut_cells_claim_rescale eut_cells_num_entries eut_cells_get_lists eut_cells_gather "
    "eut_cells_scatter eut_cells_chemotaxis_moments"
).split()


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = lib().eut_last_error().decode("utf-8", "replace")
    if rc == ERR_INDEX:
        raise EutropiaIndexError(rc, msg)
    raise EutropiaError(rc, msg)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
