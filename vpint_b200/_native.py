"""ctypes binding of libvpint_b200.so (C ABI in include/vpint_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (or
``make -C vpint_b200/csrc``).  There is no CPU fallback: a missing library or a
missing CUDA device raises :class:`VPintError` as soon as a solve is attempted.
"""

from __future__ import annotations

import ctypes as C
import os

from .errors import VPintError

LIB_NAME = "libvpint_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = 0, 1, 2, 3, 4
F64, F32, U16, U8 = 0, 1, 2, 3
FILL_DEVICE, FILL_READY, FILL_HOST = 0, 1, 2
ABI_VERSION = 2
W_SCALAR, W_PLANES = 0, 1
METHOD_EXACT, METHOD_COSINE, METHOD_EXACT_INVERSE = 0, 1, 2
LOOP_CHUNKED, LOOP_GRAPH = 0, 1
FEAT_RESISTANCE = 1
FEAT_BLOCKING_WAIT = 2

METHODS = {"exact": METHOD_EXACT, "cosine_similarity": METHOD_COSINE, "exact_inverse": METHOD_EXACT_INVERSE}


class Desc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("nprob", C.c_int32), ("nprob_inner", C.c_int32),
        ("H", C.c_int32), ("W", C.c_int32), ("T", C.c_int32),
        ("storage", C.c_int32), ("weight_mode", C.c_int32), ("k_block", C.c_int32),
        ("global_H", C.c_int32), ("row_offset", C.c_int32), ("own_row0", C.c_int32), ("own_rows", C.c_int32),
        ("features", C.c_int32),
    ]


class RunParams(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32), ("auto_terminate", C.c_int32),
        ("threshold", C.c_double), ("clip_val", C.c_double),
        ("loop_mode", C.c_int32), ("chunk", C.c_int32),
        ("delta_out", C.c_void_p),
        ("resistance", C.c_int32), ("epsilon", C.c_double), ("mu", C.c_double),
    ]


class IngestParams(C.Structure):
    _fields_ = [
        ("target", C.c_void_p), ("features", C.c_void_p),
        ("raster_dtype", C.c_int32), ("value_dtype", C.c_int32),
        ("cloud", C.c_void_p),
        ("clip_target", C.c_int32), ("clip_features", C.c_int32),
        ("target_min", C.c_double), ("target_max", C.c_double), ("features_min", C.c_double), ("features_max", C.c_double),
        ("fill_mode", C.c_int32),
        ("fill_host", C.POINTER(C.c_double)),
        ("scratch", C.c_void_p),
        ("out", C.c_void_p),
        ("out_dtype", C.c_int32),
        ("out_stride", C.c_int64 * 4),
    ]


# name -> (restype, argtypes); every symbol include/vpint_b200.h declares
SIGNATURES = {
    "vpint_abi_version": (C.c_int, []),
    "vpint_last_error": (C.c_char_p, []),
    "vpint_row_pitch": (C.c_int64, [C.c_int64]),
    "vpint_solver_create": (C.c_int, [C.POINTER(Desc), C.POINTER(C.c_void_p)]),
    "vpint_solver_destroy": (None, [C.c_void_p]),
    "vpint_solver_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "vpint_solver_bind": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "vpint_solver_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int,
                                     C.POINTER(C.c_int64), C.c_void_p]),
    "vpint_solver_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_int,
                                        C.c_int, C.c_double, C.c_double, C.c_void_p]),
    "vpint_solver_priority": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpint_solver_set_discounts": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]),
    "vpint_solver_set_trial_params": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                 C.c_void_p]),
    "vpint_solver_run_async": (C.c_int, [C.c_void_p, C.POINTER(RunParams), C.c_void_p]),
    "vpint_solver_run_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "vpint_cloud_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_void_p, C.c_void_p]),
    "vpint_buffer_clouds": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p]),
    "vpint_fill_scratch_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "vpint_fill_leaf_count": (C.c_int64, [C.c_int64, C.c_int]),
    "vpint_fill_count_offset": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "vpint_solver_fill_leaves": (C.c_int, [C.c_void_p, C.POINTER(IngestParams), C.c_void_p]),
    "vpint_solver_fill_combine": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vpint_solver_fill_ptr": (C.c_void_p, [C.c_void_p]),
    "vpint_solver_ingest_bands": (C.c_int, [C.c_void_p, C.POINTER(IngestParams), C.c_void_p]),
    "vpint_solver_result_unknown": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]),
    "vpint_solver_run": (C.c_int, [C.c_void_p, C.POINTER(RunParams), C.c_void_p]),
    "vpint_solver_begin_run": (C.c_int, [C.c_void_p, C.POINTER(RunParams), C.c_void_p]),
    "vpint_solver_step_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vpint_solver_step_end": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vpint_solver_step_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "vpint_solver_sums_ptr": (C.c_void_p, [C.c_void_p]),
    "vpint_solver_sums_count": (C.c_int, [C.c_void_p]),
    "vpint_solver_state_ptr": (C.c_void_p, [C.c_void_p, C.c_int]),
    "vpint_solver_active_ptr": (C.c_void_p, [C.c_void_p]),
    "vpint_solver_cur_ptr": (C.c_void_p, [C.c_void_p]),
    "vpint_solver_known_ptr": (C.c_void_p, [C.c_void_p]),
    "vpint_solver_known_count": (C.c_int, [C.c_void_p]),
    "vpint_solver_any_dirty": (C.c_int, [C.c_void_p]),
    "vpint_solver_halo_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpint_solver_halo_unpack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpint_solver_step_commit_lagged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpint_solver_halo_push": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vpint_solver_halo_pull": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vpint_solver_peer_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "vpint_solver_peer_bind": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vpint_solver_run_sharded": (C.c_int, [C.c_void_p, C.POINTER(RunParams), C.c_void_p, C.POINTER(C.c_int64)]),
    "vpint_solver_result": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]),
    "vpint_solver_active_tiles": (C.c_int64, [C.c_void_p]),
    "vpint_solver_total_tiles": (C.c_int64, [C.c_void_p]),
    "vpint_solver_launches": (C.c_int64, [C.c_void_p]),
    "vpint_solver_k_block": (C.c_int, [C.c_void_p]),
}

_lib = None


def load_library(path=None):
    """dlopen the C-ABI library and attach prototypes.  Loud on failure."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise VPintError(
            "vpint_b200: CUDA library %s not found; build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C vpint_b200/csrc` (there is no CPU fallback)" % p)
    try:
        lib = C.CDLL(p)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise VPintError("vpint_b200: cannot load %s: %s" % (p, exc)) from exc
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.vpint_abi_version() != ABI_VERSION:
        raise VPintError("vpint_b200: ABI version mismatch in %s" % p)
    if path is None:
        _lib = lib
    return lib


def check(lib, rc, what):
    if rc != OK:
        msg = lib.vpint_last_error()
        raise VPintError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else ""))
