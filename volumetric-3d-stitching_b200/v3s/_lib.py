"""ctypes binding of libv3s.so (the C-ABI declared in include/v3s.h).

The product path has no CPU fallback: importing this module without the built library, or
calling a kernel without an sm_100a device, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libv3s.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "v3s: libv3s.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C volumetric-3d-stitching_b200/csrc`; there is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

p, i32, i64, f64 = C.c_void_p, C.c_int, C.c_longlong, C.c_double


class Geometry(C.Structure):
    _fields_ = [("pixel", f64), ("zd", f64), ("lambda_", f64), ("n_p_nobin", i32), ("grid_scale", f64)]


class EigProblem(C.Structure):
    """v3s_eig_problem (include/v3s.h)."""
    _fields_ = [("P_rows", C.c_void_p), ("ldp", C.c_longlong), ("sparse", C.c_void_p), ("n_rows", C.c_longlong),
                ("N", C.c_longlong), ("row0", C.c_longlong), ("rank", C.c_int), ("world", C.c_int),
                ("ybuf_ptrs", C.c_void_p), ("xf_ptrs", C.c_void_p), ("flag_ptrs", C.c_void_p), ("ticket_dev", C.c_void_p),
                ("status_dev", C.c_void_p), ("mv_workspace", C.c_void_p), ("epoch", C.c_int)]


class SparseMarkovDesc(C.Structure):
    """v3s_sparse_markov (include/v3s.h): device pointers of the sparse operator's arrays."""
    _fields_ = [("nbr", p), ("a_val", p), ("t_ptr", p), ("t_col", p), ("t_val", p), ("k", i32),
                ("row_order", p), ("order_row0", i64), ("order_rows", i64)]


# name -> (restype, argtypes); mirrors include/v3s.h one to one
SIGNATURES = {
    "v3s_last_error": (C.c_char_p, []),
    "v3s_version": (i32, []),
    "v3s_kernel_launches": (i64, []),
    "v3s_check_device": (i32, []),
    "v3s_peer_fill": (i32, [p, C.c_float, i64, p]),
    "v3s_default_geometry": (None, [C.POINTER(Geometry)]),
    "v3s_synth_workspace_bytes": (i64, [i32]),
    "v3s_synth_prepare": (i32, [i32, i32, C.POINTER(Geometry), p, C.POINTER(i32), p]),
    "v3s_synth_patterns": (i32, [p, i32, p, i64, i32, p, i32, i32, p, i64, p]),
    "v3s_synth_patterns_direct": (i32, [p, i32, p, i64, i32, C.POINTER(Geometry), p, p, i64, p]),
    "v3s_poisson_noise": (i32, [p, i64, i32, i64, i64, f64, C.c_ulonglong, p]),
    "v3s_column_sums": (i32, [p, i64, i32, i64, p, p]),
    "v3s_center_normalize": (i32, [p, i64, i32, i64, p, p, p, p, i32, p, p, p]),
    "v3s_center_normalize_f16": (i32, [p, i64, i32, i64, p, p, p, i32, p, p]),
    "v3s_split_bf16": (i32, [p, i64, i32, i32, p, p, p]),
    "v3s_gram_rows": (i32, [p, p, i64, p, p, i64, i32, p, i64, p]),
    "v3s_set_gram_kblock": (i32, [i32]),
    "v3s_gram_symmetric": (i32, [p, p, i64, i32, p, i64, p, i32, p]),
    "v3s_gram_symmetric_sharded": (i32, [p, p, i64, i32, C.POINTER(C.c_ulonglong), i32, i64, i32, i64, p, i32, p]),
    "v3s_gram_rows_simt": (i32, [p, i64, p, i64, i32, p, i64, p]),
    "v3s_round_distance_dense": (i32, [p, i64, i32, p, p, p]),
    "v3s_knn_from_gram": (i32, [p, i64, i64, i64, i64, p, p, i32, i32, i32, C.c_float, i32, p, p, p, p, p, p, p, p]),
    "v3s_split_f16": (i32, [p, i64, i32, i32, p, p]),
    "v3s_set_gram_topk_cta_group": (i32, [i32]),
    "v3s_set_gram_reserved_sms": (i32, [i32]),
    "v3s_set_gram_sync_lag": (i32, [i32]),
    "v3s_gram_topk_plan": (i32, [i64, i64, i32, C.POINTER(i32)]),
    "v3s_gram_topk": (i32, [p, i64, p, i64, i32, i32, C.c_float, p, p, p, p, p]),
    "v3s_gram_topk_limits": (i32, [p, i64, p, i64, i32, i32, C.c_float, i32, p, p, p, p, p, p]),
    "v3s_gram_topk_sym_steps": (i32, [i64, i32]),
    "v3s_gram_topk_sym": (i32, [p, i64, i32, p, i32, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), i32, i32, i64, i32, p, p]),
    "v3s_knn_merge_lists": (i32, [p, p, i32, i32, i64, i64, i64, i32, C.c_float, i32, p, p, p, p, p, i32, p, i32, i32, p, p]),
    "v3s_knn_merge_segment_lists": (i32, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), i32, i32, i64, i64, i64, i32, C.c_float,
                                          i32, p, p, p, p, p, i32, p, i32, i32, p, p]),
    "v3s_knn_refine_tile_rows": (i32, [i32]),
    "v3s_knn_refine_tiled": (i32, [p, p, i32, i64, i64, i64, i32, i32, p, p, p, p, p, p, p, p, i32, p, p, p, i32, p]),
    "v3s_knn_refine_tiled_pairs": (i32, [p, p, i32, i64, i64, i64, i32, p, p, p, p, p, i32, p, p, p, i32, p, p, p]),
    "v3s_knn_refine_tiled_emit": (i32, [i64, i64, i64, i32, i32, p, p, p, p, p, p, p, C.POINTER(C.c_ulonglong), i32, i64, i32, p]),
    "v3s_knn_refine": (i32, [p, p, i32, i64, i64, i64, i32, i32, p, p, p, p, p, p, p, p]),
    "v3s_knn_from_dense": (i32, [p, i64, i32, i32, p, p, p, p, p, p]),
    "v3s_s2_scalars": (i32, [p, i64, i32, f64, i32, i32, p, p]),
    "v3s_s2_threshold": (i32, [p, i64, i32, p, p]),
    "v3s_build_markov": (i32, [p, p, i64, i32, p, i64, i64, p, i64, p, p, p]),
    "v3s_w_dense": (i32, [p, p, i64, i32, f64, p, p]),
    "v3s_anisotropic_norm_dense": (i32, [p, i64, p, p, p]),
    "v3s_dm_cov_dense": (i32, [p, i64, p, p, p, p]),
    "v3s_block_matvec_workspace_bytes": (i64, [i64, i64]),
    "v3s_block_matvec": (i32, [p, i64, i64, i64, p, i32, p, p, p]),
    "v3s_xf_floats": (i64, [i64]),
    "v3s_block_matvec_f32": (i32, [p, i64, i64, i64, p, p, p, p]),
    "v3s_lanczos_workspace_doubles": (i64, [i64, i32]),
    "v3s_lanczos_start": (i32, [p, i64, i64, i32, p, p, p, p]),
    "v3s_lanczos_step": (i32, [p, i64, i64, i32, i32, p, p, p, p, p, p]),
    "v3s_cheb_combine": (i32, [p, p, p, f64, f64, f64, p, p, i64, p]),
    "v3s_matvec_combine": (i32, [p, i64, i64, i64, i64, p, p, p, f64, f64, f64, C.POINTER(C.c_ulonglong),
                                 C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), i32, i32, i32, p, p, p, p]),
    "v3s_peer_barrier": (i32, [C.POINTER(C.c_ulonglong), i32, i32, i32, p, p]),
    "v3s_cheb_filter": (i32, [p, i64, i64, i64, i64, p, p, i32, i32, f64, f64, C.POINTER(C.c_ulonglong),
                              C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), i32, i32, i32, p, p, p, C.POINTER(i32), p]),
    "v3s_build_markov_sparse": (i32, [p, p, i64, i32, p, p, p, p, p, p, p, p, p]),
    "v3s_sparse_block_matvec": (i32, [C.POINTER(SparseMarkovDesc), i64, i64, i64, p, i32, p, p]),
    "v3s_sparse_matvec_combine": (i32, [C.POINTER(SparseMarkovDesc), i64, i64, i64, p, p, p, f64, f64, f64,
                                        C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong),
                                        i32, i32, i32, p, p, p, p]),
    "v3s_sparse_cheb_filter": (i32, [C.POINTER(SparseMarkovDesc), i64, i64, i64, p, i32, i32, f64, f64,
                                     C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), i32, i32, i32, p, p, p,
                                     C.POINTER(i32), p]),
    "v3s_eig_topk_workspace_doubles": (i64, [i64, i32]),
    "v3s_eig_topk": (i32, [C.POINTER(EigProblem), i32, f64, i32, i32, i32, C.c_ulonglong, p, C.POINTER(f64), p, C.POINTER(f64),
                           C.POINTER(i32), p]),
    "v3s_sym_eig_host": (i32, [C.POINTER(f64), i32, C.POINTER(f64)]),
    "v3s_profile_matvec": (i32, [i32]),
    "v3s_profile_matvec_read": (i32, [C.POINTER(C.c_float), i32]),
    "v3s_read_small": (i32, [p, p, i64, p]),
    "v3s_profile_matvec_gaps": (i32, [C.POINTER(C.c_float), i32]),
    "v3s_fit_gram": (i32, [p, i32, p, i64, p, p]),
    "v3s_fit_solve": (i32, [p, p, p]),
    "v3s_fit_project": (i32, [p, i32, p, i64, p, i32, p, p, p, p, p]),
    "v3s_or_func": (i32, [p, p, i32, i64, i64, i32, f64, p, p, p]),
    "v3s_sigma_pairs": (i32, [p, p, i64, i64, i64, p, p, p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

launches = 0   # number of C-ABI compute calls issued (bench.py reports kernels per step from it)


def last_error():
    return lib.v3s_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Call a status-returning entry point; raise Exception(message) like the reference's
    Utilities.check (utils.py:91-93) when it fails."""
    global launches
    launches += 1
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise Exception(last_error())


_device_ok = False


def require_device():
    """Fail loudly unless torch sees a CUDA device and it is sm_100a."""
    global _device_ok
    if _device_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("v3s: no CUDA device visible; the v3s stages run on a B200 only (no CPU fallback)")
    torch.cuda.current_device()
    if lib.v3s_check_device() != 0:
        raise RuntimeError(last_error())
    _device_ok = True


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())
