"""ctypes binding of ``libvipb200.so`` (the C ABI declared in ``include/vip_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised.  Build the library with ``python __graft_entry__.py`` (or
``make -C vip_b200/csrc``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvipb200.so")

_lib = None

_vp, _sz, _i, _d, _f = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_float

# name -> (restype, argtypes); mirrors include/vip_b200.h one to one
SIGNATURES = {
    "vb_version": (_i, []),
    "vb_last_error": (C.c_char_p, []),
    "vb_launch_count": (C.c_longlong, []),
    "vb_gram_workspace_bytes": (_sz, [_i, _sz]),
    "vb_gram_f32": (_i, [_vp, _i, _sz, _i, _vp, _vp, _sz, _vp]),
    "vb_upload_gram_f32": (_i, [_vp, _i, _sz, _vp, _vp, _vp, _sz, _i, _vp]),
    "vb_cross_gram_workspace_bytes": (_sz, [_i, _i]),
    "vb_cross_gram_f32": (_i, [_vp, _i, _vp, _i, _sz, _vp, _vp, _sz, _vp]),
    "vb_eigh_workspace_bytes": (_sz, [_i]),
    "vb_eigh_f64": (_i, [_vp, _i, _vp, _vp, _i, _d, _vp, _sz, C.POINTER(_i), _vp]),
    "vb_chol_whiten_f64": (_i, [_vp, _i, _vp, _vp]),
    "vb_eigh_topk_workspace_bytes": (_sz, [_i, _i]),
    "vb_eigh_topk_f64": (_i, [_vp, _i, _i, _d, _i, _vp, _vp, _vp, _sz, C.POINTER(_i), _vp]),
    "vb_eigh_topk_async_f64": (_i, [_vp, _i, _i, _d, _i, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vb_pcs_f32": (_i, [_vp, _vp, _i, _i, _sz, _vp, _vp]),
    "vb_pcs_hilo_f32": (_i, [_vp, _vp, _i, _i, _sz, _vp, _vp, _vp]),
    "vb_project_subtract_f32": (_i, [_vp, _vp, _i, _vp, _i, _i, _sz, _vp, _vp]),
    "vb_project_subtract_hp_f32": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _sz, _vp, _vp]),
    "vb_project_subtract_hp_rows_f32": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _sz, _vp, _vp, _vp]),
    "vb_sub_f32": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "vb_derotate_scratch_bytes": (_sz, [_i, _i, _i, _sz]),
    "vb_derotate_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _f, _i, _i, _vp, _sz, _i, _vp]),
    "vb_derotate_scatter_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _f, _i, _i, _vp, _sz, _vp, _i, _i,
                                    C.c_longlong, _i, _vp]),
    "vb_collapse_f32": (_i, [_vp, _i, _sz, _i, _vp, _i, _i, _vp, _vp]),
    "vb_annular_weights_f64": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _d, _i, _vp, _vp, _vp]),
    "vb_annular_direct_f64": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "vb_annular_auto_f64": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _d, _d, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "vb_gather_columns_f32": (_i, [_vp, _i, _sz, _vp, _i, _vp, _vp]),
    "vb_scatter_columns_f32": (_i, [_vp, _i, _i, _vp, _sz, _vp, _vp]),
    "vb_gemm_f32": (_i, [_vp, C.c_longlong, C.c_longlong, _i, _vp, C.c_longlong, C.c_longlong, _i, _i, _vp,
                        C.c_longlong, C.c_longlong, _i, _i, _i, _f, _f, _i, _vp]),
    "vb_split3_bf16": (_i, [_vp, C.c_longlong, _i, C.c_longlong, _vp, C.c_longlong, C.c_longlong, _vp]),
    "vb_gemm_bf16x3_tc": (_i, [_vp, C.c_longlong, C.c_longlong, _i, _vp, C.c_longlong, C.c_longlong, _i, _i, _i, _i,
                              _vp, C.c_longlong, C.c_longlong, _vp, C.c_longlong, C.c_longlong, _i, _vp]),
    "vb_shift_operators_f32": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "vb_checker_correct_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "vb_memcpy2d_h2d": (_i, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "vb_aperture_sums_f64": (_i, [_vp, _i, _i, _vp, _vp, _i, _d, _vp, _vp]),
    "vb_snr_points_f64": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _d, _d, _d, _i, _i, _vp, _vp, _vp]),
    "vb_fp32_probe": (_i, [_vp, _i, _i, _vp]),
    "vb_local_max_mask_f32": (_i, [_vp, _i, _i, _i, _f, _vp, _vp]),
    "vb_fits_decode_f32": (_i, [_vp, _i, _sz, _d, _d, _vp, _vp]),
    "vb_memcpy_h2d_staged": (_i, [_vp, _vp, _sz, _vp]),
    "vb_memcpy2d_h2d_staged": (_i, [_vp, _vp, _sz, _sz, _sz, _vp]),
    "vb_profile_enable": (None, [_i]),
    "vb_profile_read": (_i, [C.POINTER(_f)]),
}


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"vip_b200: CUDA library not found at {LIB_PATH}. This package has no CPU "
                "fallback; build it with `python __graft_entry__.py` (nvcc, sm_100a).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().vb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"vip_b200: {what} failed (rc={rc}): {msg}")


def launch_count():
    return int(lib().vb_launch_count())
