"""ctypes binding of libspinrelax_b200.so (the C ABI in include/spinrelax_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if that fails or no
CUDA device is present every compute call raises.
"""
import ctypes
import os

from . import build as _build

_c = ctypes
_LIB = None
ABI_VERSION = 2


class SpinRelaxError(RuntimeError):
    pass


def _declare(lib):
    ll, i, vp, sz = _c.c_longlong, _c.c_int, _c.c_void_p, _c.c_size_t
    dp = _c.POINTER(_c.c_double)
    sig = {
        "sr_abi_version": (i, []),
        "sr_last_error": (_c.c_char_p, []),
        "sr_device_info": (i, [_c.POINTER(i), _c.POINTER(i), _c.POINTER(i), _c.POINTER(sz)]),
        "sr_ct_row_pitch": (ll, [ll]),
        "sr_ct_workspace_bytes": (sz, [i, ll, i]),
        "sr_pack_vectors_f32": (i, [vp, i, ll, i, dp, vp, ll, vp]),
        "sr_ct_lag_sums": (i, [vp, ll, i, ll, i, ll, vp, vp]),
        "sr_pack_vectors_f32_chunks": (i, [vp, i, i, i, ll, i, dp, vp, ll, vp]),
        "sr_ct_lag_sums_chunks": (i, [vp, ll, i, i, i, ll, i, ll, vp, vp]),
        "sr_ct_palmer_finalize": (i, [vp, i, ll, i, ll, vp, vp, vp]),
        "sr_ct_palmer_device": (i, [vp, i, ll, i, vp, vp, vp, sz, vp]),
        "sr_ct_palmer_host": (i, [vp, i, ll, i, vp, vp]),
        "sr_release_host_cache": (None, []),
        "sr_sphere_hist_table_doubles": (i, [i, i]),
        "sr_vec_block_moments": (i, [vp, ll, i, ll, vp, vp]),
        "sr_rotate_vectors_f32_f64": (i, [vp, ll, dp, vp, vp]),
        "sr_xyz_to_rtp_f32": (i, [vp, ll, vp, i, vp]),
        "sr_xyz_to_rtp_f64": (i, [vp, ll, vp, i, vp]),
        "sr_jomega_f64": (i, [vp, vp, vp, ll, vp]),
        "sr_jomega_f32": (i, [vp, vp, vp, ll, vp]),
        "sr_jomega_host_f64": (i, [vp, ll, vp, ll, vp, ll, ll]),
        "sr_jomega_host_f32": (i, [vp, ll, vp, ll, vp, ll, ll]),
        "sr_relax_a_moments": (i, [vp, i, vp, i, i, vp, vp]),
        "sr_relax_eval": (i, [i, dp, _c.c_double, _c.c_double, _c.c_double, _c.c_double, _c.c_double, i, i, i, i, i,
                              vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "sr_ct_fit_workspace_bytes": (sz, [i, ll, i]),
        "sr_ct_fit_trf": (i, [vp, vp, vp, i, ll, i, vp, vp, vp, i, _c.c_double, _c.c_double, _c.c_double, vp, vp, vp, vp,
                              vp, vp, sz, vp]),
        "sr_dq_moments": (i, [vp, ll, vp, i, ll, i, vp, vp]),
        "sr_dq_moments_pooled": (i, [vp, ll, vp, i, ll, i, i, i, i, vp, vp]),
        "sr_dq_self": (i, [vp, ll, ll, vp, vp]),
        "sr_expdecay_chi2": (i, [vp, vp, ll, _c.c_double, _c.c_double, _c.c_double, vp, i, vp, vp]),
        "sr_dq_hist3d": (i, [vp, ll, ll, vp, i, vp, vp, i, vp, vp]),
        "sr_vec_second_moments": (i, [vp, ll, i, vp, vp]),
        "sr_sphere_hist": (i, [vp, ll, i, dp, i, i, vp, _c.c_double, _c.c_double, vp, vp, i, vp, vp]),
        "sr_sphere_hist_host": (i, [vp, ll, i, dp, i, i, vp, vp, i, vp]),
        "sr_xh_vectors": (i, [vp, ll, i, vp, vp, i, vp, vp]),
        "sr_xh_vectors_superposed": (i, [vp, ll, i, vp, vp, i, vp, vp, i, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sig


def lib_path():
    return _build.LIB


def load():
    """Load (building first if needed) the shared library. Raises if it cannot be produced."""
    global _LIB
    if _LIB is not None:
        return _LIB
    # build() compares the digest of csrc/ + the header with the stamp next to the .so and returns at once when they
    # match; after an edit it recompiles, so a stale binary is never loaded.  On a box without nvcc (the .so travels
    # with the snapshot) a matching stamp is required.
    path = _build.build()
    lib = _c.CDLL(path)
    _declare(lib)
    if lib.sr_abi_version() != ABI_VERSION:
        raise SpinRelaxError("libspinrelax_b200.so ABI version mismatch")
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sr_last_error().decode("utf-8", "replace")
        raise SpinRelaxError("%s failed (%d): %s" % (what or "libspinrelax_b200 call", rc, msg))


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise SpinRelaxError(
            "spinrelax_b200 needs a CUDA device (sm_100a); there is no CPU fallback for this path"
        )
    return torch


def current_stream_ptr():
    import torch

    return _c.c_void_p(torch.cuda.current_stream().cuda_stream)
