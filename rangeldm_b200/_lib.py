"""ctypes binding of librldm.so (the C ABI declared in include/rldm.h).

The library is built IN-TREE by `__graft_entry__.build()` / `rangeldm_b200.build.build_library()`
(`nvcc -gencode arch=compute_100a,code=sm_100a`).  There is no CPU fallback: every compute entry
point raises if the library is missing or a call fails.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RLDM_LIB") or os.path.join(_HERE, "librldm.so")   # RLDM_LIB: experiment builds

c_int, c_float, c_void_p, c_i64 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64

# kind codes of rldm_op (include/rldm.h)
OP_GN_STATS, OP_PREP, OP_CONV_TC, OP_CONV_IN, OP_CONV_OUT, OP_ATTENTION, OP_TEMB, OP_SCHED_STEP, \
    OP_MEMSET, OP_CONV_REF, OP_AXPY, OP_NORM_CONV_OUT, OP_FUSED, OP_CONV_UP2 = range(1, 15)


class RldmOp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("i", ctypes.c_int32 * 23), ("f", ctypes.c_float * 2),
                ("p", ctypes.c_void_p * 24), ("n", ctypes.c_int64)]


class ConvEmit(ctypes.Structure):
    """rldm_conv_emit (include/rldm.h)"""
    _fields_ = [("out", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("eps", ctypes.c_float),
                ("G", ctypes.c_int), ("silu", ctypes.c_int), ("circular", ctypes.c_int)]


class ConvSrc(ctypes.Structure):
    """rldm_conv_src (include/rldm.h)"""
    _fields_ = [("x0", ctypes.c_void_p), ("x1", ctypes.c_void_p), ("pairs0", ctypes.c_void_p), ("pairs1", ctypes.c_void_p),
                ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("eps", ctypes.c_float), ("c0", ctypes.c_int),
                ("c1", ctypes.c_int), ("G", ctypes.c_int), ("silu", ctypes.c_int), ("up", ctypes.c_int),
                ("circular", ctypes.c_int)]


# name -> (restype, argtypes); must list every symbol include/rldm.h declares
SIGNATURES = {
    "rldm_version": (c_int, []),
    "rldm_last_error": (ctypes.c_char_p, []),
    "rldm_reload_env": (None, []),
    "rldm_gn_stats": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rldm_prep": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int,
                          c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rldm_conv_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
                     + [c_int] * 10 + [c_void_p, c_void_p]),
    "rldm_conv_tc_shortcut": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
                              + [c_int] * 10 + [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "rldm_conv_tc_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
                        + [c_int] * 10 + [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "rldm_conv_tc_fused": (c_int, [c_void_p, c_void_p] + [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
                           + [c_int] * 10 + [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "rldm_conv_tc_fusable": (c_int, [c_int] * 10),
    "rldm_conv_tc_up2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_void_p, c_int, c_void_p]),
    "rldm_conv_tc_up2_ok": (c_int, [c_int] * 5),
    "rldm_conv_tc_emit": (c_int, [c_void_p] + [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
                          + [c_int] * 10 + [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "rldm_conv_tc_emittable": (c_int, [c_int] * 11),
    "rldm_conv_ref": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
                      + [c_int] * 9 + [c_void_p]),
    "rldm_conv_in": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p] + [c_int] * 5
                     + [c_void_p]),
    "rldm_norm_conv_out": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    "rldm_conv_in_stats": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p] + [c_int] * 5
                           + [c_void_p, c_void_p]),
    "rldm_conv_out": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    "rldm_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rldm_temb": (c_int, [c_void_p] * 9 + [c_int] * 4 + [c_void_p]),
    "rldm_sched_step": (c_int, [c_void_p] * 7 + [c_i64, c_void_p]),
    "rldm_range_to_points": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_float,
                                     c_float, c_void_p, c_void_p, c_void_p]),
    "rldm_points_to_voxel": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p]),
    "rldm_points_to_range": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float,
                                     c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rldm_scale": (c_int, [c_void_p, c_float, c_void_p, c_i64, c_void_p]),
    "rldm_ref_to_cl": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rldm_cl_to_ref": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rldm_run": (c_int, [ctypes.POINTER(RldmOp), c_int, c_void_p]),
    "rldm_fused_supported": (c_int, [ctypes.POINTER(RldmOp)]),
    "rldm_fused_ws_bytes": (ctypes.c_longlong, [ctypes.POINTER(RldmOp), c_int]),
    "rldm_fused_create": (c_int, [ctypes.POINTER(RldmOp), c_int, c_void_p, ctypes.c_longlong, ctypes.POINTER(c_void_p)]),
    "rldm_fused_run": (c_int, [c_void_p, c_void_p]),
    "rldm_fused_destroy": (None, [c_void_p]),
    "rldm_run_timed": (c_int, [ctypes.POINTER(RldmOp), c_int, c_void_p, c_void_p]),
}

_lib = None


class RldmError(RuntimeError):
    pass


def lib():
    """Load librldm.so (once).  Raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RldmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(rangeldm_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RldmError(f"librldm call failed (rc={rc}): {lib().rldm_last_error().decode(errors='replace')}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    check(getattr(lib(), name)(*args, stream_ptr()))
