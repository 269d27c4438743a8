"""ctypes binding of libmuscade_b200.so (include/muscade_b200.h).  No fallback: if the CUDA library is missing or no
device is present, every compute entry point raises."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MB_LIB") or os.path.join(_HERE, "libmuscade_b200.so")   # MB_LIB: experiment builds only
_LIB = None

MB_OK, MB_ERR_CUDA, MB_ERR_ARG, MB_ERR_NAN, MB_ERR_STATE, MB_ERR_TOOBIG, MB_ERR_NCCL = range(7)

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class ErrInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ieletyp", C.c_int32), ("iele", C.c_int64), ("step", C.c_int64)]


class DevPtrs(C.Structure):
    _fields_ = [("X0", C.c_void_p), ("X1", C.c_void_p), ("X2", C.c_void_p), ("U0", C.c_void_p), ("Llambda", C.c_void_p),
                ("nzval", C.c_void_p), ("colptr0", C.c_void_p), ("rowval0", C.c_void_p), ("ndofX", C.c_int64), ("ndofU", C.c_int64),
                ("nnz", C.c_int64)]


# every symbol include/muscade_b200.h declares: name → (restype, argtypes)
H = C.c_void_p
SYMBOLS = {
    "mb_create": (C.c_int32, [C.c_int32, C.POINTER(H)]),
    "mb_destroy": (C.c_int32, [H]),
    "mb_last_error": (C.c_char_p, [H]),
    "mb_version": (C.c_int32, []),
    "mb_add_eulerbeam3d": (C.c_int32, [H, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "mb_add_bar3d": (C.c_int32, [H, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "mb_add_soilcontact": (C.c_int32, [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "mb_add_host_elements": (C.c_int32, [H, C.c_int64, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "mb_set_host_elements": (C.c_int32, [H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_sweepx_prepare": (C.c_int32, [H, C.c_int64, C.POINTER(C.c_int64)]),
    "mb_sweepx_get_pattern": (C.c_int32, [H, i64p, i64p]),
    "mb_sweepx_get_asm": (C.c_int32, [H, C.c_int32, C.c_void_p, C.c_void_p]),
    "mb_sweepx_get_asm_range": (C.c_int32, [H, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "mb_sweepx_assemble": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, f64p,
                                        C.c_void_p, C.c_void_p, C.POINTER(ErrInfo)]),
    "mb_sweepx_assemble_dev": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_double, f64p]),
    "mb_beam_results": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_sweepx_set_state": (C.c_int32, [H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_sweepx_get_state": (C.c_int32, [H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_sweepx_set_dof_scale": (C.c_int32, [H, C.c_void_p]),
    "mb_sweepx_newmark_decrement": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mb_sync": (C.c_int32, [H, C.POINTER(ErrInfo)]),
    "mb_get_device_ptrs": (C.c_int32, [H, C.POINTER(DevPtrs)]),
    "mb_set_ndofU": (C.c_int32, [H, C.c_int64]),
    "mb_sweepx_time_dev": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_double, f64p, C.c_int32, f32p]),
    "mb_sweepx_time_step_dev": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_double, f64p, C.c_int32, f32p]),
    "mb_measure_fp64_tflops": (C.c_int32, [H, C.POINTER(C.c_double)]),
    "mb_measure_copy_gbs": (C.c_int32, [H, C.POINTER(C.c_double)]),
    "mb_measure_host_copy_ms": (C.c_int32, [H, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_double)]),
    "mb_launch_count": (C.c_int64, [H]),
    "mb_direct_prepare": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mb_direct_class_pattern": (C.c_int32, [H, C.c_int32, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]),
    "mb_direct_get_asm": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_void_p]),
    "mb_direct_set_state": (C.c_int32, [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_direct_set_time0": (C.c_int32, [H, C.c_double]),
    "mb_direct_set_lambda_scale": (C.c_int32, [H, C.c_double]),
    "mb_direct_set_host_cost": (C.c_int32, [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_direct_set_host_elements": (C.c_int32, [H, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_direct_sparser": (C.c_int32, [H, C.c_double, C.POINTER(C.c_int64)]),
    "mb_direct_get_sparse": (C.c_int32, [H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_direct_set_lambda": (C.c_int32, [H, C.c_int64, C.c_void_p]),
    "mb_direct_get_state": (C.c_int32, [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_direct_set_dof_scale": (C.c_int32, [H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_direct_decrement": (C.c_int32, [H, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "mb_direct_assemble": (C.c_int32, [H, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(ErrInfo)]),
    "mb_direct_big_pattern": (C.c_int32, [H, C.c_void_p, C.c_void_p]),
    "mb_direct_get_step_block": (C.c_int32, [H, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "mb_direct_set_host_xx": (C.c_int32, [H, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_direct_set_gauge_cost": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_void_p, C.c_double]),
    "mb_direct_set_gauge_measurements": (C.c_int32, [H, C.c_int64, C.c_int32, C.c_void_p, C.c_int32]),
    "mb_direct_rebase": (C.c_int32, [H, C.c_int64, C.POINTER(C.c_int64)]),
    "mb_direct_step_ptrs": (C.c_int32, [H, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "mb_direct_time_dev": (C.c_int32, [H, C.c_int32, f32p]),
    "mb_xua_add_eletyp": (C.c_int32, [H, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "mb_xua_prepare": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mb_xua_class_pattern": (C.c_int32, [H, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]),
    "mb_xua_get_asm": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mb_xua_big_pattern": (C.c_int32, [H, C.c_void_p, C.c_void_p]),
    "mb_xua_big_asm": (C.c_int32, [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_xua_add_device_eletyp": (C.c_int32, [H, C.c_int32, C.POINTER(C.c_int32)]),
    "mb_xua_set_lambda_scale": (C.c_int32, [H, C.c_double]),
    "mb_xua_set_time0": (C.c_int32, [H, C.c_int32, C.c_double]),
    "mb_xua_set_gauge_cost": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_double]),
    "mb_xua_set_gauge_measurements": (C.c_int32, [H, C.c_int32, C.c_void_p, C.c_int32]),
    "mb_xua_eval_device": (C.c_int32, [H, C.c_int32, C.c_int64, C.POINTER(ErrInfo)]),
    "mb_xua_get_packet": (C.c_int32, [H, C.c_int32, C.c_void_p, C.c_void_p]),
    "mb_xua_time_device_pass": (C.c_int32, [H, C.c_int32, f32p]),
    "mb_xua_get_gauge": (C.c_int32, [H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_xua_zero": (C.c_int32, [H]),
    "mb_xua_set_packet": (C.c_int32, [H, C.c_int32, C.c_void_p, C.c_void_p]),
    "mb_xua_add_A": (C.c_int32, [H]),
    "mb_xua_add_step": (C.c_int32, [H, C.c_int32, C.c_int64]),
    "mb_xua_get_out": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mb_xua_out_shape": (C.c_int32, [H, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mb_xua_allreduce_big": (C.c_int32, [H]),
    "mb_xua_get_big": (C.c_int32, [H, C.c_void_p, C.c_void_p]),
    "mb_xua_sparser": (C.c_int32, [H, C.c_double, C.POINTER(C.c_int64)]),
    "mb_xua_get_sparse": (C.c_int32, [H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_xua_set_state": (C.c_int32, [H, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_xua_get_state": (C.c_int32, [H, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_xua_set_dof_scale": (C.c_int32, [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mb_xua_decrement": (C.c_int32, [H, C.c_void_p, C.c_void_p]),
    "mb_set_stream": (C.c_int32, [H, C.c_void_p]),
    "mb_iface_setup": (C.c_int32, [H, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "mb_iface_pack_dev": (C.c_int32, [H, C.c_void_p]),
    "mb_iface_unpack_add_dev": (C.c_int32, [H, C.c_void_p]),
    "mb_iface_exchange": (C.c_int32, [H]),
    "mb_iface_buffers": (C.c_int32, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "mb_iface_get_recvbuf": (C.c_int32, [H, C.c_void_p]),
    "mb_comm_unique_id": (C.c_int32, [C.c_void_p]),
    "mb_comm_init": (C.c_int32, [H, C.c_void_p, C.c_int32, C.c_int32]),
    "mb_comm_share": (C.c_int32, [H, H]),
    "mb_comm_destroy": (C.c_int32, [H]),
    "mb_comm_info": (C.c_int32, [H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mb_comm_allreduce": (C.c_int32, [H, C.c_void_p, C.c_int64, C.c_int32]),
    "mb_comm_allreduce_dev": (C.c_int32, [H, C.c_void_p, C.c_int64, C.c_int32]),
    "mb_comm_barrier": (C.c_int32, [H]),
    "mb_direct_halo_exchange": (C.c_int32, [H]),
    "mb_host_register": (C.c_int32, [H, C.c_void_p, C.c_int64]),
    "mb_host_unregister": (C.c_int32, [H, C.c_void_p]),
}


def build(force=False):
    """nvcc build of the CUDA library for sm_100a (cross-compiles without a GPU)."""
    if force or not os.path.exists(SO_PATH):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8", "-s"] + (["-B"] if force else []))
    return SO_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("libmuscade_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class MuscadeB200Error(RuntimeError):
    """Mirror of MuscadeException (src/Exceptions.jl:2-16): message + where (dbg) it happened."""

    def __init__(self, msg, dbg=None):
        super().__init__(msg + ("" if not dbg else "  " + repr(dbg)))
        self.dbg = dbg or {}


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check(h, rc, dbg=None):
    if rc != MB_OK:
        msg = lib().mb_last_error(h).decode() if h else "mb_create failed: no CUDA device / library (no CPU fallback)"
        raise MuscadeB200Error("muscade_b200 [%d]: %s" % (rc, msg), dbg)
