"""Thin object wrapper over the C ABI (include/muscade_b200.h): one `Engine` = one `mb_handle` = one GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ErrInfo, DevPtrs, check, ptr, MuscadeB200Error


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    def __init__(self, device=0):
        self.L = _lib.lib()
        self.h = C.c_void_p()
        rc = self.L.mb_create(int(device), C.byref(self.h))
        if rc != _lib.MB_OK:
            self.h = None
            check(None, rc)
        self.groups = []          # (kind, nele, nx)
        self.ndofX = 0
        self.nnz = 0

    def close(self):
        if self.h:
            self.L.mb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- element types -------------------------------------------------------------------------------------------------
    def add_eulerbeam3d(self, eleobj, idxX, scaleX, udof=False, idxU=None, scaleU=None):
        """eleobj (nele,69); idxX (nele,12) Int64 1-based = dis.index[iele].X; scaleX (12,)"""
        eleobj = _f64(eleobj); idxX = _i64(idxX); scaleX = _f64(scaleX)
        nele = eleobj.shape[0]
        assert eleobj.shape == (nele, 69) and idxX.shape == (nele, 12)
        idxU = _i64(idxU) if udof else None
        scaleU = _f64(scaleU) if udof else None
        ityp = C.c_int32()
        check(self.h, self.L.mb_add_eulerbeam3d(self.h, nele, ptr(eleobj), int(bool(udof)), ptr(idxX), ptr(idxU), ptr(scaleX), ptr(scaleU), C.byref(ityp)))
        self.groups.append(("eulerbeam3d", nele, 12))
        return ityp.value

    def add_bar3d(self, eleobj, idxX, scaleX, udof=False, idxU=None, scaleU=None):
        """eleobj (nele,38); idxX (nele,6)"""
        eleobj = _f64(eleobj); idxX = _i64(idxX); scaleX = _f64(scaleX)
        nele = eleobj.shape[0]
        assert eleobj.shape == (nele, 38) and idxX.shape == (nele, 6)
        idxU = _i64(idxU) if udof else None
        scaleU = _f64(scaleU) if udof else None
        ityp = C.c_int32()
        check(self.h, self.L.mb_add_bar3d(self.h, nele, ptr(eleobj), int(bool(udof)), ptr(idxX), ptr(idxU), ptr(scaleX), ptr(scaleU), C.byref(ityp)))
        self.groups.append(("bar3d", nele, 6))
        return ityp.value

    def add_soilcontact(self, eleobj, idxX, scaleX):
        """eleobj (nele,5) = z₀ Kh Kv Ch Cv; idxX (nele,3)"""
        eleobj = _f64(eleobj); idxX = _i64(idxX); scaleX = _f64(scaleX)
        nele = eleobj.shape[0]
        assert eleobj.shape == (nele, 5) and idxX.shape == (nele, 3)
        ityp = C.c_int32()
        check(self.h, self.L.mb_add_soilcontact(self.h, nele, ptr(eleobj), ptr(idxX), ptr(scaleX), C.byref(ityp)))
        self.groups.append(("soilcontact", nele, 3))
        return ityp.value

    def add_host_elements(self, idxX):
        idxX = _i64(idxX)
        nele, nx = idxX.shape
        ityp = C.c_int32()
        check(self.h, self.L.mb_add_host_elements(self.h, nele, nx, ptr(idxX), C.byref(ityp)))
        self.groups.append(("host", nele, nx))
        return ityp.value

    def set_host_elements(self, ityp, Re, Rp, Ke):
        """Re (nele,nx), Rp (nele,nx) or None, Ke (nele,nx*nx) with entry i+nx*j (column-major element matrix)."""
        check(self.h, self.L.mb_set_host_elements(self.h, ityp, ptr(_f64(Re)), ptr(_f64(Rp)), ptr(_f64(Ke))))

    def set_ndofU(self, n):
        check(self.h, self.L.mb_set_ndofU(self.h, int(n)))

    # ---- SweepX -------------------------------------------------------------------------------------------------------------
    def sweepx_prepare(self, ndofX):
        nnz = C.c_int64()
        check(self.h, self.L.mb_sweepx_prepare(self.h, int(ndofX), C.byref(nnz)))
        self.ndofX = int(ndofX); self.nnz = nnz.value
        return self.nnz

    def sweepx_pattern(self):
        colptr = np.zeros(self.ndofX + 1, np.int64); rowval = np.zeros(self.nnz, np.int64)
        check(self.h, self.L.mb_sweepx_get_pattern(self.h, colptr, rowval))
        return colptr, rowval

    def sweepx_asm(self, ityp, want2=True):
        _, nele, nx = self.groups[ityp - 1]
        asm1 = np.zeros((nele, nx), np.int64)
        asm2 = np.zeros((nele, nx * nx), np.int64) if want2 else None
        check(self.h, self.L.mb_sweepx_get_asm(self.h, ityp, ptr(asm1), ptr(asm2)))
        return asm1, asm2

    def sweepx_asm_range(self, ityp, e0, e1):
        _, nele, nx = self.groups[ityp - 1]
        asm1 = np.zeros((e1 - e0, nx), np.int64); asm2 = np.zeros((e1 - e0, nx * nx), np.int64)
        check(self.h, self.L.mb_sweepx_get_asm_range(self.h, ityp, int(e0), int(e1), ptr(asm1), ptr(asm2)))
        return asm1, asm2

    def sweepx_assemble(self, OX, mission, X, newmark, U0=None, t=0., Llambda=None, nzval=None, dbg=None, nzval_on_device=False):
        """assemble!{mission}: host state in, host Lλ / nzval out (pass preallocated arrays to avoid allocation).
        nzval_on_device: the CSC values stay in HBM behind mb_get_device_ptrs (device solver hand-off); only Lλ comes back."""
        X = [_f64(x) for x in X]
        if Llambda is None:
            Llambda = np.empty(self.ndofX)
        if nzval is None and not nzval_on_device:
            nzval = np.empty(self.nnz)
        where = ErrInfo()
        rc = self.L.mb_sweepx_assemble(self.h, OX, {"step": 0, "iter": 1}[mission], ptr(X[0]), ptr(X[1]) if OX >= 1 else None,
                                       ptr(X[2]) if OX >= 2 else None, ptr(_f64(U0)), float(t), _f64(newmark), ptr(Llambda), ptr(nzval),
                                       C.byref(where))
        if rc == _lib.MB_ERR_NAN:
            d = dict(dbg or {}); d.update(ieletyp=where.ieletyp, iele=where.iele)
            raise MuscadeB200Error("residual(...) returned NaN in R, FB or derivatives", d)
        check(self.h, rc, dbg)
        return Llambda, nzval

    # ---- device-resident state
    def set_state(self, X, U0=None):
        """state.X[1..OX+1] (and state.U[1]) → device"""
        X = [_f64(x) for x in X]
        OX = len(X) - 1
        check(self.h, self.L.mb_sweepx_set_state(self.h, OX, ptr(X[0]), ptr(X[1]) if OX >= 1 else None, ptr(X[2]) if OX >= 2 else None, ptr(_f64(U0))))

    def get_state(self, OX):
        X = [np.empty(self.ndofX) for _ in range(OX + 1)]
        check(self.h, self.L.mb_sweepx_get_state(self.h, OX, ptr(X[0]), ptr(X[1]) if OX >= 1 else None, ptr(X[2]) if OX >= 2 else None))
        return X

    def set_dof_scale(self, scaleX):
        s = _f64(scaleX)
        assert s.shape == (self.ndofX,)
        check(self.h, self.L.mb_sweepx_set_dof_scale(self.h, ptr(s)))

    def newmark_decrement(self, OX, firstiter, dx, newmark, norms=True):
        """Newmarkβdecrement!{OX} on the device-resident state (SweepX.jl:98-132); returns (Σ Δx², Σ Lλ²) when norms."""
        dx = _f64(dx)
        assert dx.shape == (self.ndofX,)
        a, b = C.c_double(0.), C.c_double(0.)
        check(self.h, self.L.mb_sweepx_newmark_decrement(self.h, OX, int(bool(firstiter)), ptr(dx), ptr(_f64(newmark)),
                                                         C.byref(a) if norms else None, C.byref(b) if norms else None))
        return a.value, b.value

    def sweepx_assemble_resident(self, OX, mission, newmark, U0=None, t=0., Llambda=None, nzval=None, dbg=None):
        """assemble!{mission} at the device-resident state; host Lλ / nzval out (chunk-pipelined copy for large models)."""
        if Llambda is None:
            Llambda = np.empty(self.ndofX)
        if nzval is None:
            nzval = np.empty(self.nnz)
        where = ErrInfo()
        rc = self.L.mb_sweepx_assemble(self.h, OX, {"step": 0, "iter": 1}[mission], None, None, None, ptr(_f64(U0)), float(t), _f64(newmark),
                                       ptr(Llambda), ptr(nzval), C.byref(where))
        if rc == _lib.MB_ERR_NAN:
            d = dict(dbg or {}); d.update(ieletyp=where.ieletyp, iele=where.iele)
            raise MuscadeB200Error("residual(...) returned NaN in R, FB or derivatives", d)
        check(self.h, rc, dbg)
        return Llambda, nzval

    def beam_results(self, ityp, OX, X=None):
        """getresult for all elements of EulerBeam3D type `ityp` (Output.jl:131-181): dict of arrays eps (n), rsm (n,3,3), kappa (n,3) and per Gauss
        point x, kappa_gp, mi, fe, me (n,4,3), fi (n,4)  [ε, rₛₘ, ♢κ, x, κgp, mᵢ, fₑ, mₑ, fᵢ of the reference].  X=None: device-resident state."""
        nele = self.groups[ityp - 1][1]
        raw = np.empty((nele, 77))
        X = None if X is None else [_f64(x) for x in X]
        check(self.h, self.L.mb_beam_results(self.h, int(ityp), int(OX), ptr(X[0]) if X else None, ptr(X[1]) if X and OX >= 1 else None,
                                             ptr(X[2]) if X and OX >= 2 else None, ptr(raw)))
        gp = raw[:, 13:].reshape(nele, 4, 16)
        return {"raw": raw, "eps": raw[:, 0], "rsm": raw[:, 1:10].reshape(nele, 3, 3).transpose(0, 2, 1), "kappa": raw[:, 10:13], "x": gp[:, :, 0:3],
                "kappa_gp": gp[:, :, 3:6], "fi": gp[:, :, 6], "mi": gp[:, :, 7:10], "fe": gp[:, :, 10:13], "me": gp[:, :, 13:16]}

    def sweepx_assemble_dev(self, OX, mission, newmark, t=0.):
        check(self.h, self.L.mb_sweepx_assemble_dev(self.h, OX, {"step": 0, "iter": 1}[mission], float(t), _f64(newmark)))

    def sync(self):
        where = ErrInfo()
        rc = self.L.mb_sync(self.h, C.byref(where))
        if rc == _lib.MB_ERR_NAN:
            raise MuscadeB200Error("residual(...) returned NaN in R, FB or derivatives", dict(ieletyp=where.ieletyp, iele=where.iele))
        check(self.h, rc)

    def device_ptrs(self):
        p = DevPtrs()
        check(self.h, self.L.mb_get_device_ptrs(self.h, C.byref(p)))
        return p

    def time_dev(self, OX, mission, newmark, reps=3):
        ms = np.zeros(2, np.float32)
        check(self.h, self.L.mb_sweepx_time_dev(self.h, OX, {"step": 0, "iter": 1}[mission], 0., _f64(newmark), reps, ms))
        return float(ms[0]), float(ms[1])

    def time_step_dev(self, OX, mission, newmark, reps=3):
        """average ms of `reps` whole device-resident steps (sweepx_assemble_dev as configured)"""
        ms = np.zeros(1, np.float32)
        check(self.h, self.L.mb_sweepx_time_step_dev(self.h, OX, {"step": 0, "iter": 1}[mission], 0., _f64(newmark), reps, ms))
        return float(ms[0])

    def fp64_tflops(self):
        v = C.c_double()
        check(self.h, self.L.mb_measure_fp64_tflops(self.h, C.byref(v)))
        return v.value

    def copy_gbs(self):
        v = C.c_double()
        check(self.h, self.L.mb_measure_copy_gbs(self.h, C.byref(v)))
        return v.value

    def host_copy_ms(self, host_in, host_out, reps=3):
        """ms per round of len(host_in) bytes H2D and as many D2H at once (pinned buffers): the ceiling of the host-buffer e2e path"""
        v = C.c_double()
        check(self.h, self.L.mb_measure_host_copy_ms(self.h, ptr(host_in), ptr(host_out), int(host_in.nbytes), int(reps), C.byref(v)))
        return v.value

    def set_stream(self, cuda_stream):
        check(self.h, self.L.mb_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def iface_setup(self, send_nz, send_v, recv_nz, recv_v):
        a = [_i64(x) for x in (send_nz, send_v, recv_nz, recv_v)]
        check(self.h, self.L.mb_iface_setup(self.h, len(a[0]), ptr(a[0]), len(a[1]), ptr(a[1]), len(a[2]), ptr(a[2]), len(a[3]), ptr(a[3])))

    def iface_pack(self, dev_ptr):
        check(self.h, self.L.mb_iface_pack_dev(self.h, C.c_void_p(int(dev_ptr))))

    def iface_unpack_add(self, dev_ptr):
        check(self.h, self.L.mb_iface_unpack_add_dev(self.h, C.c_void_p(int(dev_ptr))))

    def iface_exchange(self):
        """pack → NCCL send to rank+1 / receive from rank−1 → unpack-add, inside the shim, asynchronous on the handle's stream"""
        check(self.h, self.L.mb_iface_exchange(self.h))

    def iface_buffers(self):
        """(send pointer, n_send, receive pointer, n_recv) of the device buffers mb_iface_exchange uses"""
        ps, pr, ns, nr = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        check(self.h, self.L.mb_iface_buffers(self.h, C.byref(ps), C.byref(ns), C.byref(pr), C.byref(nr)))
        return ps.value, ns.value, pr.value, nr.value

    def iface_recvbuf(self):
        n = self.iface_buffers()[3]
        out = np.zeros(n)
        check(self.h, self.L.mb_iface_get_recvbuf(self.h, ptr(out)))
        return out

    # ---- NCCL inside the shim ------------------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (one process creates it, the host hands it to the others)"""
        buf = np.zeros(128, np.uint8)
        rc = _lib.lib().mb_comm_unique_id(ptr(buf))
        if rc != _lib.MB_OK:
            raise MuscadeB200Error("muscade_b200 [%d]: ncclGetUniqueId failed (libnccl.so.2 missing?)" % rc)
        return buf

    def comm_init(self, uid, rank, world):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        assert uid.size == 128
        check(self.h, self.L.mb_comm_init(self.h, ptr(uid), int(rank), int(world)))

    def comm_init_from(self, owner):
        """borrow the NCCL communicator of another Engine on the same GPU (which must outlive this one)"""
        check(self.h, self.L.mb_comm_share(self.h, owner.h))

    def comm_info(self):
        r, w, v = C.c_int32(), C.c_int32(), C.c_int32()
        check(self.h, self.L.mb_comm_info(self.h, C.byref(r), C.byref(w), C.byref(v)))
        return r.value, w.value, v.value

    def comm_allreduce(self, values, op="sum"):
        a = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64).copy()
        check(self.h, self.L.mb_comm_allreduce(self.h, ptr(a), a.size, {"sum": 0, "max": 1, "min": 2}[op]))
        return a

    def comm_barrier(self):
        check(self.h, self.L.mb_comm_barrier(self.h))

    def pin(self, a):
        """page-lock a numpy array in place (cudaHostRegister)"""
        check(self.h, self.L.mb_host_register(self.h, ptr(a), a.nbytes))

    def unpin(self, a):
        check(self.h, self.L.mb_host_unregister(self.h, ptr(a)))

    def launch_count(self):
        return int(self.L.mb_launch_count(self.h))
