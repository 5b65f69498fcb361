"""ctypes binding of include/rustpde_b200.h.

`product_lib()` loads librustpde_b200.so (CUDA, sm_100a) that sits next to
this file and nothing else: if the library is missing, is not the CUDA build,
or no GPU is usable, it raises -- there is no CPU fallback in this package.
"""
import ctypes as C
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(HERE, "librustpde_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)

# name -> argtypes (restype is int unless listed in _RESTYPE)
_SIGS = {
    "rp_init": [C.c_int],
    "rp_last_error": [],
    "rp_version": [],
    "rp_is_emulated": [],
    "rp_set_lapack_library": [C.c_char_p],
    "rp_field_create": [C.c_int, C.c_int, C.c_int, C.c_int, vpp],
    "rp_field_destroy": [vp],
    "rp_field_shape": [vp, c_int_p, c_int_p, c_int_p, c_int_p],
    "rp_field_coords": [vp, C.c_int, c_double_p, C.c_size_t],
    "rp_field_dx": [vp, C.c_int, c_double_p, C.c_size_t],
    "rp_field_upload_v": [vp, c_double_p, C.c_size_t],
    "rp_field_download_v": [vp, c_double_p, C.c_size_t],
    "rp_field_upload_vhat": [vp, c_double_p, C.c_size_t],
    "rp_field_download_vhat": [vp, c_double_p, C.c_size_t],
    "rp_field_upload_vhat_rows": [vp, C.c_int, C.c_int, c_double_p, C.c_size_t],
    "rp_field_download_vhat_rows": [vp, C.c_int, C.c_int, c_double_p, C.c_size_t],
    "rp_field_forward": [vp],
    "rp_field_backward": [vp],
    "rp_field_to_ortho": [vp, c_double_p, C.c_size_t],
    "rp_field_from_ortho": [vp, c_double_p, C.c_size_t],
    "rp_field_gradient": [vp, C.c_int, C.c_int, c_double_p, c_double_p, C.c_size_t],
    "rp_field_average": [vp, c_double_p],
    "rp_field_average_axis": [vp, C.c_int, c_double_p, C.c_size_t],
    "rp_hholtz_create": [vp, C.c_double, C.c_double, C.c_double, vpp],
    "rp_hholtz_adi_create": [vp, C.c_double, C.c_double, vpp],
    "rp_poisson_create": [vp, C.c_double, C.c_double, vpp],
    "rp_hholtz_create_with_eig": [vp, C.c_double, C.c_double, C.c_double, c_double_p, c_double_p, c_double_p, vpp],
    "rp_poisson_create_with_eig": [vp, C.c_double, C.c_double, c_double_p, c_double_p, c_double_p, vpp],
    "rp_solver_eig_size": [vp, c_int_p, c_int_p],
    "rp_solver_export_eig": [vp, c_double_p, c_double_p, c_double_p],
    "rp_solver_solve": [vp, c_double_p, C.c_size_t, c_double_p, C.c_size_t, C.c_int],
    "rp_solver_solve_resident": [vp, C.c_int, C.c_int],
    "rp_solver_sync": [vp],
    "rp_solver_path": [vp, c_int_p, c_int_p, c_int_p],
    "rp_solver_destroy": [vp],
    "rp_navier_create": [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, vpp],
    "rp_navier_create_with_eig": [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, c_double_p, c_double_p, c_double_p, vpp],
    "rp_navier_destroy": [vp],
    "rp_navier_set_velocity": [vp, C.c_double, C.c_double, C.c_double],
    "rp_navier_set_temperature": [vp, C.c_double, C.c_double, C.c_double],
    "rp_navier_set_tempbc_ortho": [vp, c_double_p, C.c_size_t],
    "rp_navier_set_dealias": [vp, C.c_int],
    "rp_navier_set_solid": [vp, c_double_p, c_double_p, C.c_size_t],
    "rp_navier_update": [vp, C.c_int],
    "rp_navier_stage_state": [vp, c_double_p, C.c_size_t, c_double_p, C.c_size_t, c_double_p, C.c_size_t, c_double_p, C.c_size_t],
    "rp_navier_commit_staged": [vp],
    "rp_navier_fetch_state": [vp, c_double_p, C.c_size_t, c_double_p, C.c_size_t, c_double_p, C.c_size_t, c_double_p, C.c_size_t],
    "rp_navier_fetch_wait": [vp],
    "rp_navier_div_async": [vp],
    "rp_navier_div_poll": [vp, C.c_int, c_double_p, c_int_p],
    "rp_navier_write_snapshot": [vp, C.c_char_p],
    "rp_navier_read_snapshot": [vp, C.c_char_p],
    "rp_navier_sync": [vp],
    "rp_navier_get_time": [vp, c_double_p],
    "rp_navier_get_dt": [vp, c_double_p],
    "rp_navier_reset_time": [vp],
    "rp_navier_params": [vp, c_double_p, c_double_p, c_double_p],
    "rp_navier_eval": [vp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p],
    "rp_navier_field": [vp, C.c_int, vpp],
    "rp_navier_export_eig": [vp, c_double_p, c_double_p, c_double_p],
    "rp_navier_launches_per_step": [vp, c_int_p],
    "rp_navier_set_graph": [vp, C.c_int],
    "rp_navier_kernel_path": [vp, c_int_p, c_int_p],
    "rp_navier_slab_phase1": [vp, C.c_int, C.c_int, C.POINTER(C.c_void_p)],
    "rp_navier_slab_phase2": [vp, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_void_p)],
    "rp_navier_slab_phase3": [vp, C.c_int, C.c_int, C.POINTER(C.c_void_p)],
    "rp_navier_slab_phase1_p2p": [vp, C.c_int, C.c_int, C.c_int, c_int_p, C.POINTER(C.c_void_p)],
    "rp_navier_slab_phase2_p2p": [vp, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_int, c_int_p, C.POINTER(C.c_void_p)],
    "rp_dev_alloc": [C.c_size_t, vpp],
    "rp_dev_free": [vp],
    "rp_ipc_export": [vp, C.c_char_p],
    "rp_ipc_open": [C.c_char_p, vpp],
    "rp_ipc_close": [vp],
    "rp_adjoint_create": [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, vpp],
    "rp_adjoint_destroy": [vp],
    "rp_adjoint_set_velocity": [vp, C.c_double, C.c_double, C.c_double],
    "rp_adjoint_set_temperature": [vp, C.c_double, C.c_double, C.c_double],
    "rp_adjoint_update": [vp, C.c_int],
    "rp_adjoint_get_time": [vp, c_double_p],
    "rp_adjoint_reset_time": [vp],
    "rp_adjoint_eval": [vp, c_double_p, c_double_p, c_double_p, c_double_p],
    "rp_adjoint_residuals": [vp, c_double_p, c_double_p],
    "rp_adjoint_exit": [vp, c_int_p],
    "rp_adjoint_field": [vp, C.c_int, vpp],
    "rp_adjoint_solver": [vp, C.c_int, vpp],
    "rp_navier_profile": [vp, C.c_int, c_double_p, C.c_size_t, c_int_p],
    "rp_navier_op_info": [vp, C.c_int, C.c_char_p, C.c_size_t, c_double_p, c_double_p],
}
_RESTYPE = {"rp_last_error": C.c_char_p}

EXPORTED_SYMBOLS = sorted(_SIGS)


class RustpdeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rustpde_b200 error %d: %s" % (code, msg))
        self.code = code


def find_lapack():
    """A LAPACK provider for solver set-up (dgeev/dgetri): scipy's bundled OpenBLAS."""
    env = os.environ.get("RUSTPDE_B200_LAPACK")
    if env:
        return env
    for p in sys.path:
        for pat in ("scipy.libs/libscipy_openblas-*.so", "scipy.libs/libscipy_openblas*.so", "numpy.libs/libscipy_openblas*.so"):
            hits = sorted(glob.glob(os.path.join(p, pat)))
            if hits:
                return hits[0]
    return None


class Lib:
    """A loaded librustpde_b200 with typed entry points; `call` raises on a non-zero status."""

    def __init__(self, path, device=0):
        if not os.path.exists(path):
            raise RuntimeError("rustpde_b200: shared library not found: %s (run `python -m rustpde_b200.build`)" % path)
        self.path = path
        self.c = C.CDLL(path)
        for name, args in _SIGS.items():
            fn = getattr(self.c, name)  # AttributeError if a declared symbol is not exported
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        self.emulated = bool(self.c.rp_is_emulated())
        lp = find_lapack()
        if lp:
            self.c.rp_set_lapack_library(lp.encode())
        self.call("rp_init", device)

    def call(self, name, *args):
        rc = getattr(self.c, name)(*args)
        if rc != 0:
            msg = self.c.rp_last_error()
            raise RustpdeError(rc, msg.decode() if msg else "")
        return rc


_product = None


def product_lib(device=None):
    """The CUDA library, or an exception.  Never returns an emulated/CPU build."""
    global _product
    if _product is None:
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        lib = Lib(PRODUCT_LIB, device)
        if lib.emulated:
            raise RuntimeError("rustpde_b200: %s is not the CUDA build" % PRODUCT_LIB)
        _product = lib
    return _product
