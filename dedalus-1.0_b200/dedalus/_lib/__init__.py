"""ctypes binding of libddl_b200.so (C ABI: include/ddl.h).

There is no fallback: if the shared library has not been built
(``python dedalus-1.0_b200/build.py`` / ``__graft_entry__.build()``) importing this module
raises, and every compute entry point of the package depends on it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libddl_b200.so")
# The one other library this module will bind is the g++ host-emulation build of the same sources, and only inside the test
# harness (tests/conftest.py sets both variables in the test process): a stray DEDALUS_DDL_LIB in production is an error.
if os.environ.get("DEDALUS_DDL_LIB"):
    if os.environ.get("DDL_TEST_HOST_EMUL") != "1":
        raise ImportError("DEDALUS_DDL_LIB is honoured only by the test harness (DDL_TEST_HOST_EMUL=1); unset it")
    LIB_PATH = os.environ["DEDALUS_DDL_LIB"]

EXPORTS = [
    "ddl_plan_create", "ddl_plan_create_slab", "ddl_plan_destroy", "ddl_workspace_bytes", "ddl_rhs_workspace_bytes",
    "ddl_forward", "ddl_backward", "ddl_dealias", "ddl_dealias_array", "ddl_copy_boxes", "ddl_deriv", "ddl_rhs", "ddl_stage",
    "ddl_rk4_stage", "ddl_cn_step", "ddl_step_array", "ddl_stage_outside", "ddl_rhs_stage", "ddl_slab_assemble_stage", "ddl_slab_info", "ddl_slab_rows", "ddl_slab_theta", "ddl_slab_zinv", "ddl_slab_yinv",
    "ddl_slab_xfused", "ddl_slab_xc2r", "ddl_slab_xr2c", "ddl_slab_yfwd", "ddl_slab_zfwd", "ddl_slab_assemble",
    "ddl_p2p_create", "ddl_p2p_connect", "ddl_p2p_base", "ddl_p2p_exchange", "ddl_p2p_wait", "ddl_p2p_destroy",
    "ddl_p2p_peer_base", "ddl_p2p_signal", "ddl_p2p_push", "ddl_slab_yfwd_planes", "ddl_slab_zinv_peer", "ddl_slab_yfwd_peer", "ddl_slab_xfused_planes", "ddl_launch_count",
    "ddl_reduce_invariants", "ddl_reduce_outside_mask", "ddl_reduce_max_square", "ddl_rhs_capture_max", "ddl_set_shear", "ddl_profile_enable", "ddl_profile_report", "ddl_measure_fp64", "ddl_set_option", "ddl_sync", "ddl_last_error", "ddl_version",
]

HYDRO, BOUSSINESQ, MHD = 0, 1, 2
ADV = 3     # DDL_*_ADV = base id + 3: advective-form policies for states that are not solenoidal
EULER, ETD1, ETD2RK1, ETD2RK2 = 0, 1, 2, 3
RHS_ZERO_FILL, RHS_DEALIAS_STATE = 1, 2
STAGE_RETAINED_ONLY = 1
# include/ddl.h DDL_INV_*: entries of the vector ddl_reduce_invariants fills
INV = dict(ekin=0, e2=1, div_sum=2, mag_div_sum=3, enstrophy=4, current2=5, hel_kin=6, hel_cross=7, div_re=8, div_im=9,
           mag_div_re=10, mag_div_im=11, cenk_num=12, cenk_den=13, msq=14, hel_mag=20, grad2_T=21, div2=22, mag_div2=23)
NINV = 24


class PhysParams(C.Structure):
    _fields_ = [("rho0", C.c_double), ("g", C.c_double), ("alpha_t", C.c_double), ("beta", C.c_double),
                ("boussinesq_dir", C.c_int), ("reserved", C.c_int)]


FUSE_RK4, FUSE_CN = 4, 5


class StageFuse(C.Structure):
    """include/ddl.h: ddl_stage_fuse."""
    _fields_ = [("y", C.c_void_p), ("total", C.c_void_p), ("out", C.c_void_p), ("coeff", C.c_void_p),
                ("visc_order", C.c_int), ("first", C.c_int), ("last", C.c_int), ("wdiv", C.c_double), ("dt_step", C.c_double),
                ("kind", C.c_int), ("reserved", C.c_int), ("deriv1", C.c_void_p), ("k_out", C.c_void_p)]


class DDLError(RuntimeError):
    pass


def bind_slab(lib):
    """argtypes of the slab phase API (also applied to the host-emulation library in tests)."""
    vp, i32 = C.c_void_p, C.c_int
    lib.ddl_slab_info.argtypes = [vp, vp]
    lib.ddl_slab_rows.argtypes = [vp, vp]
    lib.ddl_slab_zinv.argtypes = [vp, i32, vp, vp, vp]
    lib.ddl_slab_theta.argtypes = [vp, i32, vp, vp]
    lib.ddl_slab_yinv.argtypes = [vp, i32, vp, vp, vp]
    lib.ddl_slab_xfused.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.ddl_slab_xc2r.argtypes = [vp, vp, vp, vp]
    lib.ddl_slab_xr2c.argtypes = [vp, vp, vp, vp]
    lib.ddl_slab_yfwd.argtypes = [vp, i32, vp, vp, vp]
    lib.ddl_slab_zfwd.argtypes = [vp, i32, vp, vp, i32, vp]
    lib.ddl_slab_assemble.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.ddl_dealias.argtypes = [vp, vp, vp]
    lib.ddl_slab_assemble_stage.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.ddl_rhs_stage.argtypes = [vp, i32, vp, vp, vp, C.c_size_t, i32, vp, vp]
    lib.ddl_reduce_invariants.argtypes = [vp, i32, vp, i32, vp, vp]
    lib.ddl_reduce_max_square.argtypes = [vp, i32, vp, vp, vp, C.c_size_t, i32, vp, vp]
    lib.ddl_rhs_capture_max.argtypes = [vp, vp]
    lib.ddl_stage_outside.argtypes = [vp, i32, i32, vp, vp, vp, i32, C.c_double, vp]
    lib.ddl_reduce_outside_mask.argtypes = [vp, i32, vp, vp, vp]
    lib.ddl_dealias_array.argtypes = [i32, vp, vp, vp, vp, vp, i32, vp, vp]
    lib.ddl_copy_boxes.argtypes = [vp, vp, vp, i32, vp, i32, i32, vp]
    lib.ddl_step_array.argtypes = [i32, i32, C.c_longlong, vp, vp, vp, vp, vp, C.c_double, vp]
    lib.ddl_set_shear.argtypes = [vp, i32, C.c_double, C.c_double, C.c_double]
    if hasattr(lib, "ddl_p2p_create"):
        lib.ddl_p2p_create.argtypes = [C.POINTER(vp), i32, i32, C.c_size_t, C.c_char_p]
        lib.ddl_p2p_connect.argtypes = [vp, C.c_char_p]
        lib.ddl_p2p_base.argtypes = [vp]
        lib.ddl_p2p_base.restype = vp
        lib.ddl_p2p_exchange.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        lib.ddl_p2p_exchange.restype = C.c_longlong
        lib.ddl_p2p_wait.argtypes = [vp, C.c_longlong, vp]
        lib.ddl_p2p_destroy.argtypes = [vp]
        lib.ddl_p2p_peer_base.argtypes = [vp, i32]
        lib.ddl_p2p_peer_base.restype = vp
        lib.ddl_p2p_signal.argtypes = [vp, vp]
        lib.ddl_p2p_signal.restype = C.c_longlong
        lib.ddl_p2p_push.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp]
        lib.ddl_p2p_push.restype = C.c_longlong
        lib.ddl_slab_yfwd_planes.argtypes = [vp, i32, vp, vp, i32, i32, vp]
        lib.ddl_slab_zinv_peer.argtypes = [vp, i32, vp, vp, vp]
        lib.ddl_slab_yfwd_peer.argtypes = [vp, i32, vp, vp, i32, i32, vp]
        lib.ddl_slab_xfused_planes.argtypes = [vp, i32, vp, vp, vp, i32, i32, vp]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libddl_b200.so is missing (%s): build it with `python dedalus-1.0_b200/build.py`. "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
    lib.ddl_plan_create.argtypes = [C.POINTER(vp), i32, vp, vp, vp, vp, vp, vp, vp]
    lib.ddl_plan_create_slab.argtypes = [C.POINTER(vp), i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32]
    lib.ddl_plan_destroy.argtypes = [vp]
    bind_slab(lib)
    lib.ddl_workspace_bytes.argtypes = [vp, i32, i32]
    lib.ddl_workspace_bytes.restype = sz
    lib.ddl_rhs_workspace_bytes.argtypes = [vp, i32]
    lib.ddl_rhs_workspace_bytes.restype = sz
    lib.ddl_forward.argtypes = [vp, vp, vp, vp, sz, vp]
    lib.ddl_backward.argtypes = [vp, vp, vp, vp, sz, vp]
    lib.ddl_dealias.argtypes = [vp, vp, vp]
    lib.ddl_deriv.argtypes = [vp, vp, vp, i32, vp]
    lib.ddl_rhs.argtypes = [vp, i32, C.POINTER(PhysParams), vp, vp, vp, sz, i32, vp]
    lib.ddl_stage.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, dbl, i32, vp]
    lib.ddl_rk4_stage.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, dbl, dbl, i32, i32, i32, vp]
    lib.ddl_cn_step.argtypes = [vp, i32, vp, vp, vp, i32, dbl, i32, vp]
    lib.ddl_launch_count.restype = C.c_longlong
    lib.ddl_profile_enable.argtypes = [i32]
    lib.ddl_profile_report.argtypes = [C.c_char_p, sz]
    lib.ddl_set_option.argtypes = [C.c_char_p, i32]
    lib.ddl_measure_fp64.argtypes = [vp, vp]
    lib.ddl_sync.argtypes = [vp]
    lib.ddl_last_error.restype = C.c_char_p
    lib.ddl_version.restype = C.c_char_p
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise DDLError(lib.ddl_last_error().decode())


def ptr_array(tensors):
    """(void*)[n] of device pointers; keep the tensors alive while it is in use."""
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def launch_count():
    return int(lib.ddl_launch_count())


def profile(enable):
    check(lib.ddl_profile_enable(1 if enable else 0))


def profile_report():
    """{label: {"n": launches, "ms": device ms}} for the launches recorded since profile(True)."""
    import json
    buf = C.create_string_buffer(1 << 16)
    check(lib.ddl_profile_report(buf, len(buf)))
    return json.loads(buf.value.decode())


def measure_fp64(stream=None):
    """(DFMA TFLOP/s, DADD TFLOP/s) of the current device, measured now (include/ddl.h ddl_measure_fp64)."""
    out = (C.c_double * 2)()
    check(lib.ddl_measure_fp64(out, stream))
    return float(out[0]), float(out[1])


def set_option(name, value):
    check(lib.ddl_set_option(name.encode(), int(value)))
