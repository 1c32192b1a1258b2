"""ctypes binding of libprt.so (C ABI in include/prt.h).

There is no CPU implementation behind this module: if the CUDA library is missing, or no B200 is
visible, every compute entry point raises.  PyTorch supplies device memory, streams and (for
multi-GPU) torch.distributed; the arithmetic is all in libprt.so.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PRT_LIB_PATH") or os.path.join(_HERE, "libprt.so")  # override: build variants

PRT_SUCCESS = 0
RAY_STATUS = {
    0: "OK", 1: "STEP_TOO_SMALL", 2: "VERTICAL", 3: "BBOX", 4: "BACKWARDS", 5: "BOTTOM_OOB",
    6: "BRENTQ_SIGN", 7: "BRENTQ_ITER", 9: "MAX_BOUNCES",
}
EIG_CONVERGED, EIG_RAY_DROPPED, EIG_MAX_ITER = 0, 1, 2
COUNTER_NAMES = ("nfev", "attempts", "accepted", "segments", "bisect", "dense")
N_COUNTERS = 6
EIG_TABLE_WIDTH = 10
EIG_TABLE_FIELDS = ("rcv", "T", "z", "p", "launch_angle", "received_angle", "ray_id", "n_bottom", "n_surface", "bracket")
DEFAULT_MAX_BOUNCES = 1 << 20

EXPORTS = (
    "prt_abi_version", "prt_set_concurrency", "prt_set_lanes", "prt_last_error", "prt_device_count", "prt_env_create", "prt_ensemble_create",
    "prt_ensemble_create_dev",
    "prt_shoot_ensemble", "prt_env_destroy",
    "prt_env_device_bytes", "prt_shoot_fan", "prt_shoot_fan_host", "prt_pack_arrivals", "prt_bracket_workspace_bytes", "prt_count_brackets", "prt_fill_brackets",
    "prt_refine_workspace_bytes", "prt_refine_brackets", "prt_received_angles", "prt_eigen_table", "prt_flat_earth", "prt_iw_workspace_bytes", "prt_iw_perturb", "prt_launch_count", "prt_measure_fp64_peak",
)


class PrtError(RuntimeError):
    pass


_lib = None


def load():
    """Load libprt.so and declare the prototypes of include/prt.h.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PrtError(
            f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; "
            "g.build()').  pygenray_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    d, i32, i64, vp = C.c_double, C.c_int32, C.c_int64, C.c_void_p
    L.prt_abi_version.restype = C.c_int
    L.prt_last_error.restype = C.c_char_p
    L.prt_set_concurrency.restype = C.c_int
    L.prt_set_concurrency.argtypes = [C.c_int]
    L.prt_set_lanes.restype = C.c_int
    L.prt_set_lanes.argtypes = [C.c_int]
    L.prt_device_count.restype = C.c_int
    L.prt_env_create.restype = C.c_int
    L.prt_env_create.argtypes = [C.c_int, i32, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, C.POINTER(vp)]
    L.prt_ensemble_create.restype = C.c_int
    L.prt_ensemble_create.argtypes = [C.c_int, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, C.POINTER(vp)]
    L.prt_ensemble_create_dev.restype = C.c_int
    L.prt_ensemble_create_dev.argtypes = [C.c_int, i32, i32, i32, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, C.POINTER(vp)]
    L.prt_env_destroy.restype = C.c_int
    L.prt_env_destroy.argtypes = [vp]
    L.prt_env_device_bytes.restype = i64
    L.prt_env_device_bytes.argtypes = [vp]
    fan_tail = [d, d, d, d, d, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, i32]
    L.prt_shoot_fan.restype = C.c_int
    L.prt_shoot_fan.argtypes = [vp, i64, vp, vp] + fan_tail + [vp]
    L.prt_shoot_ensemble.restype = C.c_int
    L.prt_shoot_ensemble.argtypes = [vp, i32, i64, vp] + fan_tail + [vp]
    L.prt_shoot_fan_host.restype = C.c_int
    L.prt_shoot_fan_host.argtypes = [vp, i64, vp, vp] + fan_tail
    L.prt_pack_arrivals.restype = C.c_int
    L.prt_pack_arrivals.argtypes = [i64, vp, vp, vp, vp, vp, vp]
    L.prt_count_brackets.restype = C.c_int
    L.prt_bracket_workspace_bytes.restype = i64
    L.prt_bracket_workspace_bytes.argtypes = [i64, i32]
    L.prt_count_brackets.argtypes = [i64, vp, i32, vp, vp, vp, vp]
    L.prt_fill_brackets.restype = C.c_int
    L.prt_fill_brackets.argtypes = [i64, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.prt_refine_workspace_bytes.restype = i64
    L.prt_refine_workspace_bytes.argtypes = [i64]
    L.prt_refine_brackets.restype = C.c_int
    L.prt_refine_brackets.argtypes = [vp, i64, vp, vp, vp, vp, vp, d, d, d, d, d, i32, d, i32,
                                      vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    L.prt_received_angles.restype = C.c_int
    L.prt_received_angles.argtypes = [i64, d, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.prt_eigen_table.restype = C.c_int
    L.prt_eigen_table.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp, d, i32, i32, vp, vp, vp, vp, vp]
    L.prt_flat_earth.restype = C.c_int
    L.prt_flat_earth.argtypes = [i32, i32, i32, vp, vp, vp, i32, vp, vp]
    L.prt_iw_workspace_bytes.restype = i64
    L.prt_iw_workspace_bytes.argtypes = [i32, i32, i32]
    L.prt_iw_perturb.restype = C.c_int
    L.prt_iw_perturb.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, d, vp, vp, vp]
    L.prt_launch_count.restype = i64
    L.prt_measure_fp64_peak.restype = d
    L.prt_measure_fp64_peak.argtypes = [C.c_int, i32]
    _lib = L
    return L


def check(rc, what):
    if rc != PRT_SUCCESS:
        msg = load().prt_last_error()
        raise PrtError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def require_gpu():
    """Raise unless libprt.so is loadable AND a CUDA device is visible."""
    L = load()
    if L.prt_device_count() < 1:
        raise PrtError("no CUDA device visible: pygenray_b200 runs on B200 GPUs only (no CPU fallback)")
    return L


def launch_count():
    return int(load().prt_launch_count())
