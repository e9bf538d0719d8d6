"""ctypes binding of libranslice_b200.so (C ABI: include/ranslice_b200.h).

The library is the product: there is no CPU fallback.  Loading fails loudly if the shared
object has not been built (``python __graft_entry__.py build``).
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("RS_B200_LIB") or os.path.join(_PKG, "libranslice_b200.so")   # RS_B200_LIB: experiment builds (tools/sweep_variants.sh)

RS_ABI_VERSION = 2
FLAG_UE_CAP, FLAG_BURST_CAP, FLAG_ACTION_CLAMP, FLAG_SAME_SLOT_DEP, FLAG_MTC_QUEUE_CAP = 1, 2, 4, 8, 16

class RsConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n_envs", C.c_int32),
                ("n_prbs", C.c_int32), ("n_embb", C.c_int32), ("n_mmtc", C.c_int32),
                ("slots_per_step", C.c_int32), ("max_ues", C.c_int32), ("max_bursts", C.c_int32),
                ("mtc_queue_cap", C.c_int32), ("kernel_variant", C.c_int32), ("l1_mux", C.c_int32),
                ("penalty", C.c_double), ("prop_A", C.c_double), ("prop_B", C.c_double),
                ("base_seed", C.c_uint64), ("first_env_id", C.c_uint64)]


class RsTables(C.Structure):
    _fields_ = [("trace", C.c_void_p), ("mcs_rate", C.c_void_p), ("mcs_snr", C.c_void_p),
                ("mcs_order", C.c_void_p), ("mcs_mod", C.c_void_p)]


_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise NativeLibraryMissing(
            "libranslice_b200.so is not built (%s). Run `python __graft_entry__.py build`; "
            "there is no CPU fallback for the step path." % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, i32 = C.c_void_p, C.c_int32
    L.rs_create.argtypes = [C.POINTER(RsConfig), C.POINTER(RsTables), C.POINTER(vp)]
    L.rs_destroy.argtypes = [vp]
    L.rs_reset.argtypes = [vp, vp]
    L.rs_step.argtypes = [vp] + [vp] * 6
    L.rs_step_device.argtypes = [vp] + [vp] * 6 + [vp]
    L.rs_step_async.argtypes = [vp] + [vp] * 6 + [C.POINTER(i32)]
    L.rs_wait.argtypes = [vp, i32]
    L.rs_get_info.argtypes = [vp, i32, vp, vp]
    L.rs_get_n_ues.argtypes = [vp, vp]
    L.rs_state_size.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.rs_get_state.argtypes = [vp, vp, C.c_size_t]
    L.rs_set_state.argtypes = [vp, vp, C.c_size_t]
    L.rs_get_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.rs_n_variables.argtypes = [vp]
    L.rs_active_variant.argtypes = [vp]
    L.rs_active_variant.restype = C.c_int
    L.rs_set_debug_check.argtypes = [vp, i32]
    L.rs_get_diag.argtypes = [vp, C.POINTER(C.c_double), i32]
    L.rs_set_profiling.argtypes = [vp, i32]
    L.rs_get_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.rs_set_route_limits.argtypes = [vp, i32, i32, i32, i32]
    L.rs_get_routes.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.rs_set_heavy_threshold.argtypes = [vp, i32, i32]
    L.rs_last_error.restype = C.c_char_p
    for name in ("rs_create", "rs_destroy", "rs_reset", "rs_step", "rs_step_device", "rs_get_info", "rs_get_n_ues",
                 "rs_state_size", "rs_get_state", "rs_set_state", "rs_get_counters", "rs_n_variables",
                 "rs_set_profiling", "rs_get_profile", "rs_set_debug_check", "rs_get_diag", "rs_step_async", "rs_wait",
                 "rs_set_route_limits", "rs_get_routes", "rs_set_heavy_threshold"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError("libranslice_b200: error %d: %s" % (rc, lib().rs_last_error().decode()))
