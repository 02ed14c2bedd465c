"""ctypes binding of libtapir_cuda.so (include/tapir_cuda.h).  No torch, no CPU fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtapir_cuda.so")

TAPIR_OK = 0
TAPIR_E_INVALID, TAPIR_E_CUDA, TAPIR_E_NOMEM, TAPIR_E_STATE = -1, -2, -3, -4
WANT_PROFILE = 1
FP32_MODE = 2
MAX_EPOCHS = 64
MAX_TIMES = 4096


class TapirCudaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libtapir_cuda error {0}: {1}".format(code, message))
        self.code = code


class Problem(ctypes.Structure):
    _fields_ = [
        ("n_loci", ctypes.c_int64),
        ("offsets", ctypes.c_void_p),
        ("rates", ctypes.c_void_p),
        ("n_grid", ctypes.c_int64),
        ("n_times", ctypes.c_int32),
        ("time_idx", ctypes.c_void_p),
        ("n_epochs", ctypes.c_int32),
        ("epoch_ab", ctypes.c_void_p),
        ("flags", ctypes.c_uint32),
    ]


RESULT_FIELDS = ["pi_net", "pi_times", "integral", "abserr", "max_pi", "argmax_t", "argmax_epoch", "mean_rate",
                 "n_finite", "n_multi_panel"]


class Results(ctypes.Structure):
    _fields_ = [(name, ctypes.c_void_p) for name in RESULT_FIELDS]


class Timing(ctypes.Structure):
    _fields_ = [
        ("total_ms", ctypes.c_float),
        ("profile_ms", ctypes.c_float),
        ("quad_ms", ctypes.c_float),
        ("finalize_ms", ctypes.c_float),
        ("launches", ctypes.c_int32),
        ("profile_tile_points", ctypes.c_int32),
        ("profile_moment_form", ctypes.c_int32),
        ("moment_epochs", ctypes.c_int32),
        ("profile_fp32", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class MomEpoch(ctypes.Structure):
    _fields_ = [
        ("eligible", ctypes.c_int32),
        ("tile", ctypes.c_int32),
        ("hb", ctypes.c_int32),
        ("tb", ctypes.c_double),
        ("p", ctypes.c_double * 32),
        ("r_min", ctypes.c_double),
        ("r_max", ctypes.c_double),
        ("w_lo", ctypes.c_double),
        ("w_hi", ctypes.c_double),
    ]


# every symbol include/tapir_cuda.h declares: (restype, argtypes)
_vp, _i64, _dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
SYMBOLS = {
    "tapir_cuda_version": (ctypes.c_int, []),
    "tapir_cuda_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "tapir_cuda_last_error": (ctypes.c_char_p, [_vp]),
    "tapir_cuda_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "tapir_cuda_destroy": (ctypes.c_int, [_vp]),
    "tapir_cuda_host_alloc": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_size_t]),
    "tapir_cuda_host_free": (ctypes.c_int, [_vp]),
    "tapir_cuda_run": (ctypes.c_int, [_vp, ctypes.POINTER(Problem), ctypes.POINTER(Results)]),
    "tapir_cuda_upload": (ctypes.c_int, [_vp, ctypes.POINTER(Problem)]),
    "tapir_cuda_compute": (ctypes.c_int, [_vp, ctypes.POINTER(Timing)]),
    "tapir_cuda_download": (ctypes.c_int, [_vp, ctypes.POINTER(Results)]),
    "tapir_cuda_device_results": (ctypes.c_int, [_vp, ctypes.POINTER(Results)]),
    "tapir_cuda_device_small_block": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_size_t),
                                                     ctypes.POINTER(ctypes.c_size_t * 9)]),
    "tapir_cuda_rank": (ctypes.c_int, [_vp, ctypes.c_int32, _vp, _vp, _vp]),
    "tapir_cuda_gather_create": (ctypes.c_int, [_vp, ctypes.c_size_t, ctypes.c_int, _vp]),
    "tapir_cuda_gather_attach": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, ctypes.c_int]),
    "tapir_cuda_gather_read": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t]),
    "tapir_cuda_gather_detach": (ctypes.c_int, [_vp]),
    "tapir_cuda_set_serial_kernels": (ctypes.c_int, [_vp, ctypes.c_int]),
    "tapir_cuda_transfer_bytes": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    "tapir_cuda_pi_matrix": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "tapir_cuda_nansum_rows": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "tapir_cuda_site_integrals": (ctypes.c_int, [_vp, _dbl, _dbl, _vp, _i64, _vp, _vp, _vp]),
    "tapir_mom_epoch_plan": (ctypes.c_int, [_dbl, _dbl, _i64, ctypes.c_int32, ctypes.c_int32, _dbl, ctypes.POINTER(MomEpoch)]),
    "tapir_ingest_rates": (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), _i64, ctypes.c_int, _dbl, _vp, _vp, _i64,
                                          ctypes.POINTER(_i64)]),
    "tapir_ingest_open": (ctypes.c_int, [ctypes.POINTER(ctypes.c_char_p), _i64, ctypes.c_int, ctypes.POINTER(_vp), _vp,
                                         ctypes.POINTER(_i64)]),
    "tapir_ingest_copy": (ctypes.c_int, [_vp, _dbl, _vp, _i64]),
    "tapir_ingest_close": (ctypes.c_int, [_vp]),
    "tapir_ingest_last_error": (ctypes.c_char_p, []),
    "tapir_cuda_fp64_peak": (ctypes.c_int, [_vp, ctypes.POINTER(_dbl), ctypes.POINTER(ctypes.c_float)]),
    "tapir_cuda_flush_l2": (ctypes.c_int, [_vp]),
}

_lib = None


def load():
    """Load libtapir_cuda.so from the package directory.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "{0} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C tapir_b200/csrc` (there is no CPU fallback)".format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc, ctx=None):
    if rc != TAPIR_OK:
        msg = load().tapir_cuda_last_error(ctx)
        raise TapirCudaError(rc, msg.decode("utf-8", "replace") if msg else "?")
