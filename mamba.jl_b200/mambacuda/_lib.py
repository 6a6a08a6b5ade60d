"""ctypes binding of libmambacuda.so (the C ABI in include/mambacuda.h).

The shared library is built in-tree by mamba.jl_b200/build.py.  There is no fallback of any
kind: if the library is missing, or no CUDA device is usable, the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCU_LIB_PATH") or os.path.join(_HERE, "libmambacuda.so")   # override: kernel-variant experiments only
MAX_BLOCK_NODES = 8

OK, ERR_ARG, ERR_DIM, ERR_STATE, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
TPL = {"line": 0, "seeds": 1, "rats": 2, "pumps": 3, "glm": 4, "surgical": 5, "dyes": 6, "salm": 7, "equiv": 8, "blocker": 9, "stacks": 10, "magnesium": 11, "oxford": 12, "epil": 13}
KIND = {"amwg": 0, "slice_uni": 1, "slice_multi": 2, "rwm": 3, "nuts": 4, "hmc": 5, "amm": 6, "gibbs": 7, "mala": 8}
ADAPT = {"all": 0, "burnin": 1, "none": 2}
PROPOSAL = {"normal": 0, "symuniform": 1, "symtriangular": 2, "cosine": 3, "epanechnikov": 4, "biweight": 5, "triweight": 6}
GRAD = {"analytic": 0, "forward": 1, "central": 2}
RUN_NO_STORE, RUN_FORCE_GENERIC, RUN_GLM_REFERENCE, RUN_PARTIAL, RUN_ASYNC, RUN_MPSRF = 1, 2, 4, 8, 16, 32


class BlockDesc(C.Structure):
    """mcu_block_desc"""
    _fields_ = [
        ("kind", C.c_int32), ("n_nodes", C.c_int32), ("nodes", C.c_int32 * MAX_BLOCK_NODES),
        ("transform", C.c_int32), ("adapt", C.c_int32), ("batchsize", C.c_int32),
        ("proposal", C.c_int32), ("L", C.c_int32), ("grad", C.c_int32), ("max_depth", C.c_int32),
        ("n_scale", C.c_int32), ("target", C.c_double), ("epsilon", C.c_double),
        ("beta", C.c_double), ("amm_scale", C.c_double), ("scale", C.POINTER(C.c_double)),
    ]


class MambaCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmambacuda error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None

# every symbol include/mambacuda.h declares
SYMBOLS = [
    "mcu_create", "mcu_destroy", "mcu_last_error", "mcu_abi_version", "mcu_set_data", "mcu_set_scheme",
    "mcu_dims", "mcu_names", "mcu_tune_size", "mcu_set_inits", "mcu_run", "mcu_kept", "mcu_get_state",
    "mcu_set_state", "mcu_logpdf", "mcu_gradlogpdf", "mcu_glm_gradient", "mcu_minmax", "mcu_link_codes", "mcu_moments",
    "mcu_gelman_from_moments", "mcu_gelman", "mcu_summarystats", "mcu_summary_sums",
    "mcu_summary_from_sums", "mcu_summary_streaming", "mcu_set_rng_mode", "mcu_device_count",
    "mcu_launch_count", "mcu_last_kernel_ms", "mcu_fp64_peak_tflops",
    "mcu_chains_quantile", "mcu_chains_hpd", "mcu_chains_autocor", "mcu_chains_changerate", "mcu_chains_gelman",
    "mcu_chains_geweke", "mcu_chains_heidel", "mcu_chains_raftery", "mcu_chains_summarystats", "mcu_factor_counts", "mcu_factor_parents", "mcu_logpdf_nodes", "mcu_predict",
    "mcu_diag_sizes", "mcu_monitor_links", "mcu_n_kept", "mcu_diag_round1", "mcu_diag_round2", "mcu_diag_finish",
    "mcu_comm_unique_id", "mcu_comm_init", "mcu_comm_size", "mcu_diag_global", "mcu_wait", "mcu_get_samples", "mcu_work_count",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python mamba.jl_b200/build.py` "
            "(there is no CPU fallback for the engine)")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    i64 = C.c_int64
    L.mcu_create.argtypes = [C.c_int, i64, i64, C.c_int, C.c_uint64, C.POINTER(vp)]
    L.mcu_destroy.argtypes = [vp]
    L.mcu_last_error.restype = C.c_char_p
    L.mcu_last_error.argtypes = [vp]
    L.mcu_set_data.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(i64), dp]
    L.mcu_set_scheme.argtypes = [vp, C.c_int, C.POINTER(BlockDesc)]
    L.mcu_dims.argtypes = [vp, ip, ip, ip]
    L.mcu_names.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t]
    L.mcu_tune_size.argtypes = [vp, C.POINTER(i64)]
    L.mcu_set_inits.argtypes = [vp, dp, i64, C.c_double]
    L.mcu_run.argtypes = [vp, i64, i64, i64, dp, C.c_uint32]
    L.mcu_kept.restype = i64
    L.mcu_kept.argtypes = [i64, i64, i64, i64]
    L.mcu_get_state.argtypes = [vp, dp, dp, C.POINTER(i64)]
    L.mcu_set_state.argtypes = [vp, dp, dp, i64]
    L.mcu_logpdf.argtypes = [vp, C.c_int, i64, dp, dp, dp]
    L.mcu_gradlogpdf.argtypes = [vp, C.c_int, C.c_int, i64, dp, dp, dp, dp]
    L.mcu_glm_gradient.argtypes = [vp, C.c_int, dp, dp, dp]
    L.mcu_minmax.argtypes = [vp, dp]
    L.mcu_link_codes.argtypes = [vp, C.c_int, dp, ip]
    L.mcu_moments.argtypes = [vp, ip, dp, dp, C.POINTER(i64)]
    L.mcu_gelman_from_moments.argtypes = [i64, C.c_int, dp, dp, C.c_double, dp]
    L.mcu_gelman.argtypes = [vp, C.c_double, C.c_int, dp]
    L.mcu_summarystats.argtypes = [vp, C.c_int, C.c_int, dp]
    L.mcu_summary_sums.argtypes = [vp, dp, dp]
    L.mcu_summary_from_sums.argtypes = [i64, C.c_int, dp, dp, dp]
    L.mcu_summary_streaming.argtypes = [vp, dp]
    L.mcu_set_rng_mode.argtypes = [vp, C.c_int, dp, C.c_size_t]
    L.mcu_launch_count.restype = i64
    L.mcu_launch_count.argtypes = [vp]
    L.mcu_fp64_peak_tflops.restype = C.c_double
    L.mcu_fp64_peak_tflops.argtypes = [vp]
    L.mcu_last_kernel_ms.restype = C.c_double
    L.mcu_last_kernel_ms.argtypes = [vp]
    L.mcu_chains_quantile.argtypes = [dp, i64, C.c_int, i64, dp, C.c_int, dp]
    L.mcu_chains_hpd.argtypes = [dp, i64, C.c_int, i64, C.c_double, dp]
    L.mcu_chains_autocor.argtypes = [dp, i64, C.c_int, i64, C.POINTER(i64), C.c_int, dp]
    L.mcu_chains_changerate.argtypes = [dp, i64, C.c_int, i64, dp]
    L.mcu_chains_gelman.argtypes = [dp, i64, C.c_int, i64, C.c_double, ip, C.c_int, dp]
    L.mcu_chains_geweke.argtypes = [dp, i64, C.c_int, i64, C.c_double, C.c_double, C.c_int, C.c_int, dp]
    L.mcu_chains_heidel.argtypes = [dp, i64, C.c_int, i64, C.c_double, C.c_double, C.c_int, C.c_int, i64, dp]
    L.mcu_chains_raftery.argtypes = [dp, i64, C.c_int, i64, C.c_double, C.c_double, C.c_double, C.c_double, i64, i64, dp]
    L.mcu_factor_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.mcu_factor_parents.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)]
    L.mcu_logpdf_nodes.argtypes = [C.c_void_p, C.c_uint32, i64, dp, dp]
    L.mcu_predict.argtypes = [C.c_void_p, i64, dp, C.c_uint32, dp, C.POINTER(i64)]
    L.mcu_chains_summarystats.argtypes = [dp, i64, C.c_int, i64, C.c_int, C.c_int, dp]
    L.mcu_diag_sizes.argtypes = [C.c_int, ip, ip]
    L.mcu_monitor_links.argtypes = [vp, ip]
    L.mcu_n_kept.argtypes = [vp, C.POINTER(i64)]
    L.mcu_diag_round1.argtypes = [vp, dp]
    L.mcu_diag_round2.argtypes = [vp, C.c_int, dp, dp]
    L.mcu_diag_finish.argtypes = [i64, C.c_int, C.c_double, ip, C.c_int, dp, dp, dp, dp, ip, C.POINTER(C.c_double)]
    L.mcu_comm_unique_id.argtypes = [C.c_char_p]
    L.mcu_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.mcu_comm_size.argtypes = [vp, ip, ip]
    L.mcu_diag_global.argtypes = [vp, C.c_double, C.c_int, dp, dp, ip, C.POINTER(C.c_double)]
    L.mcu_wait.argtypes = [vp]
    L.mcu_work_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mcu_get_samples.argtypes = [vp, dp]
    _lib = L
    return L
