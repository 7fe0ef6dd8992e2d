"""ctypes binding of libvmsm.so (include/vmsm.h).  Thin on purpose: argument types only.

There is no CPU fallback.  If the shared library is missing, or the process has no CUDA device, the loader /
``Context`` raise ``VmsmError`` -- nothing here (or anywhere in this package) imports ``oracle/``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VMSM_LIB") or os.path.join(_HERE, "libvmsm.so")  # VMSM_LIB: A/B builds of the same ABI

OK = 0
ERR_INVALID, ERR_CUDA, ERR_POINT, ERR_NOMEM, ERR_UNSUPPORTED, ERR_TIMEOUT = -1, -2, -3, -4, -5, -6
CURVE_ED25519, CURVE_BN256_G1, CURVE_BN256_G2 = 0, 1, 2
OPT_WINDOW_BITS, OPT_PHASE_TIMING, OPT_SORT_BUCKETS, OPT_CHECK_POINTS, OPT_REDUCE_RADIX = 1, 2, 3, 4, 5
OPT_QUAD_THRESHOLD, OPT_ASYNC_TAIL, OPT_CAP_FACTOR, OPT_SHARD_SEQ, OPT_ASYNC_SORT = 6, 7, 8, 9, 10
OPT_SORT_BLOCKS = 11
OPT_FOLD_QUAD_MAX = 12
OPT_BN_QUAD_ACC = 13
OPT_PRE_SETS, OPT_PRE_MIN_TERMS, OPT_SEG_LEN, OPT_SEG_MODE, OPT_HOST_NORMALIZE = 14, 15, 16, 17, 18
OPT_DUAL_HEAD = 19
OPT_BLOCK_SORT, OPT_BLOCK_SORT_MIN, OPT_ACC_CARVEOUT = 20, 21, 22
OPT_BN_PRE_SETS, OPT_BN_SEG_LEN, OPT_BN_QUAD_FIX = 23, 24, 25
FOLD_WITNESS, FOLD_FORM = 0, 1
AXPY_ADD_SCALED, AXPY_SCALE_ADD, AXPY_SCALE = 0, 1, 2
PHASES = ("digits", "scan", "scatter", "order", "handoff", "accumulate", "reduce", "final")


class VmsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vmsm error {code}: {msg}")
        self.code = code


_u64, _u32, _i32, _i64 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int32, ctypes.c_int64
_p = ctypes.c_void_p
_pu64 = ctypes.POINTER(ctypes.c_uint64)

# name -> argtypes; every function returns int32 unless listed in _RESTYPES.  Mirrors include/vmsm.h 1:1
# (tests/test_abi.py parses the header and checks both directions).
SIGNATURES = {
    "vmsm_version": [],
    "vmsm_last_error": [],
    "vmsm_device_count": [ctypes.POINTER(_i32)],
    "vmsm_ctx_create": [_i32, _pu64],
    "vmsm_ctx_destroy": [_u64],
    "vmsm_ctx_set_option": [_u64, _i32, _i64],
    "vmsm_sync": [_u64],
    "vmsm_timer_start": [_u64],
    "vmsm_timer_stop": [_u64, ctypes.POINTER(ctypes.c_float)],
    "vmsm_phase_times": [_u64, ctypes.POINTER(ctypes.c_double), _pu64],
    "vmsm_launch_count": [_u64, _pu64],
    "vmsm_points_upload": [_u64, _i32, _p, _u64, _pu64],
    "vmsm_points_fixed_base": [_u64, _i32, _p, _u64, _u64, _pu64],
    "vmsm_points_precompute": [_u64, _u64, _u32],
    "vmsm_points_download": [_u64, _u64, _u64, _u64, _p],
    "vmsm_points_text": [_u64, _u64, _u64, _u64, _p, _u64, _pu64],
    "vmsm_points_download_ptr": [_u64, _u64, _u64, _u64, ctypes.POINTER(_p)],
    "vmsm_scalars_download_ptr": [_u64, _u64, _u64, _u64, ctypes.POINTER(_p)],
    "vmsm_points_text_ptr": [_u64, _u64, _u64, _u64, ctypes.POINTER(_p), _pu64],
    "vmsm_points_count": [_u64, _u64, _pu64],
    "vmsm_points_free": [_u64, _u64],
    "vmsm_scalars_upload": [_u64, _p, _u64, _pu64],
    "vmsm_scalars_synth": [_u64, _i32, _u64, _u64, _pu64],
    "vmsm_scalars_download": [_u64, _u64, _u64, _u64, _p],
    "vmsm_scalars_free": [_u64, _u64],
    "vmsm_msm": [_u64, _u64, _u64, _u64, _p, _p],
    "vmsm_msm_ext": [_u64, _u64, _u64, _u64, _u64, _u64, _u64, _p, _p],
    "vmsm_points_concat": [_u64, _u64, _u64, _u64, _u64, _u64, _u64, _pu64],
    "vmsm_mailbox_create": [_u64, _u32, _p],
    "vmsm_mailbox_open_ipc": [_u64, _p, _u32, _u32],
    "vmsm_mailbox_open_local": [_u64, _u64, _u32],
    "vmsm_msm_dev_shard": [_u64, _u64, _u64, _u64, _u64, _u64, _u32, _u32],
    "vmsm_msm_async": [_u64, _u64, _u64, _u64, _p, _u32],
    "vmsm_msm_dev": [_u64, _u64, _u64, _u64, _u64, _u64, _u32],
    "vmsm_msm_dev_ext": [_u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _p, _u32],
    "vmsm_msm_dev_ext_dot": [_u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u64, _u32],
    "vmsm_scalars_fold": [_u64, _u64, _u64, _p, _i32],
    "vmsm_scalars_axpy": [_u64, _u64, _u64, _u64, _u64, _u64, _p, _i32],
    "vmsm_scalars_dot": [_u64, _u64, _u64, _u64, _u64, _u64, _p],
    "vmsm_scalars_text_ptr": [_u64, _u64, _u64, _u64, _i32, ctypes.POINTER(_p), _pu64],
    "vmsm_result_affine": [_u64, _u32, _p],
    "vmsm_result_extended": [_u64, _u32, _p],
    "vmsm_fold": [_u64, _u64, _u64, _p],
    "vmsm_lincomb_async": [_u64, _i32, _p, _p, _u64, _u32],
    "vmsm_lincomb": [_u64, _i32, _p, _p, _u64, _p],
    "vmsm_host_alloc": [_u64, ctypes.POINTER(_p)],
    "vmsm_host_free": [_p],
    "vmsm_selftest_fe": [_u64, _i32, _p, _p, _u64, _p],
    "vmsm_microbench_imad": [_u64, ctypes.POINTER(ctypes.c_double)],
}
_RESTYPES = {"vmsm_last_error": ctypes.c_char_p}

_lib = None


def load():
    """Load libvmsm.so (built in-tree by ``__graft_entry__.build()`` / ``make -C verifiable_mpc_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VmsmError(ERR_CUDA, f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _i32)
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise VmsmError(rc, (load().vmsm_last_error() or b"").decode("utf-8", "replace"))
