# coding: utf-8
"""
ctypes binding of ``libjoeys2t_b200.so`` (the C ABI declared in ``include/joeys2t_b200.h``).

The shared library is built in-tree (``joeys2t_b200/csrc/libjoeys2t_b200.so``) by
:func:`build` / ``__graft_entry__.build()`` with ``nvcc -gencode arch=compute_100a,code=sm_100a``.
There is **no CPU fallback**: if the library is missing or no CUDA device is usable, every entry
point raises.
"""
import ctypes
import os
import subprocess
import threading
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
# JS2T_LIB selects another build of the same library (tuning A/B runs: tools/build_variant.py)
LIB_PATH = Path(os.environ["JS2T_LIB"]).resolve() if os.environ.get("JS2T_LIB") else CSRC / "libjoeys2t_b200.so"
SOURCES = ["fbank_kernels.cu", "ingest_kernels.cu", "capi.cu"]
HEADERS = ["js2t_internal.h", "mel_structure.inc", "../../include/joeys2t_b200.h"]

# status codes (include/joeys2t_b200.h)
OK, ERR_INVALID, ERR_CUDA, ERR_SHORT_INPUT, ERR_TABLES, ERR_NCCL, ERR_STATE = range(7)
CMVN_NONE, CMVN_UTTERANCE, CMVN_GLOBAL, CMVN_STATS_ONLY = range(4)
LAYOUT_RAGGED, LAYOUT_PADDED = 0, 1
MASK_VALUE_MEAN, MASK_VALUE_CONST = 0, 1

# every symbol include/joeys2t_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = [
    "js2t_version", "js2t_last_error", "js2t_num_frames", "js2t_ctx_create", "js2t_ctx_destroy",
    "js2t_ctx_set_tables", "js2t_reference_mel_bank", "js2t_plan_create", "js2t_plan_create_features", "js2t_plan_destroy", "js2t_plan_destroy_completed",
    "js2t_plan_total_frames", "js2t_plan_out_rows", "js2t_plan_get_frames", "js2t_plan_get_out_rows",
    "js2t_plan_set_cmvn", "js2t_plan_set_global_stats", "js2t_plan_set_masks", "js2t_plan_set_dither", "js2t_fbank_execute",
    "js2t_features_execute", "js2t_plan_enable_profiling", "js2t_plan_kernel_times_ms",
    "js2t_plan_set_option", "js2t_plan_debug_times", "js2t_plan_utt_stats", "js2t_plan_copy_utt_stats", "js2t_global_stats_accumulate",
    "js2t_global_stats_allreduce", "js2t_nccl_unique_id", "js2t_nccl_comm_create", "js2t_nccl_comm_destroy",
    "js2t_global_stats_finalize", "js2t_normalize_execute", "js2t_plan_copy_global_stats",
    "js2t_reformat_48k_to_16k", "js2t_pack_pcm", "js2t_batch_fbank", "js2t_specaug_replay",
]


class BatchOpts(ctypes.Structure):
    """``js2t_batch_opts`` (include/joeys2t_b200.h)."""
    _fields_ = [("layout", ctypes.c_int), ("pad_tmax", ctypes.c_int), ("pad_value", ctypes.c_float),
                ("cmvn_mode", ctypes.c_int), ("norm_means", ctypes.c_int), ("norm_vars", ctypes.c_int),
                ("before", ctypes.c_int), ("global_mean80", ctypes.c_void_p), ("global_istd80", ctypes.c_void_p),
                ("n_fmask", ctypes.c_int), ("n_tmask", ctypes.c_int), ("mask_table", ctypes.c_void_p),
                ("mask_value_mode", ctypes.c_int), ("mask_value_const", ctypes.c_float),
                ("max_frames", ctypes.c_void_p)]


class Js2tError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[js2t status {code}] {message}")
        self.code = code


def nvcc_command(out: Path = LIB_PATH, extra=()):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return [
        nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "--shared", "-Xcompiler", "-fPIC", *extra, "-o", str(out), *[str(CSRC / s) for s in SOURCES],
        "-ldl"
    ]


def kernel_source_sha16() -> str:
    """Hash of the device code of the hot path: profiles keyed by it (ncu counters,
    DRAM traffic) are only attached to a bench line when they belong to the build being measured."""
    import hashlib
    h = hashlib.sha256()
    for f in ("fbank_kernels.cu", "js2t_internal.h", "mel_structure.inc"):  # device code of the hot path
        h.update((CSRC / f).resolve().read_bytes())
    return h.hexdigest()[:16]


def is_stale() -> bool:
    if os.environ.get("JS2T_LIB"):
        return False
    if not LIB_PATH.is_file():
        return True
    t = LIB_PATH.stat().st_mtime
    return any((CSRC / f).resolve().stat().st_mtime > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = nvcc_command(extra=("-Xptxas", "-v") if verbose else ())
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=str(CSRC))
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({' '.join(cmd)}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def _declare(lib):
    c = ctypes
    vp, i32, i64, f32 = c.c_void_p, c.c_int, c.c_int64, c.c_float
    P = c.POINTER
    lib.js2t_version.restype = i32
    lib.js2t_last_error.restype = c.c_char_p
    lib.js2t_num_frames.restype = i64
    lib.js2t_num_frames.argtypes = [i64]
    lib.js2t_ctx_create.argtypes = [i32, P(vp)]
    lib.js2t_ctx_destroy.argtypes = [vp]
    lib.js2t_ctx_set_tables.argtypes = [vp, vp, vp]
    lib.js2t_reference_mel_bank.argtypes = [vp]
    lib.js2t_plan_create.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, f32, P(vp)]
    lib.js2t_plan_create_features.argtypes = [vp, i32, vp, vp, i32, i32, f32, P(vp)]
    lib.js2t_plan_destroy.argtypes = [vp]
    lib.js2t_plan_destroy_completed.argtypes = [vp]
    lib.js2t_plan_total_frames.restype = i64
    lib.js2t_plan_total_frames.argtypes = [vp]
    lib.js2t_plan_out_rows.restype = i64
    lib.js2t_plan_out_rows.argtypes = [vp]
    lib.js2t_plan_get_frames.argtypes = [vp, vp]
    lib.js2t_plan_get_out_rows.argtypes = [vp, vp]
    lib.js2t_plan_set_cmvn.argtypes = [vp, i32, i32, i32, i32]
    lib.js2t_plan_set_global_stats.argtypes = [vp, vp, vp, vp]
    lib.js2t_plan_set_masks.argtypes = [vp, i32, i32, vp, i32, f32, vp]
    lib.js2t_plan_set_dither.argtypes = [vp, vp]
    lib.js2t_fbank_execute.argtypes = [vp, vp, vp, vp]
    lib.js2t_features_execute.argtypes = [vp, vp, vp, vp]
    lib.js2t_plan_enable_profiling.argtypes = [vp, i32]
    lib.js2t_plan_kernel_times_ms.argtypes = [vp, vp, i32, P(i32)]
    lib.js2t_plan_set_option.argtypes = [vp, c.c_char_p, i32]
    lib.js2t_plan_debug_times.argtypes = [vp, vp, i64]
    lib.js2t_plan_utt_stats.argtypes = [vp, P(vp)]
    lib.js2t_plan_copy_utt_stats.argtypes = [vp, vp, vp]
    lib.js2t_global_stats_accumulate.argtypes = [vp, vp, vp]
    lib.js2t_global_stats_allreduce.argtypes = [vp, vp, vp]
    lib.js2t_nccl_unique_id.argtypes = [vp]
    lib.js2t_nccl_comm_create.argtypes = [vp, i32, i32, i32, P(vp)]
    lib.js2t_nccl_comm_destroy.argtypes = [vp]
    lib.js2t_global_stats_finalize.argtypes = [vp, vp, vp]
    lib.js2t_normalize_execute.argtypes = [vp, vp, vp]
    lib.js2t_plan_copy_global_stats.argtypes = [vp, vp, vp]
    lib.js2t_reformat_48k_to_16k.argtypes = [vp, vp, i32, i64, vp, vp, vp]
    lib.js2t_pack_pcm.argtypes = [i32, vp, vp, vp, vp, i64, i32]
    lib.js2t_specaug_replay.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, i32, c.c_double, vp]
    lib.js2t_batch_fbank.argtypes = [vp, i32, vp, vp, vp, P(BatchOpts), vp, i64, vp, vp]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is c.c_int and name not in ("js2t_version",):
            fn.restype = i32


def load():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.is_file():
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                    "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                    "There is no CPU fallback.")
            lib = ctypes.CDLL(str(LIB_PATH))
            _declare(lib)
            _lib = lib
    return _lib


def check(status: int):
    if status != OK:
        raise Js2tError(status, load().js2t_last_error().decode("utf-8", "replace"))
