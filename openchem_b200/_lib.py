"""ctypes binding of libocgcn.so (the C ABI declared in include/ocgcn.h) and its in-tree build.

There is deliberately no fallback: if the shared library is missing or the device is not a B200 the
product path raises. The oracle under ``oracle/`` is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libocgcn.so")
SOURCES = ["ocgcn_api.cu", "ocgcn_tile.cuh", "ocgcn_aux.cuh", "ocgcn_common.cuh", "ocgcn_tc.cuh", "ocgcn_dw.cuh", "ocgcn_compact.cuh", "ocgcn_head.cuh"]
HEADER = os.path.join(os.path.dirname(_HERE), "include", "ocgcn.h")

MAX_LAYERS = 16
MODES = {"fp32": 0, "tf32": 1, "bf16": 2}

_f32p = ctypes.POINTER(ctypes.c_float)


class Desc(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int32), ("N", ctypes.c_int32), ("L", ctypes.c_int32), ("E", ctypes.c_int32),
        ("F", ctypes.c_int32 * (MAX_LAYERS + 1)),
        ("mode", ctypes.c_int32), ("training", ctypes.c_int32), ("has_bias", ctypes.c_int32), ("need_dx", ctypes.c_int32),
        ("eps", ctypes.c_float), ("momentum", ctypes.c_float),
        ("x_stride", ctypes.c_int64 * 3), ("adj_stride", ctypes.c_int64 * 3),
    ]


class Params(ctypes.Structure):
    _fields_ = [
        ("W", ctypes.c_void_p * MAX_LAYERS), ("b", ctypes.c_void_p * MAX_LAYERS),
        ("gamma", ctypes.c_void_p * MAX_LAYERS), ("beta", ctypes.c_void_p * MAX_LAYERS),
        ("running_mean", ctypes.c_void_p * MAX_LAYERS), ("running_var", ctypes.c_void_p * MAX_LAYERS),
        ("num_batches_tracked", ctypes.c_void_p * MAX_LAYERS),
        ("Wd", ctypes.c_void_p), ("bd", ctypes.c_void_p),
    ]


class Grads(ctypes.Structure):
    _fields_ = [
        ("dW", ctypes.c_void_p * MAX_LAYERS), ("db", ctypes.c_void_p * MAX_LAYERS),
        ("dgamma", ctypes.c_void_p * MAX_LAYERS), ("dbeta", ctypes.c_void_p * MAX_LAYERS),
        ("dWd", ctypes.c_void_p), ("dbd", ctypes.c_void_p),
    ]


HEAD_MAX_LAYERS = 4
HEAD_MAX_WIDTH = 256
ACTS = {"identity": 0, "relu": 1, "sigmoid": 2, "tanh": 3}
LOSSES = {"none": 0, "mse": 1, "multitask": 2}


class HeadDesc(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int32), ("n_layers", ctypes.c_int32),
        ("width", ctypes.c_int32 * (HEAD_MAX_LAYERS + 1)), ("act", ctypes.c_int32 * HEAD_MAX_LAYERS),
        ("training", ctypes.c_int32), ("loss", ctypes.c_int32), ("need_dx", ctypes.c_int32),
        ("eps", ctypes.c_float), ("momentum", ctypes.c_float), ("ignore_index", ctypes.c_float),
    ]


class HeadParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * HEAD_MAX_LAYERS) for n in
                ("W", "b", "gamma", "beta", "running_mean", "running_var", "num_batches_tracked")]


class HeadGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * HEAD_MAX_LAYERS) for n in ("dW", "db", "dgamma", "dbeta")]


EXPORTS = [
    "ocgcn_query_workspace", "ocgcn_encoder_forward", "ocgcn_encoder_backward", "ocgcn_layer_forward",
    "ocgcn_layer_backward", "ocgcn_last_launch_count", "ocgcn_last_cuda_error", "ocgcn_status_string",
    "ocgcn_abi_version", "ocgcn_profile_enable", "ocgcn_profile_kinds", "ocgcn_profile_kind_name", "ocgcn_profile_collect",
    "ocgcn_saved_tensor", "ocgcn_expand_compact", "ocgcn_pack_compact",
    "ocgcn_head_query_workspace", "ocgcn_head_forward", "ocgcn_head_backward", "ocgcn_pool_forward", "ocgcn_pool_backward",
    "ocgcn_query_eval_workspace", "ocgcn_encoder_forward_eval",
    "ocgcn_encoder_forward_compact", "ocgcn_encoder_forward_eval_compact", "ocgcn_encoder_backward_compact",
]


def nvcc_command(out=LIB_PATH, extra=()):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
            "-Xcompiler", "-fPIC", "-shared", *extra, "-o", out, os.path.join(CSRC, "ocgcn_api.cu")]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [HEADER]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


ABI_VERSION = 5     # include/ocgcn.h OCGCN_ABI_VERSION: bump both whenever a struct layout or an entry-point signature changes


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into csrc/libocgcn.so (nvcc cross-compiles without a GPU).

    Ranks started together (torchrun / launch.py) may all find the library stale: the build is serialised with an advisory file
    lock, written to a temporary path and renamed into place, so no process can dlopen a half-written file."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():      # another process built it while we waited
                return LIB_PATH
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = nvcc_command(out=tmp, extra=("-Xptxas", "-v") if verbose else ())
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lock = threading.Lock()
_lib = None


def load():
    """Load libocgcn.so (building it if nvcc is available and it is missing/stale). Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            try:
                build()
            except Exception as exc:  # noqa: BLE001
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "openchem_b200: libocgcn.so is missing and could not be built (%s). "
                        "There is no CPU/PyTorch fallback for this path; run `python -c 'import __graft_entry__ as g; g.build()'`." % exc)
                import warnings
                warnings.warn("openchem_b200: libocgcn.so is older than its sources and the rebuild failed (%s); loading the STALE "
                              "library (its ABI version is still checked)" % (str(exc)[:200],), RuntimeWarning)
        lib = ctypes.CDLL(LIB_PATH)
        for name in EXPORTS:
            if not hasattr(lib, name):
                raise RuntimeError("libocgcn.so does not export %s" % name)
        vp = ctypes.c_void_p
        lib.ocgcn_query_workspace.argtypes = [ctypes.POINTER(Desc), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        lib.ocgcn_encoder_forward.argtypes = [ctypes.POINTER(Desc), vp, vp, ctypes.POINTER(Params), vp, vp, vp, vp]
        lib.ocgcn_encoder_backward.argtypes = [ctypes.POINTER(Desc), vp, vp, vp, ctypes.POINTER(Params), ctypes.POINTER(Grads), vp, vp, vp]
        lib.ocgcn_layer_forward.argtypes = [ctypes.POINTER(Desc), vp, vp, ctypes.POINTER(Params), vp, vp, vp, vp]
        lib.ocgcn_layer_backward.argtypes = [ctypes.POINTER(Desc), vp, vp, vp, ctypes.POINTER(Params), ctypes.POINTER(Grads), vp, vp, vp, vp]
        lib.ocgcn_saved_tensor.argtypes = [ctypes.POINTER(Desc), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        lib.ocgcn_saved_tensor.restype = ctypes.c_int
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.ocgcn_expand_compact.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, vp, vp, vp]
        lib.ocgcn_expand_compact.restype = ctypes.c_int
        lib.ocgcn_pack_compact.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, i64p, vp, i64p, vp, vp, vp, vp]
        lib.ocgcn_pack_compact.restype = ctypes.c_int
        lib.ocgcn_query_eval_workspace.argtypes = [ctypes.POINTER(Desc), ctypes.POINTER(ctypes.c_size_t)]
        lib.ocgcn_encoder_forward_eval.argtypes = [ctypes.POINTER(Desc), vp, vp, ctypes.POINTER(Params), vp, vp, vp]
        lib.ocgcn_query_eval_workspace.restype = lib.ocgcn_encoder_forward_eval.restype = ctypes.c_int
        lib.ocgcn_encoder_forward_compact.argtypes = [ctypes.POINTER(Desc), vp, vp, ctypes.POINTER(Params), vp, vp, vp, vp]
        lib.ocgcn_encoder_forward_eval_compact.argtypes = [ctypes.POINTER(Desc), vp, vp, ctypes.POINTER(Params), vp, vp, vp]
        lib.ocgcn_encoder_backward_compact.argtypes = [ctypes.POINTER(Desc), vp, ctypes.POINTER(Params), ctypes.POINTER(Grads), vp, vp, vp]
        for name in ("ocgcn_encoder_forward_compact", "ocgcn_encoder_forward_eval_compact", "ocgcn_encoder_backward_compact"):
            getattr(lib, name).restype = ctypes.c_int
        lib.ocgcn_pool_forward.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, i64p, vp, vp, vp]
        lib.ocgcn_pool_backward.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, i64p, vp, vp, vp, vp]
        lib.ocgcn_pool_forward.restype = lib.ocgcn_pool_backward.restype = ctypes.c_int
        lib.ocgcn_head_query_workspace.argtypes = [ctypes.POINTER(HeadDesc), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        lib.ocgcn_head_forward.argtypes = [ctypes.POINTER(HeadDesc), vp, ctypes.POINTER(HeadParams), vp, vp, vp, vp, vp, vp]
        lib.ocgcn_head_backward.argtypes = [ctypes.POINTER(HeadDesc), vp, vp, vp, vp, ctypes.POINTER(HeadParams), ctypes.POINTER(HeadGrads),
                                            vp, vp, vp, vp]
        for name in ("ocgcn_head_query_workspace", "ocgcn_head_forward", "ocgcn_head_backward"):
            getattr(lib, name).restype = ctypes.c_int
        lib.ocgcn_profile_kind_name.restype = ctypes.c_char_p
        lib.ocgcn_profile_kind_name.argtypes = [ctypes.c_int]
        lib.ocgcn_profile_enable.argtypes = [ctypes.c_int]
        lib.ocgcn_profile_collect.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        lib.ocgcn_status_string.restype = ctypes.c_char_p
        lib.ocgcn_status_string.argtypes = [ctypes.c_int]
        for name in ("ocgcn_query_workspace", "ocgcn_encoder_forward", "ocgcn_encoder_backward", "ocgcn_layer_forward",
                     "ocgcn_layer_backward", "ocgcn_last_launch_count", "ocgcn_last_cuda_error", "ocgcn_abi_version"):
            getattr(lib, name).restype = ctypes.c_int
        if lib.ocgcn_abi_version() != ABI_VERSION:
            raise RuntimeError("libocgcn.so reports ABI version %d, this binding needs %d; rebuild it "
                               "(python -c 'import __graft_entry__ as g; g.build()')" % (lib.ocgcn_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        lib = load()
        msg = lib.ocgcn_status_string(status).decode()
        if status == 5:
            msg += " [cudaError %d]" % lib.ocgcn_last_cuda_error()
        raise RuntimeError(msg)


def profile_enable(on):
    load().ocgcn_profile_enable(int(bool(on)))


def profile_collect():
    """-> {kind_name: (total_ms, launches)} for the launches recorded since the last collect (synchronises)."""
    lib = load()
    n = lib.ocgcn_profile_kinds()
    ms = (ctypes.c_float * n)()
    cnt = (ctypes.c_int * n)()
    check(lib.ocgcn_profile_collect(ms, cnt, n))
    return {lib.ocgcn_profile_kind_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}
