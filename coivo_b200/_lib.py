"""ctypes binding of libcolvo_b200.so (the C ABI in include/colvo.h) and its in-tree build.

The library is the product: if it is missing or fails to load, importing callers fail loudly
(`ColvoLibraryError`).  There is no CPU or PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from typing import List, Optional

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
# COLVO_LIB selects another build of the same sources (tuning experiments); the default is the in-tree library
LIB_PATH = os.environ.get("COLVO_LIB") or os.path.join(PKG_DIR, "libcolvo_b200.so")
SOURCES = ["colvo_fwd.cu", "colvo_bwd.cu", "colvo_api.cu", "colvo_front.cu"]
HEADERS = ["colvo_math.cuh", "colvo_kernels.cuh", "colvo_photo_fwd.cuh", "colvo_f2.cuh", "colvo_pe.cuh",
           os.path.join("..", "..", "include", "colvo.h")]

# flags (include/colvo.h)
F_LCC = 1
F_LCC_DETACH = 2
F_SAVE_FOR_BWD = 4
F_NO_SRC_GRAD = 8
F_PACKED_BF16 = 16
F_HOST_U8 = 32
F_SCATTER_MERGE = 64

MAX_SCALES = 4
MAX_SOURCES = 2


class ColvoLibraryError(RuntimeError):
    pass


class ColvoDesc(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int32), ("N", ctypes.c_int32), ("S", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
        ("h", ctypes.c_int32 * MAX_SCALES), ("w", ctypes.c_int32 * MAX_SCALES),
        ("alpha", ctypes.c_float), ("c1", ctypes.c_float), ("c2", ctypes.c_float), ("eps_proj", ctypes.c_float),
        ("eps_lcc", ctypes.c_float), ("eps_disp", ctypes.c_float), ("z_min", ctypes.c_float),
        ("smooth_weight", ctypes.c_float), ("geo_weight", ctypes.c_float), ("flags", ctypes.c_uint32),
    ]


def nvcc_command(out: str = LIB_PATH, defines=()) -> List[str]:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    return [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o", out,
    ] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]


def source_hash() -> str:
    """sha256 (first 16 hex digits) over the kernel sources and headers: ties a profile record (profiles/traffic.json) to
    the code it was measured on."""
    import hashlib
    h = hashlib.sha256()
    for rel in sorted(SOURCES + HEADERS):
        with open(os.path.normpath(os.path.join(CSRC, rel)), "rb") as f:
            h.update(rel.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def needs_build() -> bool:
    if os.environ.get("COLVO_LIB"):
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA kernels + C ABI for sm_100a into coivo_b200/libcolvo_b200.so (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = nvcc_command(tmp)
    if verbose:
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise ColvoLibraryError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lock = threading.Lock()
_lib: Optional[ctypes.CDLL] = None

_vp = ctypes.c_void_p
_SIGS = {
    "colvo_version": (ctypes.c_int, []),
    "colvo_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "colvo_desc_init": (ctypes.c_int, [ctypes.POINTER(ColvoDesc)] + [ctypes.c_int32] * 5 + [ctypes.c_uint32]),
    "colvo_workspace_bytes": (ctypes.c_int, [ctypes.POINTER(ColvoDesc), ctypes.POINTER(ctypes.c_size_t)]),
    "colvo_saved_doubles": (ctypes.c_int, [ctypes.POINTER(ColvoDesc), ctypes.POINTER(ctypes.c_size_t)]),
    "colvo_photo_forward": (ctypes.c_int, [ctypes.POINTER(ColvoDesc), _vp, _vp, ctypes.POINTER(_vp), _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "colvo_photo_forward_occ": (ctypes.c_int, [ctypes.POINTER(ColvoDesc), _vp, _vp, ctypes.POINTER(_vp), _vp, _vp, _vp, _vp, _vp,
                                               _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "colvo_photo_backward": (ctypes.c_int, [ctypes.POINTER(ColvoDesc), _vp, _vp, ctypes.POINTER(_vp), _vp, _vp, _vp, _vp, _vp,
                                            _vp, ctypes.POINTER(_vp), _vp, _vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "colvo_consistency_workspace_bytes": (ctypes.c_int, [ctypes.c_int32] * 3 + [ctypes.POINTER(ctypes.c_size_t)]),
    "colvo_consistency": (ctypes.c_int, [ctypes.c_int32] * 3 + [ctypes.c_uint32, _vp, _vp, _vp, _vp, ctypes.c_int32,
                                                                 _vp, _vp, ctypes.c_size_t, _vp]),
    "colvo_step_host_arena_bytes": (ctypes.c_int, [ctypes.POINTER(ColvoDesc), ctypes.POINTER(ctypes.c_size_t)]),
    "colvo_step_host_arena_grads": (ctypes.c_int, [ctypes.POINTER(ColvoDesc), ctypes.POINTER(ctypes.c_size_t),
                                                   ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "colvo_photo_step_host": (ctypes.c_int, [ctypes.POINTER(ColvoDesc), _vp, _vp, ctypes.POINTER(_vp), _vp, _vp, _vp,
                                             ctypes.POINTER(_vp), _vp, _vp, ctypes.c_float, _vp, ctypes.c_size_t, _vp]),
    "colvo_debug_time_kernel": (ctypes.c_int, [ctypes.c_int, _vp, _vp]),
    "colvo_pose_from_axisangle": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32, _vp, _vp, _vp, _vp]),
    "colvo_pose_from_axisangle_backward": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32, _vp, _vp, _vp, _vp,
                                                          _vp, _vp]),
    "colvo_disp_to_depth": (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_vp),
                                           ctypes.POINTER(_vp), ctypes.c_float, ctypes.c_float, _vp]),
    "colvo_disp_to_depth_backward": (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_vp),
                                                    ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.c_float,
                                                    ctypes.c_float, _vp]),
}
EXPORTS = tuple(_SIGS)

K_PHOTO_FWD, K_PHOTO_BWD, K_WARP_STATS, K_CONSISTENCY_PE = 1, 2, 3, 4
# kernels launched by one forward + one backward (S > 1, LCC on); bench.py's gpu_launches
KERNELS_FWD = ("k_warp_stats", "k_lcc_solve", "k_smooth", "k_photo_fwd", "k_smooth", "k_finalize_fwd")   # (k_smooth: two half-batch launches)
KERNELS_BWD = ("k_zero", "k_photo_bwd", "k_depth_gather")   # (k_depth_gather's launch carries the pose reduction and the unpack)


def load(auto_build: bool = True) -> ctypes.CDLL:
    """Load (building first if the sources are newer) and type the C ABI.  Raises if impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if auto_build and needs_build():
            build()
        if not os.path.exists(LIB_PATH):
            raise ColvoLibraryError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise ColvoLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise ColvoLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def error_string(rc: int) -> str:
    return load().colvo_error_string(rc).decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {error_string(rc)} (rc={rc})")


def make_desc(B: int, N: int, S: int, H: int, W: int, flags: int, alpha: float = 0.85,
              smooth_weight: float = 1e-3, geo_weight: float = 0.0) -> ColvoDesc:
    d = ColvoDesc()
    rc = load().colvo_desc_init(ctypes.byref(d), B, N, S, H, W, flags)
    if rc != 0:
        raise ValueError(f"bad problem size B={B} N={N} S={S} H={H} W={W}: {error_string(rc)}")
    d.alpha = alpha
    d.smooth_weight = smooth_weight
    d.geo_weight = geo_weight
    return d


def ptr_array(ptrs, n: int = MAX_SCALES):
    arr = (_vp * n)()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
