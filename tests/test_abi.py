"""The C-ABI library loads and exports every symbol include/colvo.h declares; host-only entry
points behave (no compute calls here: this container has no GPU)."""
import ctypes
import os
import re

import pytest

from coivo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "colvo.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(colvo_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_what_the_binding_expects():
    assert set(declared_functions()) == set(_lib.EXPORTS)


def test_library_builds_loads_and_exports_every_symbol():
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.colvo_version() == 100


def test_desc_init_and_validation():
    lib = _lib.load()
    d = _lib.make_desc(12, 2, 4, 256, 320, _lib.F_LCC)
    assert list(d.h) == [256, 128, 64, 32] and list(d.w) == [320, 160, 80, 40]
    assert abs(d.alpha - 0.85) < 1e-7 and abs(d.c2 - 9e-4) < 1e-9
    for bad in [(0, 2, 4, 256, 320), (1, 3, 4, 256, 320), (1, 2, 5, 256, 320), (1, 2, 4, 1, 320), (1, 2, 4, 4, 320)]:
        with pytest.raises(ValueError):
            _lib.make_desc(*bad, 0)
    d.h[1] = 127
    n = ctypes.c_size_t()
    assert lib.colvo_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) == -1
    assert b"descriptor" in lib.colvo_error_string(-1)
    assert lib.colvo_error_string(0) == b"success"
    assert b"invalid" in lib.colvo_error_string(1).lower()       # cudaErrorInvalidValue


def test_workspace_and_saved_sizes():
    lib = _lib.load()
    d = _lib.make_desc(12, 2, 4, 256, 320, _lib.F_LCC)
    n, m, a = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.colvo_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) == 0
    assert lib.colvo_saved_doubles(ctypes.byref(d), ctypes.byref(m)) == 0
    assert lib.colvo_step_host_arena_bytes(ctypes.byref(d), ctypes.byref(a)) == 0
    # doubles: 8 per warped frame + 2 per (b, k); then fp32 (2 per double): the smoothness adjoint fields
    # [B,h_k,w_k], the SSIM adjoint coefficients [B,S,3,H,W] and the projections [B,N,S,H,W] as float4 texels
    nf = 12 * (256 * 320 + 128 * 160 + 64 * 80 + 32 * 40) + 12 * 4 * 12 * 256 * 320 + 12 * 2 * 4 * 4 * 256 * 320
    assert m.value == 12 * 2 * 4 * 8 + 12 * 4 * 2 + (nf + 1) // 2
    iw = 4 * 12 * 2 * 4 * 4 * 256 * 320                      # cached raw warped frames as 16-byte texels (stats -> tile kernel)
    assert iw < n.value < iw + (32 << 20)
    assert a.value > n.value + 4 * 12 * (3 + 6 + 6) * 256 * 320
    assert lib.colvo_workspace_bytes(ctypes.byref(d), None) == -3
    c = ctypes.c_size_t()
    assert lib.colvo_consistency_workspace_bytes(2000, 256, 320, ctypes.byref(c)) == 0 and c.value > 0
    assert lib.colvo_consistency_workspace_bytes(1, 256, 320, ctypes.byref(c)) == -1


def test_null_pointer_and_alignment_checks_without_touching_the_gpu():
    lib = _lib.load()
    d = _lib.make_desc(1, 2, 2, 16, 24, _lib.F_LCC)
    depth = _lib.ptr_array([None, None])
    rc = lib.colvo_photo_forward(ctypes.byref(d), None, None, depth, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -3
    assert lib.colvo_debug_time_kernel(7, None, None) == -5
    assert lib.colvo_debug_time_kernel(0, None, None) == 0


def test_step_host_arena_gradient_offsets():
    """colvo_step_host_arena_grads: byte offsets of the device-resident gradients inside the caller's arena --
    16-byte aligned, disjoint, inside colvo_step_host_arena_bytes."""
    lib = _lib.load()
    d = _lib.make_desc(3, 2, 4, 64, 96, _lib.F_LCC)
    total = ctypes.c_size_t()
    assert lib.colvo_step_host_arena_bytes(ctypes.byref(d), ctypes.byref(total)) == 0
    offs = (ctypes.c_size_t * _lib.MAX_SCALES)()
    oT, oS = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.colvo_step_host_arena_grads(ctypes.byref(d), offs, ctypes.byref(oT), ctypes.byref(oS)) == 0
    spans = [(offs[k], 4 * 3 * (64 >> k) * (96 >> k)) for k in range(4)] + [(oT.value, 4 * 3 * 2 * 16), (oS.value, 4 * 3 * 2 * 3 * 64 * 96)]
    for off, n in spans:
        assert off % 16 == 0 and off + n <= total.value
    spans.sort()
    for (a, n), (b, _) in zip(spans, spans[1:]):
        assert a + n <= b
    assert lib.colvo_step_host_arena_grads(ctypes.byref(d), None, ctypes.byref(oT), ctypes.byref(oS)) == -3
