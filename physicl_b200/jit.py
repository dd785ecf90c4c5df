"""Run-time compiled kernels: the stand-in for ``cl.Program(ctx, src).build()``.

The reference assembles its OpenCL kernels from Python strings and builds them on first use
(``CLProgram.build_kernel``, physicl/__init__.py:583-597).  Two of its features exist only as such
run-time text: the number-density expression of ``ScatterIsotropicStep(variable_n=True)``
(light.py:295-299) and user-written ``CLProgram`` kernels.  Here the text is CUDA C++ compiled by NVRTC
for sm_100a inside the C library (``csrc/jit.cu``); this module assembles the translation units and
owns the module handles.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

# NVRTC has no system headers: the few fixed-width names the sources use
_STDINT = """#pragma once
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t; typedef unsigned long size_t;
"""

# OpenCL-C built-ins that CUDA C++ spells differently (double precision, as the reference computes)
OPENCL_COMPAT = """
__device__ __forceinline__ double powr(double x, double y) { return pow(x, y); }
__device__ __forceinline__ double pown(double x, int y) { return pow(x, y); }
__device__ __forceinline__ double mad(double a, double b, double c) { return a * b + c; }
__device__ __forceinline__ double clamp(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
__device__ __forceinline__ double mix(double a, double b, double t) { return a + (b - a) * t; }
__device__ __forceinline__ double native_exp(double x) { return exp(x); }
__device__ __forceinline__ double native_log(double x) { return log(x); }
__device__ __forceinline__ double native_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double native_powr(double x, double y) { return pow(x, y); }
__device__ __forceinline__ double native_sin(double x) { return sin(x); }
__device__ __forceinline__ double native_cos(double x) { return cos(x); }
__device__ __forceinline__ double half_exp(double x) { return exp(x); }
__device__ __forceinline__ double half_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double radians(double d) { return d * 0.017453292519943295; }
__device__ __forceinline__ double degrees(double r) { return r * 57.29577951308232; }
"""

_headers = None


def headers():
    """(names, texts) of the in-memory include files handed to NVRTC."""
    global _headers
    if _headers is None:
        files = [("physicl_b200.h", os.path.join(_INCLUDE, "physicl_b200.h")),
                 ("pcl_device.cuh", os.path.join(_CSRC, "pcl_device.cuh")),
                 ("pcl_photon_body.cuh", os.path.join(_CSRC, "pcl_photon_body.cuh")),
                 ("pcl_jit_photon.cuh", os.path.join(_CSRC, "pcl_jit_photon.cuh"))]
        names, texts = ["stdint.h", "pcl_opencl_compat.cuh"], [_STDINT, "#pragma once\n" + OPENCL_COMPAT]
        for name, path in files:
            with open(path) as f:
                names.append(name)
                texts.append(f.read())
        _headers = (names, texts)
    return _headers


def _c_headers():
    names, texts = headers()
    n = len(names)
    return n, (C.c_char_p * n)(*[x.encode() for x in names]), (C.c_char_p * n)(*[x.encode() for x in texts])


def check(source: str) -> int:
    """Compile ``source`` without loading it (needs no GPU).  Returns the cubin size; raises
    ``PclError`` with the compiler log otherwise."""
    lib = _capi.load()
    n, names, texts = _c_headers()
    log = C.create_string_buffer(1 << 16)
    size = C.c_uint64(0)
    rc = lib.pcl_jit_check(source.encode(), n, names, texts, log, C.c_uint64(len(log)), C.byref(size))
    if rc != 0:
        raise _capi.PclError("run-time kernel does not compile (%d): %s" % (rc, log.value.decode(errors="replace")))
    return int(size.value)


class Module:
    """A built module on one context; kernels are looked up by name and cached."""

    def __init__(self, ctx, source: str):
        self.ctx = ctx
        self.source = source
        n, names, texts = _c_headers()
        h = C.c_void_p()
        ctx.call("pcl_jit_build", source.encode(), n, names, texts, C.byref(h))
        self.handle = h
        self._kernels = {}

    def kernel(self, name: str):
        k = self._kernels.get(name)
        if k is None:
            k = C.c_void_p()
            self.ctx.call("pcl_jit_get_kernel", self.handle, name.encode(), C.byref(k))
            self._kernels[name] = k
        return k

    def launch(self, name: str, stream, n: int, *args):
        """``prog.<name>(queue, (n,), None, *args)``: args are ctypes values (pointers as c_void_p)."""
        ptrs = (C.c_void_p * len(args))(*[C.cast(C.pointer(a), C.c_void_p) for a in args])
        self.ctx.call("pcl_jit_launch", stream, self.kernel(name), C.c_uint64(n), ptrs)

    def close(self):
        if getattr(self, "handle", None) and getattr(self.ctx, "handle", None):
            self.ctx.lib.pcl_jit_free(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def photon_source(expr: str, wavelength: bool, delete: bool = False) -> str:
    """Translation unit of the variable-density photon kernels: the user's expression for n(r)
    (OpenCL-C over ``r0[gid]``, ``r1[gid]``, ``r2[gid]``, ``E[gid]``, ``norm``, ``A``, ``n``; light.py:295-299)
    spliced into ``csrc/pcl_jit_photon.cuh``."""
    if not isinstance(expr, str) or not expr.strip():
        raise ValueError("variable_n_fn must be a non-empty expression string (light.py:299)")
    if "\n#" in "\n" + expr.replace("\\\n", " "):
        raise ValueError("variable_n_fn must be an expression, not preprocessor text")
    one_line = " ".join(expr.split())
    return ('#include "stdint.h"\n#include "pcl_opencl_compat.cuh"\n#define PCL_JIT_WAVE %d\n#define PCL_JIT_DEL %d\n'
            '#define PCL_USER_N_EXPR %s\n#include "pcl_jit_photon.cuh"\n' % (int(bool(wavelength)), int(bool(delete)), one_line))
