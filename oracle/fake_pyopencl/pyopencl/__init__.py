"""Stand-in for ``pyopencl`` so the UNMODIFIED reference (``/root/reference/physicl``) can run in a
container with no OpenCL runtime.  TEST INFRASTRUCTURE (golden-vector generation only).

It implements exactly the entry points the reference calls:
``create_some_context`` (physicl/__init__.py:428), ``CommandQueue`` (:429), ``Program(ctx, src).build()``
(:597, light.py:160), ``prog.<kernel>(queue, global, local, *args)`` (:656, light.py:199),
``array.to_device`` (:614), ``array.empty`` (:653), ``Array.get`` (:662), ``get_platforms`` (:476).

``Program.build`` compiles the reference's own OpenCL-C kernel text as C99 with gcc: ``__kernel`` and
``__global`` become empty macros, ``get_global_id(0)`` reads a thread-local loop index, and a generated
``<name>__launch`` wrapper loops the work-items.  Arithmetic is the host libm in double precision,
i.e. the same IEEE operations an OpenCL CPU device would perform.

Every launch is appended to ``LAUNCH_LOG`` (inputs before, outputs after) for the golden recorder.
"""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np

LAUNCH_LOG = []
RECORD = False
_CACHE_DIR = os.path.join(tempfile.gettempdir(), "fake_pyopencl_cache")

_PRELUDE = r"""
#include <math.h>
#define __kernel
#define __global
static __thread long fcl_gid;
#define get_global_id(d) (fcl_gid)
"""


class Context:
    pass


class CommandQueue:
    def __init__(self, ctx, *a, **k):
        self.context = ctx


def create_some_context(*a, **k):
    return Context()


def get_platforms():
    return []


class platform_info:
    NAME = 0


class device_info:
    NAME = 0


class _Kernel:
    def __init__(self, name, fn, argspec):
        self.name, self.fn, self.argspec = name, fn, argspec

    def __call__(self, queue, global_shape, local_shape, *args):
        n = int(np.prod(global_shape))
        assert len(args) == len(self.argspec), (self.name, len(args), len(self.argspec))
        cargs, arrays = [], []
        for (ctype, is_ptr, aname), a in zip(self.argspec, args):
            if is_ptr:
                assert isinstance(a, np.ndarray), (aname, type(a))
                arrays.append((aname, a))
                cargs.append(a.ctypes.data_as(ctypes.c_void_p))
            elif ctype == "double":
                cargs.append(ctypes.c_double(float(a)))
            elif ctype == "int":
                cargs.append(ctypes.c_int(int(a)))
            else:
                raise TypeError(ctype)
        before = {k: v.copy() for k, v in arrays} if RECORD else None
        scalars = {aname: float(a) for (ctype, is_ptr, aname), a in zip(self.argspec, args) if not is_ptr}
        self.fn(ctypes.c_long(n), *cargs)
        if RECORD:
            LAUNCH_LOG.append({"kernel": self.name, "n": n, "before": before, "scalars": scalars,
                               "after": {k: v.copy() for k, v in arrays}})


class Program:
    def __init__(self, ctx, src):
        self.src = src
        self._kernels = {}

    def build(self, *a, **k):
        os.makedirs(_CACHE_DIR, exist_ok=True)
        wrappers = []
        specs = {}
        for m in re.finditer(r"__kernel\s+void\s+(\w+)\s*\(([^)]*)\)", self.src):
            name, arglist = m.group(1), m.group(2)
            spec = []
            for a in arglist.split(","):
                a = a.replace("__global", "").strip()
                is_ptr = "*" in a
                toks = a.replace("*", " ").split()
                spec.append((toks[0], is_ptr, toks[-1]))
            specs[name] = spec
            decl = ", ".join(("%s *%s" if p else "%s %s") % (t, n) for t, p, n in spec)
            call = ", ".join(n for _, _, n in spec)
            wrappers.append("void %s__launch(long fcl_n, %s){\n#pragma omp parallel for\nfor (long i = 0; i < fcl_n; ++i)"
                            "{ fcl_gid = i; %s(%s); }\n}\n" % (name, decl, name, call))
        csrc = _PRELUDE + self.src + "\n" + "\n".join(wrappers)
        tag = hashlib.sha1(csrc.encode()).hexdigest()[:16]
        so = os.path.join(_CACHE_DIR, "k_%s.so" % tag)
        if not os.path.exists(so):
            cfile = os.path.join(_CACHE_DIR, "k_%s.c" % tag)
            with open(cfile, "w") as f:
                f.write(csrc)
            subprocess.run(["/usr/bin/gcc", "-O2", "-std=gnu99", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off",
                            "-o", so, cfile, "-lm"], check=True)
        lib = ctypes.CDLL(so)
        for name, spec in specs.items():
            self._kernels[name] = _Kernel(name, getattr(lib, name + "__launch"), spec)
        return self

    def __getattr__(self, name):
        ks = self.__dict__.get("_kernels", {})
        if name in ks:
            return ks[name]
        raise AttributeError(name)
