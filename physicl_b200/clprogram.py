"""User-written kernels: ``CLInput`` / ``CLOutput`` / ``CLProgram`` (physicl/__init__.py:543-664).

The reference lets a step describe a kernel declaratively: per-object inputs gathered from
``sim.objects`` by generated Python, scalar constants, outputs, and an OpenCL-C body that is wrapped
into ``__kernel void name(args){body}`` and built with pyopencl at first use.  This module keeps that
interface (same classes, same attributes, same generated argument order, same dict of NumPy arrays
back from ``run``) and retargets it: the body is compiled for sm_100a by NVRTC inside the C library
(``csrc/jit.cu``) and launched on the simulation's stream.

What is translated, not changed: the body stays OpenCL-C as the user wrote it.  ``get_global_id(0)``
is the index of a grid-stride loop, ``__global`` / ``__kernel`` / ``__constant`` are accepted,
``NAN`` / ``INFINITY`` / ``M_PI`` exist, OpenCL spellings such as ``pown`` or ``native_exp`` map to the
CUDA double-precision functions, and ``return`` leaves the work item.  Arithmetic is float64 unless
the user declares another ``ctype``, as in the reference.

Differences (deliberate): inputs are converted to their declared ``ctype`` (the reference always
uploads ``np.double`` whatever the kernel signature says, physicl/__init__.py:613); the output
dtype ``int`` means int32, matching the kernel's ``int *`` (the reference writes ``np.int``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

# OpenCL-C scalar types -> (CUDA C++ spelling, NumPy dtype, ctypes type)
_CTYPES = {
    "double": ("double", np.float64, C.c_double),
    "float": ("float", np.float32, C.c_float),
    "int": ("int", np.int32, C.c_int32),
    "uint": ("unsigned int", np.uint32, C.c_uint32),
    "unsigned int": ("unsigned int", np.uint32, C.c_uint32),
    "long": ("long long", np.int64, C.c_int64),
    "ulong": ("unsigned long long", np.uint64, C.c_uint64),
    "unsigned long": ("unsigned long long", np.uint64, C.c_uint64),
    "short": ("short", np.int16, C.c_int16),
    "ushort": ("unsigned short", np.uint16, C.c_uint16),
    "char": ("signed char", np.int8, C.c_int8),
    "uchar": ("unsigned char", np.uint8, C.c_uint8),
}

_PRELUDE = """#include "stdint.h"
#include "pcl_opencl_compat.cuh"
#define __global
#define __kernel
#define __constant const
#define __private
#define __local
#define get_global_id(d) ((int)pcl_gid)
#define get_global_size(d) ((int)pcl_n)
#ifndef NAN
#define NAN (__longlong_as_double(0x7ff8000000000000LL))
#endif
#ifndef INFINITY
#define INFINITY (__longlong_as_double(0x7ff0000000000000LL))
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
typedef unsigned int uint;
typedef unsigned long long ulong;
typedef unsigned short ushort;
typedef unsigned char uchar;
"""


def _ctype(name):
    try:
        return _CTYPES[name.strip()]
    except KeyError:
        raise ValueError("unsupported kernel argument type %r (known: %s)" % (name, ", ".join(sorted(_CTYPES)))) from None


class CLInput:
    """physicl/__init__.py:543-560.  ``type``: ``obj`` (``obj_attr`` of every object -> array), ``obj_def``
    (a Python expression evaluated once per object -> array), ``obj_track`` (keeps the objects themselves
    in ``program.<name>``), ``obj_action`` / ``other`` (raw Python run inside / after the gather loop),
    ``const`` (a scalar, ``const_value`` is its source text)."""

    types = ["obj", "obj_def", "obj_action", "const", "other"]

    def __init__(self, **kwargs):
        self.name = kwargs["name"]
        self.type = kwargs["type"]
        if kwargs["type"] == "obj":
            self.code = "self." + self.name + ".append(obj." + kwargs["obj_attr"] + ")"
            self.ctype = "double" if "ctype" not in kwargs else kwargs["ctype"]
        elif kwargs["type"] == "obj_def":
            self.code = "self." + self.name + ".append(" + kwargs["obj_def"] + ")"
            self.ctype = "double" if "ctype" not in kwargs else kwargs["ctype"]
        elif kwargs["type"] == "obj_track":
            self.code = "self." + self.name + ".append(" + kwargs["obj_track"] + ")"
        elif kwargs["type"] in ["obj_action", "other"]:
            self.code = kwargs["code"]
        elif kwargs["type"] == "const":
            self.const_value = kwargs["const_value"]
            self.ctype = "double" if "ctype" not in kwargs else kwargs["ctype"]


class CLOutput:
    """physicl/__init__.py:562-565"""

    def __init__(self, **kwargs):
        self.name = kwargs["name"]
        self.ctype = kwargs["ctype"] if "ctype" in kwargs else "double"


class CLProgram:
    """physicl/__init__.py:567-664: ``prep_metadata`` (inputs), ``output_metadata`` (outputs),
    ``build_kernel()`` once, then ``run()`` per timestep -> ``{output name: ndarray}``."""

    def __init__(self, sim, name, kernel_code):
        self.variables = {}
        self.sim = sim
        self.prog = None
        self.prog_name = name
        self.prep_metadata = []
        self.output_metadata = []
        self.kernel_code = kernel_code
        self.source = None

    # ---- text ------------------------------------------------------------------------------------
    def _args(self):
        """Kernel arguments in the reference's order (physicl/__init__.py:586-592): obj / obj_def arrays
        and constants in ``prep_metadata`` order, then the outputs."""
        ins = [it for it in self.prep_metadata if it.type in ("obj", "obj_def", "const")]
        return ins, list(self.output_metadata)

    def kernel_source(self):
        """The translation unit handed to NVRTC: the user's body inside a grid-stride loop."""
        ins, outs = self._args()
        cat = ["unsigned long long pcl_n"]
        for it in ins:
            ct = _ctype(it.ctype)[0]
            cat.append((ct + " *" + it.name) if it.type != "const" else (ct + " " + it.name))
        cat.extend(_ctype(o.ctype)[0] + " *" + o.name for o in outs)
        return (_PRELUDE + 'extern "C" __global__ void ' + self.prog_name + "(" + ", ".join(cat) + ") {\n"
                "    for (unsigned long long pcl_gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; pcl_gid < pcl_n;\n"
                "         pcl_gid += (unsigned long long)gridDim.x * blockDim.x) {\n"
                "        [&]() {\n" + self.kernel_code + "\n        }();\n    }\n}\n")

    def build_kernel(self):
        """physicl/__init__.py:583-597.  Compiles now (NVRTC needs no device), loads on first ``run``."""
        from . import jit

        self.source = self.kernel_source()
        jit.check(self.source)
        self.prog = None

    # ---- gather + launch ---------------------------------------------------------------------------
    def _gather(self):
        """The reference's generated gather loop (physicl/__init__.py:606-629), executed as it stands:
        one pass over ``sim.objects`` running every item's code in ``prep_metadata`` order."""
        import physicl_b200
        import physicl_b200.light  # noqa: F401  (user code says physicl.light.PhotonObject)

        # the reference executes the generated text inside its own module (physicl/__init__.py:631-635), so user code
        # sees that module's names
        env = {"physicl": physicl_b200, "physicl_b200": physicl_b200, "np": np, "self": self,
               "Measurement": physicl_b200.Measurement, "Object": physicl_b200.Object}
        initial, loop, other = "", "for obj in self.sim.objects:\n\tpass", ""
        for item in self.prep_metadata:
            if item.type in ("obj", "obj_def", "obj_track"):
                initial += "self." + item.name + " = []\n"
            if item.type in ("obj", "obj_def", "obj_action", "obj_track"):
                loop += "\n\t" + item.code
            elif item.type == "other":
                other += item.code + "\n"
        exec(initial, env)
        exec(loop, env)
        if other:
            exec(other, env)
        for item in self.prep_metadata:
            if item.type in ("obj", "obj_def"):
                vals = [float(x) for x in getattr(self, item.name)]
                setattr(self, item.name + "_np", np.array(vals, dtype=_ctype(item.ctype)[1]))

    def run(self):
        """physicl/__init__.py:602-664: gather, upload, launch over the length of the first ``obj`` input,
        download every output."""
        import torch

        from . import jit

        if self.source is None:
            self.build_kernel()
        sim = self.sim
        if not getattr(sim, "cl_on", False):
            raise RuntimeError("physicl_b200 has no CPU path: CLProgram needs Simulation(cl_on=True)")
        ctx = sim.cl_ctx
        if self.prog is None or self.prog.ctx is not ctx:
            self.prog = jit.Module(ctx, self.source)
        self._gather()
        n = None
        for item in self.prep_metadata:
            if item.type == "obj":  # physicl/__init__.py:640-644: the first obj input defines the global size
                n = int(getattr(self, item.name + "_np").shape[0])
                break
        if n is None:
            raise ValueError("CLProgram.run: no input of type 'obj', so the launch size is undefined (physicl/__init__.py:640)")
        dev = torch.device("cuda", ctx.device)
        stream = torch.cuda.current_stream(dev)
        ins, outs = self._args()
        args, keep = [C.c_uint64(n)], []
        with torch.cuda.device(dev):
            for it in ins:
                if it.type == "const":
                    # the reference pastes const_value into the call as np.double(<text>) (physicl/__init__.py:648)
                    val = eval(str(it.const_value), {"np": np})
                    npdt, cty = _ctype(it.ctype)[1], _ctype(it.ctype)[2]
                    args.append(cty(float(val)) if np.issubdtype(npdt, np.floating) else cty(int(val)))
                    continue
                arr = getattr(self, it.name + "_np")
                if arr.shape[0] != n:
                    raise ValueError("input %r has %d entries, the launch covers %d" % (it.name, arr.shape[0], n))
                t = torch.from_numpy(arr).to(dev, non_blocking=False) if n else torch.empty(0, device=dev)
                setattr(self, it.name + "_dev", t)
                keep.append(t)
                args.append(C.c_void_p(t.data_ptr() if n else 0))
            res = {}
            for o in outs:
                npdt = _ctype(o.ctype)[1]
                t = torch.empty(max(n, 1), dtype=getattr(torch, np.dtype(npdt).name), device=dev)
                setattr(self, "res_" + o.name, t)
                res[o.name] = t
                args.append(C.c_void_p(t.data_ptr()))
            if n:
                self.prog.launch(self.prog_name, C.c_void_p(stream.cuda_stream), n, *args)
            out = {name: t[:n].cpu().numpy() for name, t in res.items()}
        return out
