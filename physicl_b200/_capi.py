"""ctypes binding of ``include/physicl_b200.h`` (the C ABI of ``libphysicl_b200.so``).

This is the only place Python touches the device library.  There is NO CPU fallback: if the
shared library is missing, or no sm_100 device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PHYSICL_B200_LIB: kernel-tuning aid, points at an alternative build of the same sources
LIB_PATH = os.environ.get("PHYSICL_B200_LIB") or os.path.join(_HERE, "libphysicl_b200.so")

MAX_PLANES = 8
TALLY_COLS = 16
T_ALIVE, T_XP, T_YP, T_ZP, T_SCATTERED, T_ABSORBED, T_ESCAPED, T_LIVE_IN, T_PLANE0 = range(9)
SCATTER_WAVELENGTH, SCATTER_DELETE, SCATTER_SFU = 1, 2, 4

# Thread-level SASS instructions of the shipped fused photon kernels, from ncu (profiles/r2/ncu_full_photon_multi_v3.csv,
# ncu_full_photon_wave_inplace_v3.csv; static loop counts by scripts/sass_loop.py agree): per photon-step inside the timestep
# loop, and per photon and LAUNCH for the load + write-back / compaction phase.  bench.py turns them into the issue-rate
# roofline of the photon workloads: instructions per photon-step = LOOP + LAUNCH / (timesteps per launch).
PHOTON_INSTR_LOOP = 88.7        # compacting launches of 3 and 4 timesteps: (226.90 M - 180.37 M) warp instructions x 32 / 16 Mi photons
PHOTON_INSTR_LAUNCH = 77.8      # 180.37 M x 32 / 16 Mi - 3 x 88.7
PHOTON_INSTR_LOOP_WAVE = 86.2   # wavelength law, in place: static count of the loop (345 per four photons)
PHOTON_INSTR_LAUNCH_WAVE = 39.6  # 1529.4 M x 32 / (8 x 64 Mi) = 91.15 per photon-step at 8 timesteps per launch
GRAVITY_KERNEL = "pcl_k_gravity_x2<2,128,512,UM,8> (2 i-bodies per thread, 512-body j-tiles, packed FP32; UM = equal masses)"

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)


class PclError(RuntimeError):
    """A non-zero return from the C ABI; the message is ``pcl_last_error``."""


class Soa(C.Structure):
    _fields_ = [("n", C.c_uint64)] + [(k, C.c_void_p) for k in (
        "x", "y", "z", "vx", "vy", "vz", "dx", "dy", "dz", "ax", "ay", "az", "e", "id", "nscat")] + [("id_base", C.c_uint64),
                                                                                                  ("n_dev", C.c_void_p)]


class Pingpong(C.Structure):
    _fields_ = [("buf", Soa * 2), ("n_dev", C.c_void_p), ("cur", C.c_uint32), ("id_valid", C.c_uint32)]


class ScatterParams(C.Structure):
    _fields_ = [("k", C.c_float), ("c", C.c_float), ("mode", C.c_uint32), ("_pad", C.c_uint32)]


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("step", C.c_uint32), ("_pad", C.c_uint32),
                ("u_theta", C.c_void_p), ("u_phi", C.c_void_p), ("u_rand", C.c_void_p)]


class VarnParams(C.Structure):
    _fields_ = [("kd", C.c_double), ("e0", C.c_double), ("a_slot", C.c_double), ("n_slot", C.c_double)]


class Planes(C.Structure):
    _fields_ = [("count", C.c_uint32), ("axis", C.c_uint32 * MAX_PLANES), ("loc", C.c_float * MAX_PLANES)]


def make_planes(planes):
    """planes: iterable of (axis, loc)."""
    p = Planes()
    planes = list(planes or [])
    if len(planes) > MAX_PLANES:
        raise ValueError("at most %d measurement planes per step" % MAX_PLANES)
    p.count = len(planes)
    for i, (a, l) in enumerate(planes):
        p.axis[i] = int(a)
        p.loc[i] = float(l)
    return p


_PROTOS = {
    "pcl_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "pcl_destroy": (C.c_int, [C.c_void_p]),
    "pcl_last_error": (C.c_char_p, [C.c_void_p]),
    "pcl_abi_version": (C.c_int, []),
    "pcl_device_info": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pcl_launch_count": (C.c_uint64, [C.c_void_p]),
    "pcl_stream_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pcl_stream_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "pcl_kinematics": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.c_float, C.c_int, _f32p]),
    "pcl_kinematics_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.c_float, C.c_int, _f32p, C.c_uint32]),
    "pcl_kinematics_steps_host": (C.c_int, [C.c_void_p, C.POINTER(Soa), C.c_float, C.c_int, _f32p, C.c_uint32, C.c_uint64]),
    "pcl_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.POINTER(ScatterParams), C.POINTER(Rng), C.c_void_p, C.c_void_p]),
    "pcl_escape": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.c_float, C.c_void_p]),
    "pcl_photon_step": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.c_float, C.POINTER(ScatterParams), C.POINTER(Rng), C.c_float, C.POINTER(Planes), C.c_void_p]),
    "pcl_photon_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.c_float, C.POINTER(ScatterParams), C.POINTER(Rng), C.c_float, C.POINTER(Planes), C.c_void_p, C.c_uint32]),
    "pcl_photon_step_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.POINTER(Soa), C.c_float, C.POINTER(ScatterParams), C.POINTER(Rng), C.c_float, C.POINTER(Planes), C.c_void_p, C.c_void_p]),
    "pcl_photon_steps_pp": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Pingpong), C.c_float, C.POINTER(ScatterParams), C.POINTER(Rng), C.c_float, C.POINTER(Planes), C.c_void_p, C.c_uint32, C.c_uint32]),
    "pcl_tally": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.POINTER(Planes), C.c_void_p]),
    "pcl_plane_crossers": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.POINTER(Planes), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "pcl_trace_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.c_void_p, C.c_uint64, C.c_void_p]),
    "pcl_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Soa), C.POINTER(Soa), C.c_void_p]),
    "pcl_planck_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "pcl_gravity_accel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]),
    "pcl_gravity_accel_uniform": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]),
    "pcl_gravity_kick_drift": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pcl_gravity_kick_drift_p2p": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint64]),
    "pcl_photon_step_host": (C.c_int, [C.c_void_p, C.POINTER(Soa), C.c_float, C.POINTER(ScatterParams), C.POINTER(Rng), C.c_float, C.POINTER(Planes), C.c_void_p, C.c_uint64]),
    "pcl_photon_step_host_compact": (C.c_int, [C.c_void_p, C.POINTER(Soa), C.c_float, C.POINTER(ScatterParams), C.POINTER(Rng), C.c_float, C.POINTER(Planes), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "pcl_photon_steps_host_compact": (C.c_int, [C.c_void_p, C.POINTER(Soa), C.c_float, C.POINTER(ScatterParams), C.POINTER(Rng), C.c_float, C.POINTER(Planes), C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]),
    "pcl_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "pcl_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pcl_jit_build": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_void_p)]),
    "pcl_jit_get_kernel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "pcl_jit_launch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "pcl_jit_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pcl_jit_check": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "pcl_photon_steps_jit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Soa), C.c_float, C.POINTER(ScatterParams), C.POINTER(VarnParams), C.POINTER(Rng), C.c_float, C.POINTER(Planes), C.c_void_p, C.c_uint32]),
    "pcl_scatter_jit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Soa), C.POINTER(ScatterParams), C.POINTER(VarnParams), C.POINTER(Rng), C.c_void_p, C.c_void_p]),
    "pcl_measure_fp32_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "pcl_measure_fp32x2_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "pcl_measure_copy_peak": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_double)]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


def load():
    """dlopen the in-tree library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PclError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or make -C physicl_b200/csrc). There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class Context:
    """One ``pcl_ctx``: the stand-in for the reference's ``cl_ctx`` + ``cl_q`` pair
    (physicl/__init__.py:428-429)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.pcl_init(int(device), C.byref(h))
        if rc != 0:
            raise PclError("pcl_init(%d) failed (%d): %s" % (device, rc, (self.lib.pcl_last_error(None) or b"").decode()))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pcl_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise PclError("physicl_b200 call failed (%d): %s" % (rc, (self.lib.pcl_last_error(self.handle) or b"").decode()))

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(self.handle, *args))

    @property
    def launches(self) -> int:
        return int(self.lib.pcl_launch_count(self.handle))

    def device_info(self):
        name = C.create_string_buffer(128)
        sm, hbm, l2 = C.c_int(), C.c_uint64(), C.c_uint64()
        self.call("pcl_device_info", name, 128, C.byref(sm), C.byref(hbm), C.byref(l2))
        return {"NAME": name.value.decode(), "MAX_COMPUTE_UNITS": sm.value, "GLOBAL_MEM_SIZE": hbm.value,
                "GLOBAL_MEM_CACHE_SIZE": l2.value}

    def fp32_peak_tflops(self) -> float:
        v = C.c_double()
        self.call("pcl_measure_fp32_peak", C.byref(v))
        return v.value

    def fp32x2_peak_tflops(self) -> float:
        v = C.c_double()
        self.call("pcl_measure_fp32x2_peak", C.byref(v))
        return v.value

    def copy_peak_gbs(self, nbytes: int = 1 << 30) -> float:
        v = C.c_double()
        self.call("pcl_measure_copy_peak", C.c_uint64(nbytes), C.byref(v))
        return v.value
