"""Out-of-bounds checks with guard bands (compute-sanitizer cannot be used on the GPU pool).

Every plane a kernel may touch is carved out of ONE arena whose other words hold a sentinel: a guard band in front of
the first plane, and behind each plane the slots up to the next multiple of four plus another band.  A vectorised store
that runs past slot n - 1, a tile copy-out that overshoots the survivor count, a tail loop off by one: each lands in a
band and changes a sentinel.  Sizes are chosen around the group (4), CTA-tile (1024 x 4) and chunk boundaries.
The results themselves are checked in test_gpu_parity.py; here only "nothing outside [0, n) was written" is.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

GUARD = 64  # words; a multiple of 4, so that every plane starts 16-byte aligned
SENTINEL = 0x5A5AA5A5
C_LIGHT = 299792458.0
SIZES = [1, 5, 1023, 4096, 4099, 70_001]


@pytest.fixture(scope="module")
def ctx():
    from physicl_b200 import _capi

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    c = _capi.Context(0)
    yield c
    c.close()


class Arena:
    """names -> planes of n words inside one sentinel-filled buffer."""

    def __init__(self, names, n, device="cuda:0"):
        self.n = n
        self.stride = (n + 3) // 4 * 4 + GUARD
        self.words = torch.full((GUARD + len(names) * self.stride,), SENTINEL, dtype=torch.int32, device=device)
        self.inside = torch.zeros_like(self.words, dtype=torch.bool)
        self.off = {}
        for i, nm in enumerate(names):
            lo = GUARD + i * self.stride
            self.off[nm] = lo
            self.inside[lo:lo + n] = True

    def plane(self, nm, dtype=torch.float32):
        lo = self.off[nm]
        return self.words[lo:lo + self.n].view(dtype)

    def ptr(self, nm):
        return self.words.data_ptr() + 4 * self.off[nm]

    def check(self, what):
        torch.cuda.synchronize()
        bad = (self.words != SENTINEL) & ~self.inside
        if bool(bad.any()):
            idx = torch.nonzero(bad).flatten()[:8].tolist()
            owner = []
            for j in idx:
                nm = max((k for k in self.off if self.off[k] <= j), key=lambda k: self.off[k], default="front guard")
                owner.append("%s+%d" % (nm, j - self.off.get(nm, 0)))
            raise AssertionError("%s wrote outside its planes (n = %d): %s" % (what, self.n, owner))


def photon_arena(n, seed, with_e=True):
    names = ["x", "y", "z", "vx", "vy", "vz", "id", "nscat"] + (["e"] if with_e else [])
    a = Arena(names, n)
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(3, n))
    d /= np.linalg.norm(d, axis=0)
    for k, nm in enumerate(("x", "y", "z")):
        a.plane(nm).copy_(torch.from_numpy(rng.uniform(-1e5, 1e5, n).astype(np.float32)))
        a.plane("v" + nm).copy_(torch.from_numpy((C_LIGHT * d[k]).astype(np.float32)))
    a.plane("id", torch.int32).copy_(torch.arange(n, dtype=torch.int32))
    a.plane("nscat", torch.int32).zero_()
    if with_e:
        a.plane("e").copy_(torch.from_numpy(rng.uniform(0.4, 1.0, n).astype(np.float32)))
    return a


def soa_of(a, n=None, id_base=0, n_dev=None):
    from physicl_b200 import _capi

    s = _capi.Soa()
    s.n = a.n if n is None else n
    for nm in a.off:
        setattr(s, nm, a.ptr(nm))
    s.id_base = id_base
    s.n_dev = n_dev
    return s


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def test_the_check_fires_on_a_write_one_slot_past_the_end():
    a = Arena(["x", "y"], 5)
    a.check("nothing")
    a.words[a.off["x"] + 5] = 0
    with pytest.raises(AssertionError, match=r"x\+5"):
        a.check("a stray store")


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("accel", [0, 1])
def test_kinematics_stays_inside_its_planes(ctx, n, accel):
    names = ["x", "y", "z", "vx", "vy", "vz", "dx", "dy", "dz"] + (["ax", "ay", "az"] if accel else [])
    a = Arena(names, n)
    for nm in names:
        a.plane(nm).copy_(torch.randn(n, device="cuda:0"))
    s = soa_of(a)
    ctx.call("pcl_kinematics", stream(), C.byref(s), C.c_float(1e-3), accel, None)
    a.check("pcl_kinematics")
    ctx.call("pcl_kinematics_steps", stream(), C.byref(s), C.c_float(1e-3), accel, None, C.c_uint32(7))
    a.check("pcl_kinematics_steps")


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_in_place_photon_steps_stay_inside_their_planes(ctx, n, mode):
    from physicl_b200 import _capi

    a = photon_arena(n, seed=n + mode)
    rows = torch.zeros((8, _capi.TALLY_COLS), dtype=torch.int64, device="cuda:0")
    s = soa_of(a, id_base=7_000_000_000)
    sp = _capi.ScatterParams(k=1.3e-6, c=C_LIGHT, mode=mode)
    pl = _capi.make_planes([(0, 1.0e5)])
    rg = _capi.Rng(seed=5, step=0)
    ctx.call("pcl_photon_step", stream(), C.byref(s), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(4.0e5 ** 2), C.byref(pl),
             C.c_void_p(rows.data_ptr()))
    a.check("pcl_photon_step")
    rg = _capi.Rng(seed=5, step=1)
    ctx.call("pcl_photon_steps", stream(), C.byref(s), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(4.0e5 ** 2), C.byref(pl),
             C.c_void_p(rows.data_ptr()), C.c_uint32(8))
    a.check("pcl_photon_steps")


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("nsteps", [1, 5])
def test_compacting_launches_write_survivors_only(ctx, n, mode, nsteps):
    """The survivors land in [0, n_out) of the partner planes and nowhere else: the partner arena is cleared to the
    sentinel first, so also the slots between n_out and n must still hold it afterwards."""
    from physicl_b200 import _capi

    src = photon_arena(n, seed=3 * n + mode)
    dst = photon_arena(n, seed=1)
    dst.words.fill_(SENTINEL)
    rows = torch.zeros((8, _capi.TALLY_COLS), dtype=torch.int64, device="cuda:0")
    n_dev = torch.zeros(2, dtype=torch.int64, device="cuda:0")
    sp = _capi.ScatterParams(k=1.3e-6, c=C_LIGHT, mode=mode)
    pl = _capi.make_planes([])
    rg = _capi.Rng(seed=9, step=0)
    if nsteps == 1:
        s, d = soa_of(src), soa_of(dst)
        ctx.call("pcl_photon_step_compact", stream(), C.byref(s), C.byref(d), C.c_float(1e-3), C.byref(sp), C.byref(rg),
                 C.c_float(1.5e5 ** 2), C.byref(pl), C.c_void_p(rows.data_ptr()), C.c_void_p(n_dev.data_ptr()))
        n_out = int(n_dev[0].item())
    else:
        pp = _capi.Pingpong()
        for k, ar in enumerate((src, dst)):
            for nm in ar.off:
                setattr(pp.buf[k], nm, ar.ptr(nm))
            pp.buf[k].n = n
        n_dev[0] = n
        pp.n_dev = n_dev.data_ptr()
        pp.cur, pp.id_valid = 0, 1
        ctx.call("pcl_photon_steps_pp", stream(), C.byref(pp), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(1.5e5 ** 2),
                 C.byref(pl), C.c_void_p(rows.data_ptr()), C.c_uint32(nsteps), C.c_uint32(nsteps))
        assert pp.cur == 1
        n_out = int(n_dev[1].item())
    torch.cuda.synchronize()
    assert 0 <= n_out <= n
    src.check("the compacting launch (source planes)")
    dst.inside[:] = False
    for nm, lo in dst.off.items():
        dst.inside[lo:lo + n_out] = True
    dst.check("the compacting launch (partner planes, beyond the %d survivors)" % n_out)
    ids = dst.plane("id", torch.int32)[:n_out]
    assert len(torch.unique(ids)) == n_out  # every survivor once


@pytest.mark.parametrize("n", [1, 7, 1025, 70_001])
def test_stable_compaction_writes_survivors_only(ctx, n):
    src = photon_arena(n, seed=n)
    x = src.plane("x")
    x[torch.rand(n, device="cuda:0") < 0.4] = float("nan")
    dst = photon_arena(n, seed=2)
    dst.words.fill_(SENTINEL)
    n_dev = torch.zeros(1, dtype=torch.int64, device="cuda:0")
    s, d = soa_of(src), soa_of(dst)
    ctx.call("pcl_compact", stream(), C.byref(s), C.byref(d), C.c_void_p(n_dev.data_ptr()))
    torch.cuda.synchronize()
    n_out = int(n_dev[0].item())
    assert n_out == int((~torch.isnan(x)).sum().item())
    src.check("pcl_compact (source)")
    dst.inside[:] = False
    for nm, lo in dst.off.items():
        dst.inside[lo:lo + n_out] = True
    dst.check("pcl_compact (destination)")


@pytest.mark.parametrize("n", [1, 3, 513, 3001])
@pytest.mark.parametrize("uniform", [False, True])
def test_gravity_stays_inside_its_planes(ctx, n, uniform):
    a = Arena(["ax", "ay", "az", "vx", "vy", "vz", "x", "y", "z"], n)
    pos = Arena(["posm"], 4 * n)  # the packed (x, y, z, m) array: 4n contiguous words
    pm = pos.plane("posm").view(n, 4)
    pm.copy_(torch.randn(n, 4, device="cuda:0"))
    pm[:, 3] = 1.0 / n
    for nm in ("vx", "vy", "vz", "x", "y", "z"):
        a.plane(nm).copy_(torch.randn(n, device="cuda:0"))
    p = lambda ar, nm: C.c_void_p(ar.ptr(nm))
    fn = "pcl_gravity_accel_uniform" if uniform else "pcl_gravity_accel"
    ctx.call(fn, stream(), p(pos, "posm"), C.c_uint64(n), p(pos, "posm"), C.c_uint64(n), C.c_float(1.0), C.c_float(1e-3),
             p(a, "ax"), p(a, "ay"), p(a, "az"), 0, C.c_uint64(0), C.c_uint64(0))
    a.check(fn)
    pos.check(fn + " (bodies)")
    ctx.call("pcl_gravity_kick_drift", stream(), C.c_uint64(n), p(pos, "posm"), p(a, "vx"), p(a, "vy"), p(a, "vz"), p(a, "ax"),
             p(a, "ay"), p(a, "az"), C.c_float(1e-3), p(a, "x"), p(a, "y"), p(a, "z"))
    a.check("pcl_gravity_kick_drift")
    pos.check("pcl_gravity_kick_drift (bodies)")
    assert bool(torch.isfinite(a.plane("ax")).all())


@pytest.mark.parametrize("n", [1, 6, 4099, 300_001])
def test_host_round_trip_stays_inside_the_host_planes(ctx, n):
    """pcl_photon_steps_host_compact: the survivors come back into [0, n_out) of the caller's host planes."""
    from physicl_b200 import _capi

    names = ["x", "y", "z", "vx", "vy", "vz", "id"]
    stride = (n + 3) // 4 * 4 + GUARD
    words = torch.full((GUARD + len(names) * stride,), SENTINEL, dtype=torch.int32).pin_memory()
    off = {nm: GUARD + i * stride for i, nm in enumerate(names)}
    rng = np.random.default_rng(n)
    d = rng.normal(size=(3, n))
    d /= np.linalg.norm(d, axis=0)
    for k, nm in enumerate(("x", "y", "z")):
        words[off[nm]:off[nm] + n].view(torch.float32).copy_(torch.from_numpy(rng.uniform(-1e5, 1e5, n).astype(np.float32)))
        words[off["v" + nm]:off["v" + nm] + n].view(torch.float32).copy_(torch.from_numpy((C_LIGHT * d[k]).astype(np.float32)))
    words[off["id"]:off["id"] + n].copy_(torch.arange(n, dtype=torch.int32))
    s = _capi.Soa()
    s.n = n
    for nm in names:
        setattr(s, nm, words.data_ptr() + 4 * off[nm])
    sp = _capi.ScatterParams(k=1.3e-6, c=C_LIGHT, mode=0)
    pl = _capi.make_planes([])
    rows = np.zeros((8, _capi.TALLY_COLS), np.int64)
    n_out = C.c_uint64(0)
    for call in range(2):
        rg = _capi.Rng(seed=3, step=4 * call)
        ctx.call("pcl_photon_steps_host_compact", C.byref(s), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(1.5e5 ** 2),
                 C.byref(pl), rows.ctypes.data_as(C.c_void_p), C.c_uint64(1 << 16), C.c_uint32(4), C.byref(n_out))
        assert n_out.value <= s.n
        s.n = n_out.value
    inside = torch.zeros_like(words, dtype=torch.bool)
    for nm in names:
        inside[off[nm]:off[nm] + n] = True
    assert not bool(((words != SENTINEL) & ~inside).any()), "host planes: wrote outside [0, n)"
