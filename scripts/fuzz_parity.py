"""Randomised parity soak (GPU box): random sizes, modes, plane sets and timestep counts through the C ABI against the
CPU oracle, bit for bit, for a given number of seconds.  Covers the fused photon steps (in place, several timesteps per
launch; and through the compacting host-buffer entry point), the stable compaction, the host-buffer kinematics and the
emission sampler.  Prints one summary line.
usage: python scripts/fuzz_parity.py [seconds] [seed]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import gpu_util as u
import oracle
from physicl_b200 import _capi

PLANCK_MAX_WORK = 2.4e10  # photons x table entries per emission case (the test suite lowers it)
rng = np.random.default_rng(2026)  # re-seeded by main()
ctx = None
dev = None


def pick_n():
    return int(rng.choice([rng.integers(1, 40), rng.integers(40, 5000), rng.integers(5000, 300_000)]))


def case_photon_steps():
    n, mode, steps = pick_n(), int(rng.integers(0, 4)), int(rng.integers(1, 20))
    r, v = u.random_photons(n, seed=int(rng.integers(1 << 30)), spread=float(rng.choice([1e3, 2e5])))
    E = rng.uniform(0.2, 1.0, n) if (mode & 1 or rng.random() < 0.3) else None
    k = float(rng.choice([0.0, 2e-7, 2.5e-6, 1e-3]))
    r2 = float(rng.choice([0.0, (1.5e6) ** 2, (3e5) ** 2]))
    planes = [(int(rng.integers(0, 3)), float(rng.normal(0, 1e5))) for _ in range(int(rng.integers(0, 4)))]
    nscat = bool(rng.random() < 0.5)
    id_base = int(rng.choice([0, 12345, (1 << 32) - n - 1 if n < (1 << 20) else 0, (5 << 32) + 7]))
    st, g = u.make_store(ctx, r, v, E=E, nscat=nscat, id_base=id_base)
    host = u.host_state(g)
    sp = _capi.ScatterParams(k=k, c=u.C_LIGHT, mode=mode)
    pl = _capi.make_planes(planes)
    first = st.new_rows(steps)
    soa = g.soa()
    soa.dx = soa.dy = soa.dz = None
    seed, step0 = int(rng.integers(1 << 40)), int(rng.integers(1 << 20))
    rg = _capi.Rng(seed=seed, step=step0)
    ctx.call("pcl_photon_steps", st.stream(), C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(r2), C.byref(pl),
             st.row_ptr(first), C.c_uint32(steps))
    rows = np.array([st.read_row(first + i) for i in range(steps)])
    rows_t = np.array([oracle.photon_step_f32(host, 1e-3, k, u.C_LIGHT, mode, seed=seed, step=step0 + s, r2_escape=np.float32(r2),
                                              planes=planes, id_base=id_base) for s in range(steps)])
    assert np.array_equal(rows, rows_t), ("tally rows", n, mode, steps, k, r2, planes)
    live = ~np.isnan(host["x"])
    assert np.array_equal(~np.isnan(g.download("x")), live)
    for nm in host:
        if nm == "id":
            continue
        assert u.same_bits(g.download(nm)[live], host[nm][live]), (nm, n, mode, steps)


def case_compact():
    n = pick_n()
    r, v = u.random_photons(n, seed=int(rng.integers(1 << 30)))
    kw = {}
    if rng.random() < 0.5:
        kw["E"] = np.linspace(1.0, 2.0, n)
    if rng.random() < 0.4:
        kw["a"] = rng.normal(0, 3, (3, n))
    st, g = u.make_store(ctx, r, v, nscat=bool(rng.random() < 0.5), **kw)
    if rng.random() < 0.4:
        g.ensure("dx", "dy", "dz")
        for nm in ("dx", "dy", "dz"):
            g.upload(nm, rng.normal(0, 1, n).astype(np.float32))
    dead = rng.random(n) < float(rng.choice([0.0, 0.05, 0.5, 0.95, 1.0]))
    x = g.download("x").copy()
    x[dead] = np.nan
    g.upload("x", x)
    before = u.host_state(g)
    n_live = st.compact("photon")
    keep = ~dead
    assert n_live == int(keep.sum()) == g.n, (n, n_live, int(keep.sum()))
    if n_live:
        assert np.array_equal(g.download("id"), np.nonzero(keep)[0].astype(np.uint32))
        for nm in before:
            assert u.same_bits(g.download(nm), before[nm][keep]), (nm, n)


def case_kinematics_host():
    n, accel, with_dr = pick_n(), int(rng.integers(0, 3)), bool(rng.random() < 0.6)
    chunk = int(rng.choice([4096, 65_536, 1 << 20]))
    names = ["x", "y", "z", "vx", "vy", "vz"] + (["ax", "ay", "az"] if accel == 1 else []) + (["dx", "dy", "dz"] if with_dr else [])
    pad = 8
    full = {}
    for nm in names:
        t = torch.full((n + 2 * pad,), 777.0, dtype=torch.float32)
        if rng.random() < 0.5:
            t = t.pin_memory()
        t[pad:pad + n] = torch.from_numpy(rng.normal(0, 30, n).astype(np.float32))
        full[nm] = t
    host = {nm: t[pad:pad + n].numpy().copy() for nm, t in full.items()}
    soa = _capi.Soa()
    soa.n = n
    for nm, t in full.items():
        setattr(soa, nm, t.data_ptr() + 4 * pad)
    au = np.array([0.5, 0.0, -9.81], np.float32)
    pau = au.ctypes.data_as(C.POINTER(C.c_float)) if accel == 2 else None
    k = int(rng.integers(1, 12))
    ctx.call("pcl_kinematics_steps_host", C.byref(soa), C.c_float(2e-3), int(accel != 0), pau, C.c_uint32(k), C.c_uint64(chunk))
    for _ in range(k):
        oracle.kinematics_f32(host, 2e-3, accel, au if accel == 2 else None)
    for nm, t in full.items():
        got = t.numpy()
        assert (got[:pad] == 777.0).all() and (got[pad + n:] == 777.0).all(), nm
        assert u.same_bits(got[pad:pad + n], host[nm]), (nm, n, accel, k, chunk)


def case_planck():
    ncdf = int(rng.choice([1, 2, 199, 999, 5000, 49_999, 65_535, 65_536, 70_000]))
    n = int(rng.choice([pick_n(), rng.integers(262_144, 400_000)]))
    n = min(n, int(PLANCK_MAX_WORK // max(ncdf, 1)))  # the oracle scans linearly: a few seconds at most
    w = rng.uniform(0.0, 1.0, ncdf) ** float(rng.choice([1.0, 8.0, 40.0]))
    w[rng.integers(0, ncdf)] += 1e-3
    cdf = np.cumsum(w / w.sum())
    if rng.random() < 0.3:
        cdf = np.minimum(cdf * (1 + 1e-9), 1.0 + 1e-15)
    id_base = int(rng.choice([0, 3, 4, (7 << 32) + 2, 123]))
    cdf_d = torch.from_numpy(cdf).to(dev)
    pad = 16
    e = torch.full((n + 2 * pad,), -7.0, dtype=torch.float32, device=dev)
    b = torch.full((n + 2 * pad,), -7, dtype=torch.int32, device=dev)
    seed = int(rng.integers(1 << 40))
    ctx.call("pcl_planck_sample", C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), C.c_uint64(n), C.c_uint64(id_base), C.c_uint64(seed),
             C.c_void_p(cdf_d.data_ptr()), C.c_uint32(ncdf), C.c_float(0.25), C.c_float(1e-5), C.c_void_p(e[pad:].data_ptr()),
             C.c_void_p(b[pad:].data_ptr()))
    torch.cuda.synchronize()
    e, b = e.cpu().numpy(), b.cpu().numpy()
    assert (b[:pad] == -7).all() and (b[pad + n:] == -7).all()
    e_or, b_or = oracle.planck_sample(n, id_base, seed, cdf, np.float32(0.25), np.float32(1e-5))
    assert np.array_equal(b[pad:pad + n], b_or), (ncdf, n, id_base)
    assert u.same_bits(e[pad:pad + n], e_or)


def case_photon_host_compact():
    """pcl_photon_steps_host_compact (chunked H2D -> fused launch with retirement -> survivors D2H) against the twin."""
    n, mode = pick_n(), int(rng.integers(0, 4))
    r, v = u.random_photons(n, seed=int(rng.integers(1 << 30)), spread=float(rng.choice([1e3, 2e5])))
    wave = bool(mode & 1)
    E = rng.uniform(0.2, 1.0, n) if wave else None
    k = float(rng.choice([2e-7, 2.5e-6]))
    r2 = float(rng.choice([0.0, (1.5e6) ** 2, (3e5) ** 2]))
    planes = [(int(rng.integers(0, 3)), float(rng.normal(0, 1e5))) for _ in range(int(rng.integers(0, 3)))]
    id_base = int(rng.choice([0, 99, (3 << 32) + 11]))
    chunk = int(rng.choice([4096, 32_768, 1 << 20]))
    st, g = u.make_store(ctx, r, v, E=E, id_base=id_base)
    twin = u.host_state(g)
    twin["id"] = np.arange(n, dtype=np.uint32)
    names = list(u.PLANE_NAMES) + (["e"] if wave else [])
    host = {nm: torch.from_numpy(twin[nm].copy()) for nm in names}
    host["id"] = torch.arange(n, dtype=torch.int32)
    if rng.random() < 0.5:
        host = {nm: t.pin_memory() for nm, t in host.items()}
    sp = _capi.ScatterParams(k=k, c=u.C_LIGHT, mode=mode)
    pl = _capi.make_planes(planes)
    rows = np.zeros((8, _capi.TALLY_COLS), np.int64)
    n_out = C.c_uint64(0)
    seed, s, n_live = int(rng.integers(1 << 40)), int(rng.integers(1 << 20)), n
    for _ in range(int(rng.integers(1, 4))):
        m = int(rng.integers(1, 9))
        soa = _capi.Soa()
        soa.n, soa.id_base = n_live, id_base
        for nm, t in host.items():
            setattr(soa, nm, t.data_ptr())
        rg = _capi.Rng(seed=seed, step=s)
        ctx.call("pcl_photon_steps_host_compact", C.byref(soa), C.c_float(1e-3), C.byref(sp), C.byref(rg), C.c_float(r2), C.byref(pl),
                 rows.ctypes.data_as(C.c_void_p), C.c_uint64(chunk), C.c_uint32(m), C.byref(n_out))
        for q in range(m):
            row_t = oracle.photon_step_f32(twin, 1e-3, k, u.C_LIGHT, mode, seed=seed, step=s + q, r2_escape=np.float32(r2), planes=planes,
                                           id_base=id_base)
            assert np.array_equal(rows[q], row_t), ("host row", n, mode, m, q, rows[q], row_t)
        s += m
        n_live = int(n_out.value)
        live = ~np.isnan(twin["x"])
        assert n_live == int(live.sum()), (n_live, int(live.sum()))
        ids = host["id"].numpy()[:n_live].view(np.uint32)
        order = np.argsort(ids)
        assert np.array_equal(ids[order], np.nonzero(live)[0].astype(np.uint32))
        for nm in names:
            assert u.same_bits(host[nm].numpy()[:n_live][order], twin[nm][live]), (nm, n, mode, m)
        # the twin carries on with the survivors only, in id order, like the host planes after sorting would
        if n_live == 0:
            break


CASES = [("photon_host_compact", case_photon_host_compact), ("photon_steps", case_photon_steps), ("compact", case_compact),
         ("kinematics_host", case_kinematics_host), ("planck", case_planck)]


def main(budget=60.0, seed=2026, device=0):
    """Run random cases for `budget` seconds; returns the number of cases per family (raises on the first mismatch)."""
    global rng, ctx, dev
    rng = np.random.default_rng(seed)
    ctx = _capi.Context(device)
    dev = torch.device("cuda", device)
    counts = {name: 0 for name, _ in CASES}
    t0 = time.time()
    try:
        while time.time() - t0 < budget:
            name, fn = CASES[int(rng.integers(0, len(CASES)))]
            fn()
            counts[name] += 1
    finally:
        ctx.close()
    return counts, time.time() - t0


if __name__ == "__main__":
    counts, secs = main(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    print("fuzz ok in %.0f s: %s" % (secs, counts))
