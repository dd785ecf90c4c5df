"""NumPy float64 restatement of the reference's per-particle step path.  TEST INFRASTRUCTURE.

Each function names the reference lines it follows (paths relative to /root/reference).  Arrays
are structure-of-arrays: ``r``, ``v``, ``dr`` are ``(3, N)`` float64.  Nothing here is imported
by the product package.
"""
from __future__ import annotations

import numpy as np

# physicl/light.py:14-16 (SI values; code units scale them through Measurement)
C_SI = 299792458.0
H_SI = 6.62607015e-34
KB_SI = 1.380649e-23


def kinematics(r, v, dt):
    """physicl/newton.py:14-16: ``obj.dr = obj.v * sim.dt; obj.r += obj.dr`` (no ``a``, no ``v`` update)."""
    dr = v * dt
    return r + dr, dr


def kinematics_accel(r, v, a, dt):
    """NEW (not in the reference): ``v += a dt; dr = v dt; r += dr``."""
    v = v + a * dt
    dr = v * dt
    return r + dr, v, dr


def scale_uniforms(u_theta, u_phi):
    """physicl/light.py:285: ``np.random.random() * 2 * np.pi`` and ``np.random.random() * np.pi``."""
    return u_theta * 2 * np.pi, u_phi * np.pi


def pcoll(dr, A, n, E=None, hc=None):
    """physicl/light.py:305-306 (+ :300-301): ``sqrt(pow(d0,2)+pow(d1,2)+pow(d2,2))``; ``A * n * norm [* pow(hc/E, -4)]``."""
    norm = np.sqrt(dr[0] ** 2 + dr[1] ** 2 + dr[2] ** 2)
    p = A * n * norm
    if E is not None:
        p = p * (hc / E) ** -4.0
    return p


def eval_density_expr(expr, r, E=None, dr=None, A=None, n=None):
    """The user's OpenCL-C number-density expression of ``ScatterIsotropicStep(variable_n=True)``
    (physicl/light.py:295-299) evaluated over all photons in float64: ``r0[gid]``, ``r1[gid]``, ``r2[gid]``
    are the position components AFTER the kinematics step (the inputs are gathered when the scatter
    step runs, physicl/__init__.py:606-629)."""
    ns = {"pow": np.power, "exp": np.exp, "sqrt": np.sqrt, "log": np.log, "sin": np.sin, "cos": np.cos, "fabs": np.abs,
          "gid": slice(None), "r0": r[0], "r1": r[1], "r2": r[2], "A": A, "n": n}
    if E is not None:
        ns["E"] = E
    if dr is not None:
        ns.update(d0=dr[0], d1=dr[1], d2=dr[2], norm=np.sqrt(dr[0] ** 2 + dr[1] ** 2 + dr[2] ** 2))
    return eval(expr, {"__builtins__": {}}, ns) * np.ones(r.shape[1])


def pcoll_variable_n(dr, r, expr, kernel_A, kernel_n=None, E=None, hc=None):
    """physicl/light.py:299-306 with ``variable_n=True``: ``pcoll = A * (<expr>) * norm [* pow(hc/E, -4)]``
    where the kernel name ``A`` holds the step's ``n`` (light.py:287 swaps the two constants), so the
    step's cross-section does not enter."""
    norm = np.sqrt(dr[0] ** 2 + dr[1] ** 2 + dr[2] ** 2)
    p = kernel_A * eval_density_expr(expr, r, E, dr, kernel_A, kernel_n) * norm
    if E is not None:
        p = p * (hc / E) ** -4.0
    return p


def scatter_sphere_kernel(dr, rtheta, rphi, rnd, A, n, c, E=None, hc=None):
    """physicl/light.py:303-315 kernel body.  Returns res (3, N) with res[0] = NaN where unaffected
    (res[1], res[2] are undefined there in the reference; NaN here)."""
    p = pcoll(dr, A, n, E, hc)
    hit = p >= rnd
    res = np.full((3, rnd.size), np.nan)
    res[0, hit] = (c * np.sin(rtheta) * np.cos(rphi))[hit]
    res[1, hit] = (c * np.sin(rtheta) * np.sin(rphi))[hit]
    res[2, hit] = (c * np.cos(rtheta))[hit]
    return res


def scatter_writeback(v, res):
    """physicl/light.py:325-331: scattered photons take the new v, ``dv = v_new - v_old``; others ``dv = 0``."""
    hit = ~np.isnan(res[0])
    v_new = np.where(hit, res, v)
    dv = np.where(hit, v_new - v, 0.0)
    return v_new, dv, hit


def scatter_delete_kernel(dr, rnd, n, A):
    """physicl/light.py:146-158 / :239-249: ``result = (A*n*norm >= rand) ? 1 : 0`` (int32)."""
    return (pcoll(dr, A, n) >= rnd).astype(np.int32)


def sign_tally(v):
    """physicl/light.py:414-431: strict ``>`` on velocity components over all objects."""
    return int(v.shape[1]), int((v[0] > 0).sum()), int((v[1] > 0).sum()), int((v[2] > 0).sum())


def plane_tally(r, dr, loc):
    """physicl/light.py:385-399: the plane is the one non-NaN coordinate of ``loc``; a crossing is
    ``r - dr <= loc <= r`` or ``r - dr >= loc >= r`` (closed interval, ``r - dr`` recomputed)."""
    loc = np.asarray(loc, float)
    ax = 0 if not np.isnan(loc[0]) else (1 if not np.isnan(loc[1]) else 2)
    prev = r[ax] - dr[ax]
    hit = ((prev <= loc[ax]) & (loc[ax] <= r[ax])) | ((prev >= loc[ax]) & (loc[ax] >= r[ax]))
    return int(hit.sum())


# ---- emission: physicl/light.py:53-104 --------------------------------------------------------
def planck_density(E, T, kB=KB_SI):
    """physicl/light.py:53-60: ``15/(pi^4 kT) * (E/kT)^3 * e^(-E/kT)`` (a Wien-type law, not Planck's)."""
    x = E / (kB * T)
    return 15.0 / (np.pi ** 4 * kB * T) * x ** 3 * np.exp(-x)


def planck_bin_masses(E_min, E_max, T, bins, kB=KB_SI):
    """physicl/light.py:82-86: per-interval integrals of the density over ``linspace(E_min, E_max, bins)``.
    Closed form of what the reference gets from scipy.integrate.quad:
    int x^3 e^-x dx = -e^-x (x^3 + 3x^2 + 6x + 6)."""
    E = np.linspace(E_min, E_max, bins)
    x = E / (kB * T)
    F = -np.exp(-x) * (x ** 3 + 3 * x ** 2 + 6 * x + 6) * (15.0 / np.pi ** 4)
    return E, F[1:] - F[:-1]


def planck_cdf(E_min, E_max, T, bins, kB=KB_SI):
    """physicl/light.py:88-93: normalise by the python ``sum`` and accumulate left to right."""
    E, gamma = planck_bin_masses(E_min, E_max, T, bins, kB)
    tot = 0.0
    for g in gamma:  # builtin sum(): sequential float64 adds
        tot += g
    norm = gamma / tot
    cdf = np.empty_like(norm)
    acc = 0.0
    for i, g in enumerate(norm):
        acc = g if i == 0 else acc + g
        cdf[i] = acc
    return E, cdf


def planck_pick(cdf, u):
    """physicl/light.py:101-104: first ``x >= 1`` with ``cdf[x] >= u >= cdf[x-1]``; -1 where the
    reference falls off the loop and returns None."""
    u = np.atleast_1d(u)
    idx = np.searchsorted(cdf, u, side="left")  # first idx with cdf[idx] >= u
    out = idx.astype(np.int64)
    out[idx >= cdf.size] = -1
    zero = idx == 0
    out[zero] = np.where((u[zero] == cdf[0]) & (cdf.size > 1), 1, -1)
    return out


# ---- NEW steps: definitions used as oracle (parity unpinned by the reference) ----------------
def escape_mask(r, R):
    """NEW: photons with |r| >= R retire."""
    return (r[0] ** 2 + r[1] ** 2 + r[2] ** 2) >= R * R


def gravity_accel(pos, m, G, eps2):
    """NEW: a_i = G sum_j m_j (r_j - r_i) / (|r_ij|^2 + eps2)^(3/2), j == i included (zero term)."""
    d = pos[:, None, :] - pos[:, :, None]  # d[:, i, j] = r_j - r_i
    r2 = (d ** 2).sum(0) + eps2
    w = m[None, :] / (r2 * np.sqrt(r2))
    return G * (d * w[None]).sum(2)
