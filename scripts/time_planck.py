"""Time pcl_planck_sample (BASELINE configs[2] emission: 64 Mi photons, 50 000 bins) with CUDA events.
usage: python scripts/time_planck.py [n] [bins] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import physicl_b200 as phys
import physicl_b200.light  # noqa: F401
from physicl_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 2 ** 20
bins = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ctx = _capi.Context(0)
E_min = float(phys.light.E_from_wavelength(2500e-9))
E_max = float(phys.light.E_from_wavelength(200e-9))
for want_bins in (False, True):
    best = 1e9
    for _ in range(reps):
        t = {}
        phys.light.planck_sample_device(ctx, n, E_min, E_max, 5778.0, bins=bins, seed=2025, want_bins=want_bins, timing=t)
        best = min(best, t["device_ms"])
    nbytes = n * (8 if want_bins else 4)
    print(f"n={n} bins={bins} bin_out={want_bins}: {best * 1e3:.1f} us = {n / best / 1e6:.1f} G photons/s, "
          f"{nbytes / best / 1e6:.0f} GB/s written")
