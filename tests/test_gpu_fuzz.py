"""A short run of the randomised parity soak (scripts/fuzz_parity.py) inside the GPU suite: random sizes, modes, plane sets,
id bases, timestep counts, chunk sizes and tables through the C ABI against the CPU oracle, bit for bit."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_randomised_parity_soak():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import fuzz_parity

    fuzz_parity.PLANCK_MAX_WORK = 2.0e9  # keeps one emission case under a second of oracle time
    counts, secs = fuzz_parity.main(budget=8.0, seed=20261018)
    assert sum(counts.values()) >= 20 and all(c > 0 for c in counts.values()), counts  # every family ran
