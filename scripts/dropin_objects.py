"""The reference's own usage pattern on this backend: N PhotonObjects added one by one, the pipeline of reference
test/test_light.py:27-37, sim.start(); sim.join(), every object current on the host at the end -- the script of
oracle/run_reference.py with only the import changed.  usage: python scripts/dropin_objects.py [n] [steps] [--profile]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def run(n=10000, steps=10):
    import physicl_b200 as physicl
    import physicl_b200.light
    import physicl_b200.newton

    t_build = time.perf_counter()
    sim = physicl.Simulation(bounds=np.array([1000, 1000, 1000]), cl_on=True, exit=lambda c: len(c.ts) >= steps)
    for _ in range(n):
        sim.add_obj(physicl.light.PhotonObject(s=np.array([0] * 3, dtype=np.double),
                                               v=np.array([physicl.light.c, 0, 0], dtype=np.double), E=np.double(1)))
    sim.add_step(0, physicl.UpdateTimeStep(lambda s: np.double(0.001)))
    sim.add_step(1, physicl.newton.NewtonianKinematicsStep())
    sim.add_step(2, physicl.light.ScatterIsotropicStep(A=np.double(0.001), n=np.double(0.001)))
    sign = physicl.light.ScatterSignMeasureStep(None, True)
    sim.add_step(3, sign)
    t0 = time.perf_counter()
    sim.start()
    sim.join()
    t1 = time.perf_counter()
    moved = sum(1 for o in sim.objects if float(o.r[0]) != 0.0)  # every object is current on the host again
    t2 = time.perf_counter()
    rows = len(sign.data)
    return {"particle_steps_per_s": n * rows / (t2 - t0), "wall_s": t2 - t0, "run_s": t1 - t0, "pull_s": t2 - t1,
            "construct_s": t0 - t_build, "n": n, "steps": rows, "moved": moved, "last_row": [float(v) for v in sign.data[-1]]}


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 10000
    steps = int(args[1]) if len(args) > 1 else 10
    run(256, 2)  # context, module load
    if "--profile" in sys.argv:
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
        out = run(n, steps)
        pr.disable()
        print(out)
        pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
    else:
        print(run(n, steps))
