"""Multi-rank check, launched by tests/test_gpu_multi.py (or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mp_sharded_check.py

Every rank owns a contiguous block of the particles (Simulation(shard=True)); rank 0 also runs the
unsharded simulation and compares: integer tallies identical, surviving photons identical by global id,
gravity accelerations equal to the single-GPU result within float summation order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import physicl_b200 as phys  # noqa: E402
import physicl_b200.light  # noqa: E402
import physicl_b200.newton  # noqa: E402


def photon_run(shard, n, steps, device):
    sim = phys.Simulation(cl_on=True, device=device, shard=shard, seed=77)
    r = np.zeros((3, n))
    v = np.zeros((3, n))
    v[0] = float(phys.light.c)
    sim.add_particles(r, v)
    sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(1e-3)))
    sim.add_step(1, phys.newton.NewtonianKinematicsStep())
    sim.add_step(2, phys.light.ScatterIsotropicStep(A=np.double(1e-3), n=np.double(1e-3)))
    esc = phys.light.EscapeSphereStep(1.5e6)
    sim.add_step(3, esc)
    sign = phys.light.ScatterSignMeasureStep(None, True)
    plane = phys.light.ScatterMeasureStep(None, True, [[6.0e5, np.nan, np.nan]])
    sim.add_step(4, sign)
    sim.add_step(5, plane)
    sim.run_steps(steps)
    return sim, esc, sign, plane


def gravity_run(shard, n, steps, device, pos, vel):
    sim = phys.Simulation(cl_on=True, device=device, shard=shard)
    sim.add_particles(pos, vel, kind="object")
    sim.add_step(0, phys.UpdateTimeStep(lambda s: np.double(1e-3)))
    sim.add_step(1, phys.newton.NewtonianGravityStep(G=1.0, eps2=1e-3, masses=np.full(n, 1.0 / n, np.float32)))
    sim.run_steps(steps)
    return sim


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, steps = 200_003, 24
    sim, esc, sign, plane = photon_run(True, n, steps, local)
    rows_s, rows_p = np.array(sign.data), np.array(plane.data)  # all-reduced over ranks
    escaped = esc.escaped_all_ranks()
    snap = sim.store.snapshot("photon")
    gids = snap["id"].astype(np.int64) + sim.store.group("photon").id_base
    n_live = phys.dist.all_reduce_int(len(gids))
    ok = True
    if rank == 0:
        ref, esc1, sign1, plane1 = photon_run(False, n, steps, local)
        a, b = np.array(sign1.data), np.array(plane1.data)
        ok &= np.array_equal(a[:, 1:], rows_s[:, 1:]) and np.array_equal(b[:, 1:], rows_p[:, 1:])
        ok &= np.array_equal(esc1.escaped, escaped)
        ok &= n_live == int(a[-1, 1])
        full = ref.store.snapshot("photon")
        pick = np.isin(full["id"].astype(np.int64), gids)
        ok &= int(pick.sum()) == len(gids)
        for nm in ("x", "y", "z", "vx", "vy", "vz"):
            ok &= np.array_equal(full[nm][pick].view(np.uint32), snap[nm].view(np.uint32))
        print("photon shard check:", "ok" if ok else "MISMATCH", "alive", int(a[-1, 1]), "escaped", int(escaped.sum()))
    # gravity: N divisible by world
    ng = 4096
    rng = np.random.default_rng(3)
    pos, vel = rng.normal(size=(3, ng)), rng.normal(0, 0.1, (3, ng))
    g1 = gravity_run(False, ng, 5, local, pos, vel)  # every rank computes the unsharded answer for its own block
    whole = g1.store.snapshot("object")
    for mode in ("p2p", "nccl"):  # stores from the kick-drift kernel into peer memory / NCCL all-gather
        os.environ["PCL_GRAVITY_EXCHANGE"] = mode
        gs = gravity_run(True, ng, 5, local, pos, vel)
        used = type(gs.steps[1]._state["xchg"]).__name__
        mine = gs.store.snapshot("object")
        lo = gs.store.group("object").id_base
        m = len(mine["x"])
        err = max(np.abs(whole[nm][lo:lo + m] - mine[nm]).max() for nm in ("x", "y", "z", "vx", "vy", "vz"))
        scale = max(np.abs(whole[nm]).max() for nm in ("vx", "vy", "vz"))
        gok = bool(err <= 1e-5 * scale)
        worst = torch.tensor([0 if gok else 1], device="cuda")
        dist.all_reduce(worst)
        if rank == 0:
            print("gravity shard check (%s -> %s):" % (mode, used), "ok" if int(worst.item()) == 0 else "MISMATCH", "max err on rank 0", err)
        ok &= int(worst.item()) == 0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
