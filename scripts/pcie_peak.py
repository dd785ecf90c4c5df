"""Measurement aid (GPU box): pinned-memory copy bandwidth over PCIe, one direction at a time and both at once:
the ceiling of the host-buffer (`e2e`) path, which moves 28 B up and 28 B down per photon and timestep."""
import torch

n = 256 * 2 ** 20
h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(up, down, reps=8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


timed(True, True, 2)
print("H2D alone      %.1f GB/s" % timed(True, False))
print("D2H alone      %.1f GB/s" % timed(False, True))
print("both at once   %.1f GB/s in each direction" % timed(True, True))
