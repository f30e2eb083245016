"""PCIe probe: pinned H2D, D2H and both at once (two streams), to back the e2e analysis in DESIGN.md section 5."""
import time
import torch

n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt


for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("H2D + D2H concurrently", True, True)):
    run(a, b, 3)
    dt = run(a, b)
    moved = n * (int(a) + int(b))
    print("%-26s %.2f ms per %d MiB each way  -> %.1f GB/s total" % (name, dt * 1e3, n >> 20, moved / dt / 1e9))
