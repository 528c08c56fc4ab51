"""PCIe microbenchmark behind the e2e numbers: pinned host <-> device copies of the bench's per-step volume (5 fields of 67 MB each
way), one direction at a time and both at once on two streams."""
import time
import torch

n = 8388608
h_in = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(5)]
h_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(5)]
d_in = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(5)]
d_out = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(5)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
nbytes = 5 * n * 8


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                for a, b in zip(d_in, h_in):
                    a.copy_(b, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for a, b in zip(h_out, d_out):
                    a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for _ in range(2):
    run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D alone {a*1e3:.2f} ms ({nbytes/a/1e9:.1f} GB/s)  D2H alone {b*1e3:.2f} ms ({nbytes/b/1e9:.1f} GB/s)  both at once {c*1e3:.2f} ms "
      f"({2*nbytes/c/1e9:.1f} GB/s aggregate)")
