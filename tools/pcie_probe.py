"""PCIe floor of the end-to-end numbers: pinned host <-> device copy rates on this box, one direction at a time and both at once
(what `add` does: input up, marked audio down).  bench.py's e2e legs cannot be faster than bytes / these rates."""
import time
import torch

n = 635_040_000            # bytes of 1 h stereo 16 bit PCM
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def up():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    up(); down()


t_up, t_down, t_both = timed(up), timed(down), timed(both)
print("H2D %.1f GB/s (%.2f ms)   D2H %.1f GB/s (%.2f ms)   both at once %.2f ms = %.1f + %.1f GB/s" % (
    n / t_up / 1e9, t_up * 1e3, n / t_down / 1e9, t_down * 1e3, t_both * 1e3, n / t_both / 1e9, n / t_both / 1e9))
