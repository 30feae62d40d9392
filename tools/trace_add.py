"""End-to-end `add` with 16 bit pinned host buffers (1 h stereo): wall time per call for several piece sizes (AWM_PIECE, 1024-sample
frames per pipeline piece) next to the plain copy times of the same buffers -- shows how close the H2D / kernel / D2H pipeline of
awm_embed_s16 gets to the PCIe floor."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from audiowmark_b200 import hostapi as H

n = 60 * 60 * 44100
H.set_params()
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.rand((n, 2), device="cuda", generator=g, dtype=torch.float32) - 0.5
x16 = torch.empty((n, 2), dtype=torch.int16, pin_memory=True)
y16 = torch.empty((n, 2), dtype=torch.int16, pin_memory=True)
x16.copy_(torch.floor(x * 32768.0).clamp_(-32768, 32767).to(torch.int16))
d16 = torch.empty((n, 2), dtype=torch.int16, device="cuda")
P = "0123456789abcdef0011223344556677"
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def wall(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def up():
    with torch.cuda.stream(s1):
        d16.copy_(x16, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        y16.copy_(d16, non_blocking=True)


print("copy up %.2f ms, down %.2f ms, both %.2f ms" % (wall(up), wall(down), wall(lambda: (up(), down()))))
for piece in (sys.argv[1:] or ["2048", "4096", "6144", "8192", "12288", "24576", "1000000"]):
    os.environ["AWM_PIECE"] = piece
    print("AWM_PIECE=%-8s add_s16 %.2f ms" % (piece, wall(lambda: H.add_s16(x16.numpy(), P, None, y16.numpy()))))
