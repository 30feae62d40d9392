"""Stage timing of the frame-balanced get on one GPU (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from audiowmark_b200 import hostapi as H, sharding as S
n = int(60 * 60 * 44100)
H.set_params()
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = (torch.rand((n, 2), device="cuda", generator=g, dtype=torch.float32) - 0.5)
y = torch.empty_like(x)
P = "0123456789abcdef0011223344556677"
torch.cuda.synchronize()
H.add(x.data_ptr(), P, None, y.data_ptr(), n, 2); H.synchronize()
for it in range(3):
    t = [time.perf_counter()]
    job = S.BalancedGet(0, 1, n, y.data_ptr(), 0, n, 2); t.append(time.perf_counter())
    p1 = [job.stage_peaks()]; t.append(time.perf_counter())
    job.stage_select(p1); t.append(time.perf_counter())
    p2 = [job.stage_refine()]; t.append(time.perf_counter())
    job.stage_final(p2); t.append(time.perf_counter())
    p3 = [job.stage_decode()]; t.append(time.perf_counter())
    p4 = [job.stage_viterbi(p3)]; t.append(time.perf_counter())
    doc = job.stage_merge(p4); t.append(time.perf_counter())
    names = ["init", "peaks", "select", "refine", "final", "decode", "viterbi", "merge"]
    print("iter", it, " ".join("%s %.2f" % (nm, (b - a) * 1e3) for nm, a, b in zip(names, t, t[1:])), "total %.2f ms" % ((t[-1] - t[0]) * 1e3), len(doc["matches"]), flush=True)
t0 = time.perf_counter(); doc2 = H.get(y.data_ptr(), n_frames=n, channels=2); print("H.get %.2f ms" % ((time.perf_counter() - t0) * 1e3), doc2 == doc)
