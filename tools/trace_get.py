"""AWM_TRACE=1 python tools/trace_get.py: stage timeline (host wall clock) of one resident 1 h add + get, after warm-up"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from audiowmark_b200 import hostapi as H

n = 60 * 60 * 44100
H.set_params()
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.rand((n, 2), device="cuda", generator=g, dtype=torch.float32) - 0.5
y = torch.empty_like(x)
P = "0123456789abcdef0011223344556677"
os.environ.pop("AWM_TRACE", None)
for _ in range(2):
    H.add(x.data_ptr(), P, None, y.data_ptr(), n, 2)
    H.get(y.data_ptr(), n_frames=n, channels=2)
os.environ["AWM_TRACE"] = "1"
t0 = time.perf_counter()
H.add(x.data_ptr(), P, None, y.data_ptr(), n, 2)
H.synchronize()
t1 = time.perf_counter()
doc = H.get(y.data_ptr(), n_frames=n, channels=2)
t2 = time.perf_counter()
print("add %.3f ms, get %.3f ms, %d patterns" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, len(doc["matches"])), file=sys.stderr)
