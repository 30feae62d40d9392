"""Host-side phase timing of one `get` on a real GPU (development aid, not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
from audiowmark_b200 import hostapi as H
minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
n = int(minutes * 60 * 44100)
H.set_params()
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = (torch.rand((n, 2), device="cuda", generator=g, dtype=torch.float32) - 0.5)
y = torch.empty_like(x)
P = "0123456789abcdef0011223344556677"
for it in range(3):
    t0 = time.perf_counter(); H.add(x.data_ptr(), P, None, y.data_ptr(), n, 2); torch.cuda.synchronize(); t1 = time.perf_counter()
    if it == 2: os.environ["AWM_TRACE"] = "1"
    doc = H.get(y.data_ptr(), n_frames=n, channels=2); t2 = time.perf_counter()
    print("iter %d: add %.3f ms, get %.3f ms, %d matches" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(doc["matches"])), flush=True)
