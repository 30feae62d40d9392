"""Small workload for `ncu --set full`: one add + one get of 10 minutes stereo (resident PCM), then one
`get --detect-speed` of the same audio played 1 % fast -- every kernel of the path launches at least once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
from audiowmark_b200 import hostapi as H

n = 10 * 60 * 44100
H.set_params()
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = (torch.rand((n, 2), device="cuda", generator=g, dtype=torch.float32) - 0.5)
y = torch.empty_like(x)
P = "0123456789abcdef0011223344556677"
H.add(x.data_ptr(), P, None, y.data_ptr(), n, 2)
doc = H.get(y.data_ptr(), n_frames=n, channels=2)
n_fast = int(np.rint(n / 1.01))
z = torch.empty((n_fast, 2), device="cuda", dtype=torch.float32)
H.resample(y.data_ptr(), 1 / 1.01, n_out=n_fast, out=z.data_ptr(), n_frames=n, channels=2)
H.set_speed_params(detect_speed=True)
doc2 = H.get(z.data_ptr(), n_frames=n_fast, channels=2)
H.set_speed_params()
print(len(doc["matches"]), len(doc2["matches"]), sum(m["type"].endswith("SPEED") for m in doc2["matches"]))
