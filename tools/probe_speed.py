"""Timing of the speed detection path on one GPU (development aid): detect_speed alone and `get --detect-speed`
on a 30 minute stereo chunk played 1% fast, with the per-kernel profile; the reference binary (oracle/_ref) is timed on
the same input when present."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch
from audiowmark_b200 import hostapi as H

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
n = int(minutes * 60 * 44100)
H.set_params()
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = (torch.rand((n, 2), device="cuda", generator=g, dtype=torch.float32) - 0.5)
y = torch.empty_like(x)
H.add(x.data_ptr(), "0123456789abcdef0011223344556677", None, y.data_ptr(), n, 2); H.synchronize()
n_fast = int(np.rint(n / 1.01))
z = torch.empty((n_fast, 2), device="cuda", dtype=torch.float32)
H.resample(y.data_ptr(), 1 / 1.01, n_out=n_fast, out=z.data_ptr(), n_frames=n, channels=2); H.synchronize()
zh = z.cpu().numpy()
pin = torch.from_numpy(zh).pin_memory().numpy()
for dev, buf, kw in (("host", pin, {}), ("device", z.data_ptr(), dict(n_frames=n_fast, channels=2))):
    for it in range(3):
        H.set_speed_params(detect_speed=True)
        t0 = time.perf_counter(); r = H.detect_speed(buf, **kw); t1 = time.perf_counter()
        doc = H.get(buf, **kw); t2 = time.perf_counter()
        H.set_speed_params()
        doc1 = H.get(buf, **kw); t3 = time.perf_counter()
        print("%s it %d: detect_speed %.2f ms -> %s; get --detect-speed %.2f ms (%d matches, %d SPEED); plain get %.2f ms" % (
            dev, it, (t1 - t0) * 1e3, r, (t2 - t1) * 1e3, len(doc["matches"]), sum(m["type"].endswith("SPEED") for m in doc["matches"]), (t3 - t2) * 1e3), flush=True)
H.profile_enable(True)
H.set_speed_params(detect_speed=True)
H.get(pin)
H.set_speed_params()
print(json.dumps(H.profile_report()))
ref = os.path.join(ROOT, "oracle", "_ref", "audiowmark")
if os.path.exists(ref) and "--ref" in sys.argv:
    import awm_oracle as O
    O.write_wav16("/tmp/probe_speed.wav", zh)
    t0 = time.perf_counter()
    p = subprocess.run([ref, "get", "--detect-speed", "/tmp/probe_speed.wav"], capture_output=True, text=True)
    print("reference get --detect-speed: %.2f s, %d cores" % (time.perf_counter() - t0, os.cpu_count()))
    print(p.stdout[:300])
