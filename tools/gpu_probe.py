"""Quick per-stage timing on a real GPU (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import awm_oracle as O
import awm_testlib as T
from audiowmark_b200 import capi

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
P = O.Params(); KEY = O.Key()
ctx = capi.Context(0)
T.setup_ctx(ctx, KEY, P, T.PAYLOAD)
n = int(minutes * 60 * 44100)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = (torch.rand((n, 2), device="cuda", generator=g, dtype=torch.float32) - 0.5)
y = torch.empty_like(x)
def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize(); ctx.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ctx.synchronize(); ts.append(time.perf_counter() - t)
    print("%-28s %9.3f ms   (%.1f M PCM frames/s)" % (name, min(ts) * 1e3, n / min(ts) / 1e6), flush=True)
    return min(ts)
timed("embed (limiter)", lambda: ctx.embed(x.data_ptr(), y.data_ptr(), n_frames=n, channels=2))
timed("embed (no limiter)", lambda: ctx.embed(x.data_ptr(), y.data_ptr(), n_frames=n, channels=2, limiter_block=0))
ctx.embed(x.data_ptr(), y.data_ptr(), n_frames=n, channels=2); ctx.synchronize()
ctx.pcm_bind(y.data_ptr(), n, 2)
res = {}
def approx(): res["a"] = ctx.sync_approx(0, capi.MODE_BLOCK)
timed("sync_approx", approx)
a = res["a"]
aq = np.abs(a["raw_quality"] - a["local_mean"])
top = a[np.argsort(-aq)[:max(8, int((aq > 0.2625).sum() // 3))]]
print("candidates for refine:", len(top), "best", aq.max())
def refine(): res["r"] = ctx.sync_refine(top, 0, capi.MODE_BLOCK)
timed("sync_refine (%d cands)" % len(top), refine, reps=2)
r = res["r"]
idx = np.sort(r["index"])
def dec(): res["d"] = ctx.decode_blocks(idx, 858)
timed("decode_blocks (%d)" % len(idx), dec)
raw, valid = res["d"]
ok = raw[valid == 1]
def vit(): res["v"] = ctx.viterbi(ok, [0] * len(ok))
if len(ok):
    timed("viterbi (%d jobs, rate 6)" % len(ok), vit)
    bits, err = res["v"]
    print("payload hits:", sum(O.bit_vec_to_str(list(b)) == T.PAYLOAD for b in bits), "of", len(bits))
print("launches", ctx.launches)
