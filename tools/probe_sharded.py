"""cProfile of the frame-balanced sharded get on rank 0 (development aid; run under torchrun)."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import torch.distributed as dist
from audiowmark_b200 import hostapi as H, sharding as S

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
n = 60 * 60 * 44100
n_total = n * world
H.set_params(gpu_device=local)
mx, ov = H.chunk_geometry(44100)
plan = S.chunk_plan(n_total, mx, ov, 44100)
sl = S.rank_slices(plan, rank, world, n_total)
e0, e1, ffn = S.embed_range(min(s.lo for s in sl), max(s.hi for s in sl), n_total, 44100)
n_loc = e1 - e0
g = torch.Generator(device=dev); g.manual_seed(1 + rank)
x = torch.rand((n_loc, 2), device=dev, generator=g, dtype=torch.float32) - 0.5
y = torch.empty_like(x)
P = "0123456789abcdef0011223344556677"
def step():
    H.add(x.data_ptr(), P, None, y.data_ptr(), n_loc, 2, first_frame_number=ffn)
    job = S.BalancedGet(rank, world, n_total, y.data_ptr(), e0, n_loc, 2)
    return job.run(lambda payload: S.allgather_bytes(payload, device=dev))
for _ in range(3):
    step()
dist.barrier(); torch.cuda.synchronize()
pr = cProfile.Profile() if rank == 0 else None
t0 = time.perf_counter()
if pr: pr.enable()
for _ in range(10):
    step()
if pr: pr.disable()
torch.cuda.synchronize()
if rank == 0:
    print("10 steps: %.2f ms per step" % ((time.perf_counter() - t0) * 100), flush=True)
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue()[:9000], flush=True)
dist.destroy_process_group()
