"""Where does a step of the query-sharded C2 pass go?  (torchrun --nproc-per-node N tools/qshard_probe.py)"""
import importlib, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
hp = importlib.import_module("humanpose-and-urbanscene-matching_b200")
E, R = hp.engine, hp.retrieval
n = 1 << 20
q, t = hp.synth.sift_like(n, n, seed=1, planted=0.1)
qa, qb = R.shard_range(n, rank, world)
dq = torch.from_numpy(q[qa:qb]).to(dev).contiguous()
dt = torch.from_numpy(t).to(dev)
del q, t
ws = E.knn2_l2_workspace(qb - qa, n, 128, E.L2_AUTO, dev)
idx = torch.empty((qb - qa, 2), dtype=torch.int32, device=dev)
dst = torch.empty((qb - qa, 2), dtype=torch.float32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for it in range(6):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    ev[0].record()
    E.knn2_l2(dq, dt, mode=E.L2_AUTO, out=(idx, dst), workspace=ws)
    w1 = time.perf_counter()
    ev[1].record()
    per = -(-n // world)
    packed = torch.full((per, 4), -1, dtype=torch.int32, device=dev)
    packed[:idx.shape[0], :2] = idx
    packed[:idx.shape[0], 2:] = dst.contiguous().view(torch.int32)
    ev[2].record()
    g = R.allgather(packed).reshape(world * per, 4)[:n] if world > 1 else packed
    ev[3].record()
    gi, gd = g[:, :2].contiguous(), g[:, 2:].contiguous().view(torch.float32)
    keep, count = E.ratio_filter(gi, gd, 0.8)
    ev[4].record()
    torch.cuda.synchronize()
    w2 = time.perf_counter()
    if it >= 3:
        print("rank %d it %d: knn %.2f ms (host returned after %.2f) pack %.2f allgather %.2f ratio %.2f | wall %.2f ms | mode %d" %
              (rank, it, ev[0].elapsed_time(ev[1]), (w1 - w0) * 1e3, ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]),
               ev[3].elapsed_time(ev[4]), (w2 - w0) * 1e3, E.knn2_l2_last_mode()), flush=True)
if world > 1:
    dist.destroy_process_group()
