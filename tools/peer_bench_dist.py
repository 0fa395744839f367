"""Back-to-back BN-moment exchanges between the ranks of one node, nothing else on the GPUs: the peer-memory kernel against
NCCL all_reduce + moments kernel.      torchrun --nproc-per-node N tools/peer_bench_dist.py [C]"""
import os, sys, time, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
from vsiquantization_b200 import ops
from vsiquantization_b200.parallel import peer_exchange_for
C = int(sys.argv[1]) if len(sys.argv) > 1 else 576
stats = ops.observe(torch.randn(8, C, 20, 20, device="cuda"), ch_axis=1)
ms, vs = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
px = peer_exchange_for(None)
assert px is not None, "no peer memory between the ranks"
n = 500


def peer():
    px.bn_moments(stats, 1.0, 3200.0 * world, ms, vs)


def nccl():
    s = stats.clone()
    dist.all_reduce(s)
    ops.bn_moments_finalize(s, 3200.0 * world, ms, vs)


def nccl_only():
    dist.all_reduce(stats)


for name, fn in (("peer-memory exchange kernel", peer), ("NCCL all_reduce + moments kernel", nccl), ("NCCL all_reduce alone", nccl_only)):
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n * 1e6
    # GPU-side: capture 50 calls in a CUDA graph (no CPU launch cost) where the call can be captured
    g_us = None
    if fn is peer:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            peer(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for i in range(50):
                    peer()
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); g.replay(); g.replay(); b.record(st)
            torch.cuda.synchronize()
            g_us = a.elapsed_time(b) * 1e3 / 100
    if rank == 0:
        print(f"{world} GPUs, C={C}: {name}: {dt:.1f} us per exchange (wall, CPU-launched)" + (f"; {g_us:.1f} us inside a CUDA graph" if g_us else ""), flush=True)
print(rank, px.status()) if rank == 0 else None
dist.barrier()
dist.destroy_process_group()
