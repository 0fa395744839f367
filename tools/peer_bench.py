"""Cost of one BN-moments exchange kernel (vsiq_bn_moments_exchange) without peers: world = 1 on one buffer (fixed cost:
launch, fences, flag round trip), and world = 2 inside one process (two streams, buffers on the same GPU)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import _lib, ops
from vsiquantization_b200.parallel import PeerExchange
words = _lib.lib.vsiq_peer_buffer_bytes() // 8
C = int(sys.argv[1]) if len(sys.argv) > 1 else 576
stats = ops.observe(torch.randn(8, C, 20, 20, device="cuda"), ch_axis=1)
ms, vs = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
for world in (1, 2):
    bufs = [torch.zeros(words, dtype=torch.float64, device="cuda") for _ in range(world)]
    ranks = [PeerExchange(local_buffers=[b.data_ptr() for b in bufs], rank=r, timeout_s=5.0) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    n = 300
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    ranks[r].bn_moments(stats, 1.0, 3200.0 * world, ms, vs)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n * 1e6
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(streams[0]):
        ev[0].record()
        for i in range(n):
            ranks[0].bn_moments(stats, 1.0, 3200.0 * world, ms, vs) if world == 1 else None
        ev[1].record()
    torch.cuda.synchronize()
    print(f"world {world}: {dt:.1f} us per exchange (wall, {n} back-to-back)", f"; events world=1: {ev[0].elapsed_time(ev[1]) / n * 1e3:.1f} us" if world == 1 else "", ranks[0].status())
# the two launches the NCCL path keeps around its collective, for scale
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(300):
    ops.bn_moments_finalize(stats, 3200.0, ms, vs)
torch.cuda.synchronize(); print(f"bn_moments_finalize alone: {(time.perf_counter() - t0) / 300 * 1e6:.1f} us")
# one exchange at a time between events (GPU idle before each): the kernel's own duration + ~2 us of event overhead
bufs = [torch.zeros(words, dtype=torch.float64, device="cuda")]
px = PeerExchange(local_buffers=[bufs[0].data_ptr()], rank=0, timeout_s=5.0)
for name, fn in (("exchange world=1", lambda: px.bn_moments(stats, 1.0, 3200.0, ms, vs)),
                 ("bn_moments_finalize", lambda: ops.bn_moments_finalize(stats, 3200.0, ms, vs)),
                 ("empty-ish (fill 1 element)", lambda: ms[:1].fill_(0.0))):
    ts = []
    for i in range(40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"{name}: median {ts[20]:.1f} us, min {ts[0]:.1f} us (events around one launch)")
