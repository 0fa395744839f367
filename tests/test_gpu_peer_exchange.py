"""BN re-estimation's per-layer exchange as one kernel over peer memory (csrc/peer_exchange.cu; reference semantics:
utils/estimate_bn.py:56-99 over the union of every rank's shard).  The protocol is exercised inside one
process first (two "ranks" = two buffers and two streams of one GPU), then across two processes over cudaIpc / NVLink when
the box has two GPUs."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _buffers(n):
    from vsiquantization_b200 import _lib
    words = _lib.lib.vsiq_peer_buffer_bytes() // 8
    return [torch.zeros(words, dtype=torch.float64, device="cuda") for _ in range(n)]


@pytest.mark.parametrize("world", [1, 2, 4])
def test_exchange_inside_one_process_equals_sum_then_finalize(world):
    from vsiquantization_b200 import ops
    from vsiquantization_b200.parallel import PeerExchange
    torch.manual_seed(1)
    bufs = _buffers(world)
    ranks = [PeerExchange(local_buffers=[b.data_ptr() for b in bufs], rank=r, timeout_s=5.0) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    C, rows = 96, 5000
    for it in range(5):       # several exchanges: slot parity, monotonic sequence numbers
        shards = [torch.randn(8, C, 25, 25, device="cuda") * (1 + r) + 0.1 * it for r in range(world)]
        stats = [ops.observe(s, ch_axis=1) for s in shards]
        weights = [1.0] * world if it != 3 else [1.0] + [0.0] * (world - 1)     # a rank replaying a batch with weight zero
        count = float(sum(w * rows for w in weights))
        msum = [torch.zeros(C, device="cuda") + r for r in range(world)]
        vsum = [torch.zeros(C, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        outs = []
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                outs.append(ranks[r].bn_moments(stats[r], weights[r], count, msum[r], vsum[r]))
        torch.cuda.synchronize()
        total = torch.zeros_like(stats[0])
        for r in range(world):
            total = total + weights[r] * stats[r]
        m, vb, vu = ops.bn_moments_finalize(total, count)
        for r in range(world):
            assert torch.equal(outs[r][0], m) and torch.equal(outs[r][1], vb) and torch.equal(outs[r][2], vu), (it, r)
            assert torch.equal(msum[r], m + r) and torch.equal(vsum[r], vu)
    for r in range(world):
        assert ranks[r].status() == (5, 0)


def test_a_missing_peer_times_out_instead_of_hanging():
    from vsiquantization_b200 import ops
    from vsiquantization_b200.parallel import PeerExchange
    bufs = _buffers(2)
    lonely = PeerExchange(local_buffers=[b.data_ptr() for b in bufs], rank=0, timeout_s=0.2)
    stats = ops.observe(torch.randn(4, 8, 6, 6, device="cuda"), ch_axis=1)
    mean = lonely.bn_moments(stats, 1.0, 144.0)[0]
    torch.cuda.synchronize()
    assert lonely.status() == (0, 1)      # nothing completed, exchange 1 reported as failed
    del mean
    import time
    t0 = time.perf_counter()
    lonely.bn_moments(stats, 1.0, 144.0)  # after a failure every later exchange returns at once (no second wait)
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 0.1 and lonely.status() == (0, 1)


def _worker(rank, world, port, results):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from vsiquantization_b200 import ops, parallel
    from vsiquantization_b200.modules.fused import ConvBnReLU
    from vsiquantization_b200.utils.estimate_bn import reestimate_BN_stats
    torch.manual_seed(0)

    def build():
        torch.manual_seed(0)
        cv, bn = torch.nn.Conv2d(8, 32, 3, padding=1, bias=False), torch.nn.BatchNorm2d(32, eps=1e-3)
        m = ConvBnReLU(cv, bn, torch.nn.ReLU(), "MinMaxObserver", "UniformQuantizer", "MinMaxObserver", "UniformQuantizer",
                       True, True, False, 8, 8).cuda().to(memory_format=torch.channels_last)
        for q in (m.weight_quantizer, m.activation_quantizer):
            q.is_quantize = False
            q.is_observer_qparam = False
        return m
    g = torch.Generator().manual_seed(100 + rank)
    # uneven shards: rank 0 holds 3 batches, rank 1 two (the third exchange sees a replayed batch with weight zero)
    batches = [(torch.randint(0, 256, (4 + rank, 8, 16, 16), generator=g, dtype=torch.uint8),) for _ in range(3 - rank)]
    out = {}
    for mode in ("1", "0"):
        os.environ["VSIQ_PEER_EXCHANGE"] = mode
        parallel._peer_exchanges.clear()
        m = build()
        reestimate_BN_stats(m, batches, num_batches=3)
        out[mode] = (m.bn.running_mean.cpu(), m.bn.running_var.cpu())
        if mode == "1":
            px = parallel.peer_exchange_for(None)
            out["used"] = px is not None and px.status() == (3, 0)
    os.environ.pop("VSIQ_PEER_EXCHANGE", None)
    results[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (peer memory over NVLink)")
def test_two_processes_over_peer_memory_match_the_nccl_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(2):
        assert results[r]["used"], "the peer-memory exchange was not taken"
        pm, pv = results[r]["1"]
        nm, nv = results[r]["0"]
        assert torch.allclose(pm, nm, rtol=1e-6, atol=1e-7) and torch.allclose(pv, nv, rtol=1e-6, atol=1e-7)
    # every rank ends with bit-identical statistics (same summation order everywhere)
    assert torch.equal(results[0]["1"][0], results[1]["1"][0]) and torch.equal(results[0]["1"][1], results[1]["1"][1])
