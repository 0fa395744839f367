"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: observer MIN/MAX/SUM exchange and the flat qparam-gradient
bucket.  The collectives are device-agnostic torch.distributed code; the GPU-only step (recomputing scales from the
reduced extrema) is checked here against the oracle's formula."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(fn, world=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    return [q.get() for _ in range(world)]


def _entry(fn, rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put(fn(rank, world))
    finally:
        dist.destroy_process_group()


def _observer_case(rank, world):
    from vsiquantization_b200.parallel import reduce_observer_states
    rng = np.random.default_rng(7)
    batches = [rng.standard_normal((4, 50)).astype(np.float32) * (1 + i) for i in range(4)]  # same on every rank
    mine = batches[rank::world]  # calibration batches sharded across ranks
    states = np.zeros((3, 8))
    states[:, 2] = 1.0
    for row in range(3):  # three quantisers looking at different slices
        run_min = run_max = 0.0
        for b in mine:
            st = oracle.minmax_stats(b[:, row * 10:(row + 1) * 10 + 20])
            run_min, run_max = oracle.minmax_update(run_min, run_max, st[0, 0], st[0, 1])
            n = b[:, row * 10:(row + 1) * 10 + 20].size
            states[row, 4] += 1
            states[row, 5] += st[0, 2] / n
            states[row, 6] += st[0, 3] / n
        states[row, 0], states[row, 1] = run_min, run_max
    t = torch.tensor(states, dtype=torch.float64)
    reduce_observer_states(t)
    return t.numpy()


def test_observer_allreduce_equals_single_process_union():
    res = _run(_observer_case)
    assert np.array_equal(res[0], res[1])  # all ranks agree
    rng = np.random.default_rng(7)
    batches = [rng.standard_normal((4, 50)).astype(np.float32) * (1 + i) for i in range(4)]
    for row in range(3):
        run_min = run_max = 0.0
        mean_abs = []
        for b in batches:
            st = oracle.minmax_stats(b[:, row * 10:(row + 1) * 10 + 20])
            run_min, run_max = oracle.minmax_update(run_min, run_max, st[0, 0], st[0, 1])
            mean_abs.append(st[0, 2] / b[:, row * 10:(row + 1) * 10 + 20].size)
        # MIN/MAX are exact: extrema -- and therefore the scales derived from them -- are bit-identical
        assert (res[0][row, 0], res[0][row, 1]) == (run_min, run_max)
        assert oracle.qparams(res[0][row, 0], res[0][row, 1], 8, True) == oracle.qparams(run_min, run_max, 8, True)
        assert oracle.qparams(res[0][row, 0], res[0][row, 1], 4, False) == oracle.qparams(run_min, run_max, 4, False)
        assert res[0][row, 4] == 4  # calls summed over ranks
        # LSQ init from summed statistics: order-dependent sum, tolerance
        assert 2 * res[0][row, 5] / res[0][row, 4] == pytest.approx(2 * np.mean(mean_abs), rel=1e-12)


class _Q(torch.nn.Module):
    def __init__(self, asym):
        super().__init__()
        self.scale = torch.nn.Parameter(torch.tensor(0.05, dtype=torch.float64))
        if asym:
            self.zero_point = torch.nn.Parameter(torch.tensor(3.0))


class _Layer(torch.nn.Module):
    def __init__(self, asym):
        super().__init__()
        self.weight_quantizer = _Q(False)
        self.activation_quantizer = _Q(asym)
        self.lin = torch.nn.Linear(4, 4)


def _bucket_case(rank, world):
    from vsiquantization_b200.parallel import QParamGradBucket
    torch.manual_seed(0)
    model = torch.nn.Sequential(_Layer(False), _Layer(True))
    bucket = QParamGradBucket(model, average=True)
    assert len(bucket) == 5 and bucket.numel() == 5 and len(bucket.flat) == 2  # 4 fp64 scales, 1 fp32 zero-point
    out = []
    for step in range(2):
        bucket.zero()
        loss = sum((p * (rank + 1 + step)).sum() for p in bucket.params)  # d/dp = rank + 1 + step
        loss.backward()
        for p in bucket.params:  # autograd accumulated in place: .grad is still a view of the flat buffer
            assert p.grad.data_ptr() in {f[i:].data_ptr() for f in bucket.flat.values() for i in range(f.numel())}
        bucket.all_reduce()
        out.append([float(p.grad) for p in bucket.params])
    names = list(bucket.names)
    bucket.ddp_ignore(model)
    ignored = sorted(getattr(model, "_ddp_params_and_buffers_to_ignore"))
    return out, names, ignored


def test_qparam_grad_bucket_single_collective_per_dtype():
    res = _run(_bucket_case)
    (g0, names0, ign0), (g1, _, _) = res
    assert g0 == g1
    assert g0[0] == [1.5] * 5 and g0[1] == [2.5] * 5  # mean over ranks of (rank + 1 + step)
    assert all(n.endswith(("quantizer.scale", "quantizer.zero_point")) for n in names0)
    assert ign0 == sorted(names0)


def _flat_grad_case(rank, world):
    """GraphedQATStep's gradient exchange (one flat all-reduce per dtype) without the graph: gloo, CPU tensors."""
    from vsiquantization_b200.graph import GraphedQATStep
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3)).double()
    model.register_parameter("scale64", torch.nn.Parameter(torch.tensor(0.5, dtype=torch.float64)))
    model.register_parameter("zp32", torch.nn.Parameter(torch.tensor([1.0, 2.0], dtype=torch.float32)))
    # a channels_last 4-D weight: its gradient keeps channels_last strides, and so must the flat view handed back
    wcl = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).reshape(2, 3, 2, 2).contiguous(memory_format=torch.channels_last)
    model.register_parameter("wcl", torch.nn.Parameter(wcl))
    x = torch.full((2, 5), float(rank + 1), dtype=torch.float64)
    loss = (model(x) ** 2).sum() * model.scale64 + (model.zp32 ** 2).sum() * (rank + 1) + (model.wcl ** 2).sum() * (rank + 2)
    loss.backward()
    assert model.wcl.grad.stride() == model.wcl.stride()
    local = [p.grad.clone() for p in model.parameters()]
    step = GraphedQATStep.__new__(GraphedQATStep)  # the exchange only: no CUDA, no capture
    step.model, step.group, step.average, step.world = model, None, False, world
    step._all_reduce_grads()
    assert model.wcl.grad.stride() == model.wcl.stride()          # strides preserved: the optimizer's fast path
    flat32 = step._flat[torch.float32][1]
    assert all(p.grad.data_ptr() % 32 == 0 for p in model.parameters())  # every view starts on a 32-byte boundary
    assert model.wcl.grad.untyped_storage().data_ptr() == flat32.untyped_storage().data_ptr()
    step._all_reduce_grads()                                       # second call reuses the persistent buffers
    assert step._flat[torch.float32][1] is flat32
    return [g.numpy() for g in local], [p.grad.numpy().copy() / world for p in model.parameters()]


def test_graphed_step_flat_gradient_all_reduce_is_a_sum_over_ranks():
    res = _run(_flat_grad_case)
    locals_, reduced = zip(*res)
    for k in range(len(reduced[0])):
        want = locals_[0][k] + locals_[1][k]
        for r in range(2):
            assert reduced[r][k].dtype == want.dtype and reduced[r][k].shape == want.shape
            np.testing.assert_allclose(reduced[r][k], want, rtol=1e-12)


def _ema_sync_case(rank, world):
    """sync_observers on moving-average observers: running extrema are AVERAGED over the ranks that observed data."""
    from vsiquantization_b200.observers.moving_average import ema_update_, torch_qparams
    from vsiquantization_b200.parallel import sync_observers
    from vsiquantization_b200.quantizers.quantization_manager import QuantizationManager as M

    class Layer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.weight_quantizer = M("UniformQuantizer", "MovingAveragePerChannelMinMaxObserver", 8, True)
            self.activation_quantizer = M("UniformQuantizer", "MovingAverageMinMaxObserver", 8, False)
    layer = Layer()
    g = torch.Generator().manual_seed(10 + rank)
    for q, C in ((layer.weight_quantizer, 4), (layer.activation_quantizer, 1)):
        st = torch.zeros(C, 8, dtype=torch.float64)
        st[:, 2] = 1.0
        if not (rank == 1 and C == 1):  # rank 1 never observed an activation: it must not drag the average to zero
            for _ in range(3):
                x = torch.randn(C, 50, generator=g) * (rank + 1)
                ema_update_(st, x.min(1).values, x.max(1).values, 0.01, q.observer.quant_min, q.observer.quant_max,
                            bool(q.observer.symmetric))
        q.observer.load_state(st)
        q._calibrated = True
    before = [q.observer.state.clone() for q in (layer.weight_quantizer, layer.activation_quantizer)]
    rows = sync_observers(layer)
    after = [q.observer.state.clone() for q in (layer.weight_quantizer, layer.activation_quantizer)]
    w_obs = layer.weight_quantizer.observer
    s, z = torch_qparams(after[0][:, 0].float(), after[0][:, 1].float(), w_obs.quant_min, w_obs.quant_max, True)
    assert torch.equal(after[0][:, 2].float(), s) and torch.equal(after[0][:, 3].long(), z)
    return rank, rows, [b.numpy() for b in before], [a.numpy() for a in after]


def test_sync_observers_averages_moving_average_observers():
    res = sorted(_run(_ema_sync_case), key=lambda r: r[0])  # the queue hands results back in arrival order
    (_, rows0, b0, a0), (_, rows1, b1, a1) = res
    assert rows0 == rows1 == 5
    for k in range(2):
        assert np.array_equal(a0[k], a1[k])  # every rank ends with the same state
    # weights: both ranks observed -> plain mean of the running extrema (rounded to fp32); call counts summed
    want = ((b0[0][:, :2] + b1[0][:, :2]) / 2).astype(np.float32).astype(np.float64)
    assert np.array_equal(a0[0][:, :2], want) and np.array_equal(a0[0][:, 4], b0[0][:, 4] + b1[0][:, 4])
    # activations: only rank 0 has data -> its extrema survive unchanged
    assert np.array_equal(a0[1][:, :2], b0[1][:, :2]) and a0[1][0, 4] == 3


def _peer_case(rank, world):
    """Under gloo (no CUDA peers) the per-layer exchange of BN re-estimation must keep its collective: peer_exchange_for
    answers None on every rank, and VSIQ_PEER_EXCHANGE=0 switches it off wherever it would apply."""
    from vsiquantization_b200 import parallel
    parallel._peer_exchanges.clear()
    first = parallel.peer_exchange_for(None)
    os.environ["VSIQ_PEER_EXCHANGE"] = "0"
    second = parallel.peer_exchange_for(None)
    os.environ.pop("VSIQ_PEER_EXCHANGE")
    parallel.drop_peer_exchange(None)
    return first is None and second is None and parallel.peer_exchange_for(None) is None


def test_peer_exchange_declines_without_cuda_peers():
    assert _run(_peer_case) == [True, True]
