"""Data-parallel plumbing of the QAT path: the two small exchanges the path really has.

QAT shards as pure data parallel (batch shards per rank, replicated weights); fake-quant forward/backward need no
communication.  What must be exchanged:

  * calibration: every rank sees different batches, so observer extrema are all-reduced -- ONE packed
    all_reduce(MIN) over [min, -max] of all quantisers (exact, so post-calibration scales are bit-identical to a
    single process that saw the union of the batches), plus one all_reduce(SUM) of the LSQ-initialisation statistics;
    scale / zero-point are then recomputed on the device for all observers in one launch (vsiq_qparams_from_minmax).
    The reference has no such step: every rank calibrates on identical un-sharded data (yolov8_qat.py:86-92).
  * training: the LSQ dscale / dzero_point of all layers live in ONE flat buffer per dtype (their .grad are views into
    it) and are reduced with a single all_reduce(SUM) instead of riding in DDP's 25 MB buckets with the weights.

torch.distributed (NCCL over NVLink on the box, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

STATE_WIDTH = 8  # run_min, run_max, scale, zero_point, n_calls, sum mean|x|, sum mean x, sum std


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def quantization_managers(model) -> List[Tuple[str, torch.nn.Module]]:
    """(qualified name, manager) of every weight_quantizer / activation_quantizer, in module order."""
    out = []
    for name, module in model.named_modules():
        for attr in ("weight_quantizer", "activation_quantizer"):
            if hasattr(module, attr):
                out.append((f"{name}.{attr}" if name else attr, getattr(module, attr)))
    return out


def reduce_observer_states(states: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce a packed [rows, 8] fp64 observer-state block in place: MIN over run_min, MAX over run_max (as MIN of
    the negation, one collective for both), SUM over the call statistics.  Device-agnostic (CPU/gloo or CUDA/NCCL)."""
    if states.dtype != torch.float64 or states.dim() != 2 or states.shape[1] != STATE_WIDTH:
        raise ValueError("states must be [rows, 8] float64")
    if _world(group) == 1:
        return states
    packed = torch.cat([states[:, 0], -states[:, 1]]).contiguous()
    dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    n = states.shape[0]
    states[:, 0] = packed[:n]
    states[:, 1] = -packed[n:]
    sums = states[:, 4:8].contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    states[:, 4:8] = sums
    return states


def sync_observers(model, group=None) -> int:
    """Make every rank's observers agree after sharded calibration; returns the number of observer rows synchronised.
    Scales / zero-points are recomputed from the reduced extrema exactly as observers/minmax.py:67-74 would."""
    from . import ops
    every = [(m, m.observer) for _, m in quantization_managers(model) if m.observer.state is not None]
    # moving-average observers (observers/moving_average.py) hold averages, not extrema: their running min / max are
    # AVERAGED over the ranks and their qparams recomputed with torch's formula; the call statistics are summed
    ema = [(m, o) for m, o in every if hasattr(o, "averaging_constant")]
    entries = [(m, o) for m, o in every if not hasattr(o, "averaging_constant")]
    rows_ema = 0
    if ema:
        from .observers.moving_average import torch_qparams
        arena = torch.cat([o.state for _, o in ema]).contiguous()
        w = _world(group)
        if w > 1:
            # ranks that never observed (n_calls == 0) carry no information: weight by "has data"
            has = (arena[:, 4:5] > 0).to(arena.dtype)
            packed = torch.cat([arena[:, 0:2] * has, has, arena[:, 4:8]], dim=1).contiguous()
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            n = torch.clamp(packed[:, 2:3], min=1.0)
            arena[:, 0:2] = (packed[:, 0:2] / n).to(torch.float32).to(arena.dtype)
            arena[:, 4:8] = packed[:, 3:7]
        row = 0
        for m, o in ema:
            k = o.state.shape[0]
            blk = arena[row:row + k]
            s_, z_ = torch_qparams(blk[:, 0].to(torch.float32), blk[:, 1].to(torch.float32), o.quant_min, o.quant_max,
                                   bool(o.symmetric), float(o.eps))
            blk[:, 2], blk[:, 3] = s_.to(arena.dtype), z_.to(arena.dtype)
            o.state.copy_(blk)
            o._host = None
            m._invalidate()
            row += k
        rows_ema = row
    if not entries:
        return rows_ema
    arena = torch.cat([o.state for _, o in entries]).contiguous()
    reduce_observer_states(arena, group)
    bits = torch.cat([torch.full((o.state.shape[0],), int(o.num_bits), dtype=torch.int32) for _, o in entries])
    sym = torch.cat([torch.full((o.state.shape[0],), int(bool(o.symmetric)), dtype=torch.int32) for _, o in entries])
    eps = {float(o.eps) for _, o in entries}
    if len(eps) != 1:
        raise ValueError("observers with different eps cannot be recomputed in one launch")
    ops.qparams_from_minmax(arena, bits, sym, eps.pop())
    row = 0
    for m, o in entries:
        k = o.state.shape[0]
        o.state.copy_(arena[row:row + k])
        o._host = None
        m._invalidate()
        row += k
    return row + rows_ema


class QParamGradBucket:
    """Flat gradient buffers for the learnable quantisation parameters (…quantizer.scale / …quantizer.zero_point).

    ``p.grad`` of every such parameter becomes a view into one flat tensor per dtype (autograd then accumulates in
    place), ``all_reduce()`` reduces each flat tensor with ONE collective, and ``ddp_ignore()`` keeps DDP from
    reducing the same parameters a second time.  Use ``zero()`` instead of ``optimizer.zero_grad(set_to_none=True)``.
    ``average=True`` divides by the world size (DDP's convention; the reference then multiplies the loss by the world
    size, yolov8_qat.py:235-236)."""

    SUFFIXES = ("quantizer.scale", "quantizer.zero_point")

    def __init__(self, model, group=None, average: bool = True):
        self.group, self.average = group, average
        self.names: List[str] = []
        self.params: List[torch.nn.Parameter] = []
        for name, p in model.named_parameters():
            if name.endswith(self.SUFFIXES) and p.requires_grad:
                self.names.append(name)
                self.params.append(p)
        self.flat: Dict[Tuple[torch.dtype, torch.device], torch.Tensor] = {}
        by_key: Dict[Tuple[torch.dtype, torch.device], List[torch.nn.Parameter]] = {}
        for p in self.params:
            by_key.setdefault((p.dtype, p.device), []).append(p)
        for key, ps in by_key.items():
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=key[0], device=key[1])
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            self.flat[key] = flat

    def __len__(self):
        return len(self.params)

    def numel(self) -> int:
        return sum(f.numel() for f in self.flat.values())

    def zero(self) -> None:
        for f in self.flat.values():
            f.zero_()

    def all_reduce(self) -> None:
        w = _world(self.group)
        if w == 1:
            return
        for f in self.flat.values():
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                f.div_(w)

    def ddp_ignore(self, model) -> None:
        """Tell DistributedDataParallel (call BEFORE wrapping) not to reduce these parameters itself."""
        from torch.nn.parallel import DistributedDataParallel as DDP
        DDP._set_params_and_buffers_to_ignore_for_model(model, list(self.names))
