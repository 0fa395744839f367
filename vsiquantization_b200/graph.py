"""Whole-step CUDA-graph capture of a QAT training step.

Small models are launch-bound: YOLOv8n at batch 2 issues 228 fake-quant launches plus ~700 framework launches per step
for ~3 ms of GPU work.  Because this implementation never synchronises inside the step (qparams, grad-scales and
reductions all stay on the device; the reference reads ``.item()`` / ``.tolist()`` every step, yolov8_qat.py:248-258),
forward + backward + optimizer step can be captured once and replayed with a single launch.

    step = GraphedQATStep(model, optimizer, loss_fn, example_input)
    loss = step(batch)          # copies the batch into the static input, replays the graph, returns the loss tensor

Requirements: static shapes, model and qparams on the device, optimizer state created during the warm-up (done here).
The libvsiq.so launchers make no CUDA API call that is illegal during stream capture (device properties are cached
before the capture starts).

Data parallel (``torch.distributed`` initialised, world size > 1): pass the bare model, NOT a DistributedDataParallel
wrapper.  The step then reduces ALL gradients -- weights, biases and the LSQ step sizes / zero-points alike -- with ONE
NCCL all-reduce(SUM) over a flat buffer, captured inside the same graph between backward and the optimizer step (one
launch per rank per step, nothing on the host between them).  SUM matches the reference, which multiplies the loss by
the world size to undo DDP's averaging (yolov8_qat.py:235-236)."""
from __future__ import annotations

from typing import Callable

import torch


import os as _os

# measurement knob only (benchmarks: what does the step cost without its collective?) -- training without it is wrong
_SKIP_ALLREDUCE = _os.environ.get("VSIQ_DEBUG_SKIP_ALLREDUCE", "0") not in ("", "0")
_STEP_EVENTS = _os.environ.get("VSIQ_DEBUG_STEP_EVENTS", "0") not in ("", "0")


class GraphedQATStep:
    def __init__(self, model, optimizer, loss_fn: Callable, example_input: torch.Tensor, warmup: int = 3,
                 post_backward: Callable = None, group=None, average: bool = False):
        if not example_input.is_cuda:
            raise RuntimeError("GraphedQATStep needs a CUDA example input")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        import torch.distributed as dist
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1 and isinstance(model, torch.nn.parallel.DistributedDataParallel):
            raise ValueError("GraphedQATStep reduces the gradients itself: pass the bare model, not a DDP wrapper")
        self.static_input = example_input.clone()
        self.post_backward = post_backward
        self._marks = {}
        # timing events recorded INSIDE the captured graph (external event-record nodes): each replay re-records them, so
        # after a replay ``allreduce_ms()`` is the device time of that step's captured gradient all-reduce
        self._ar_events = None
        if self.world > 1:
            try:
                self._ar_events = (torch.cuda.Event(enable_timing=True, external=True),
                                   torch.cuda.Event(enable_timing=True, external=True))
            except TypeError:  # older torch: no external events, no in-graph timing
                self._ar_events = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 3)):  # allocator, cuDNN plans, optimizer state, workspaces
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._fwd_bwd_step()
        torch.cuda.synchronize()

    def _mark(self, name):
        """Debug timeline (VSIQ_DEBUG_STEP_EVENTS=1): external timing events captured into the graph."""
        if _STEP_EVENTS:
            ev = self._marks.setdefault(name, torch.cuda.Event(enable_timing=True, external=True))
            ev.record()

    def step_timeline_ms(self):
        """{phase: ms} of the last replayed step when VSIQ_DEBUG_STEP_EVENTS=1 (call after a synchronisation)."""
        names = [n for n in ("start", "forward", "backward", "packed", "reduced", "unpacked", "optimizer") if n in self._marks]
        return {b: float(self._marks[a].elapsed_time(self._marks[b])) for a, b in zip(names, names[1:])}

    def _fwd_bwd_step(self):
        self._mark("start")
        loss = self.loss_fn(self.model(self.static_input))
        self._mark("forward")
        loss.backward()
        self._mark("backward")
        if self.world > 1 and not _SKIP_ALLREDUCE:
            self._all_reduce_grads()
        if self.post_backward is not None:
            self.post_backward()
        self.optimizer.step()
        self._mark("optimizer")
        return loss.detach()

    def _all_reduce_grads(self) -> None:
        """One flat all-reduce per dtype over every gradient; the parameters' .grad become views of the flat buffer.

        The flat buffer is persistent and every gradient starts on a 32-byte boundary: the pack is ONE multi-tensor copy
        (torch._foreach_copy_) instead of torch.cat over 228 tensors (measured inside the graph: 0.38 ms -> the copy's
        bandwidth time), and the optimizer's multi-tensor kernels keep their vectorised path on the views (unaligned views
        of a cat'ed buffer cost the SGD step 0.33 -> 0.65 ms)."""
        import torch.distributed as dist
        by_dtype = {}
        for p in self.model.parameters():
            if p.grad is not None:
                by_dtype.setdefault(p.grad.dtype, []).append(p)
        self.allreduce_bytes = sum(p.grad.numel() * p.grad.element_size() for ps in by_dtype.values() for p in ps)
        if not hasattr(self, "_flat"):
            self._flat = {}
        packed = []
        for dtype, ps in by_dtype.items():
            align = max(1, 32 // ps[0].grad.element_size())
            sig = tuple(p.numel() for p in ps)
            ent = self._flat.get(dtype)
            if ent is None or ent[0] != sig or ent[1].device != ps[0].grad.device:
                offs, off = [], 0
                for p in ps:
                    offs.append(off)
                    off += (p.numel() + align - 1) // align * align
                ent = (sig, torch.zeros(off, dtype=dtype, device=ps[0].grad.device), offs)  # pads stay zero
                self._flat[dtype] = ent
            _, flat, offs = ent
            # each view keeps ITS gradient's strides (channels_last weights stay channels_last): the pack is then a plain
            # multi-tensor memcpy and the optimizer's multi-tensor kernels see parameters and gradients with equal strides
            # (a contiguous view of a channels_last gradient cost a permuting copy per tensor and pushed the foreach
            # optimizer onto its per-tensor path: measured 0.42 ms pack + 0.72 ms SGD step inside the graph)
            views = []
            for o, p in zip(offs, ps):
                g = p.grad
                dense = g.is_contiguous() or (g.dim() == 4 and g.is_contiguous(memory_format=torch.channels_last)) or \
                    (g.dim() == 5 and g.is_contiguous(memory_format=torch.channels_last_3d))
                views.append(torch.as_strided(flat, g.shape, g.stride(), o) if dense else flat[o:o + p.numel()].view(p.shape))
            torch._foreach_copy_(views, [p.grad for p in ps])
            packed.append((ps, flat, views))
        if hasattr(self, "_marks"):
            self._mark("packed")
        ev = getattr(self, "_ar_events", None)
        if ev is not None:
            ev[0].record()
        for _, flat, _ in packed:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        if ev is not None:
            ev[1].record()
        if hasattr(self, "_marks"):
            self._mark("reduced")
        for ps, flat, views in packed:
            if self.average:
                flat.div_(self.world)
            for p, v in zip(ps, views):
                p.grad = v
        if hasattr(self, "_marks"):
            self._mark("unpacked")

    def allreduce_ms(self):
        """Device time of the captured NCCL gradient all-reduce(s) in the last replayed step, or
        None (single process / no external-event support).  Call after a synchronisation."""
        if self._ar_events is None:
            return None
        try:
            return float(self._ar_events[0].elapsed_time(self._ar_events[1]))
        except (RuntimeError, ValueError):  # never recorded (single process, or the collective was skipped)
            return None

    def _eager_step(self):
        self.optimizer.zero_grad(set_to_none=True)
        return self._fwd_bwd_step()

    def __call__(self, batch: torch.Tensor) -> torch.Tensor:
        self.static_input.copy_(batch, non_blocking=True)
        self.graph.replay()
        return self.static_loss
