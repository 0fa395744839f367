"""Torch-eager restatement of the reference's fake-quant op sequence -- TEST / BASELINE INFRASTRUCTURE ONLY.

This is what the reference executes on the host CPU for the hot path: separate ATen elementwise kernels composed in
Python, autograd for the backward (no fusion).  bench.py times it as the CPU baseline / ``--impl reference`` arm
(kind "port": /root/reference itself is Python and cannot travel to the GPU box).  The product never imports this.

Restated from (file:line in the reference):
  quantizers/uniform.py:95      x_int = clamp(RoundStraightThrough(x / scale + zero_point), qmin, qmax)
  quantizers/uniform.py:55      x_dequant = (x_int - zero_point) * scale
  quantizers/uniform.py:258-271 RoundStraightThrough: forward torch.round, backward identity
  quantizers/uniform.py:242-255 ScaleGradient: forward identity, backward grad * scale
  quantizers/uniform.py:69-71   grad_scale = (qmax * numel) ** -0.5
Checked bit-for-bit against the golden vectors in tests/test_oracle_golden.py::test_torch_port_matches_golden.
"""
import torch


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _ScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        ctx.k = k
        return x

    @staticmethod
    def backward(ctx, g):
        return g * ctx.k, None


def fake_quant(x, scale, zero_point, qmin, qmax, learn=False):
    if learn:
        scale = _ScaleGrad.apply(scale, (qmax * x.numel()) ** -0.5)
    x_int = torch.clamp(_RoundSTE.apply(x / scale + zero_point), qmin, qmax)
    return (x_int - zero_point) * scale


def fwd_bwd(x, g, scale, zero_point, qmin, qmax, learn=False):
    """One forward + backward of the reference composition; returns (y, dx[, ds])."""
    x = x.detach().requires_grad_(True)
    if learn:
        scale = torch.nn.Parameter(torch.as_tensor(scale, dtype=torch.float64))
    y = fake_quant(x, scale, zero_point, qmin, qmax, learn)
    y.backward(g)
    return (y.detach(), x.grad, scale.grad) if learn else (y.detach(), x.grad)


class EagerQuantizer:
    """The reference's quantizer as a plugin object (torch eager on whatever device the tensors live on): per tensor it
    is quantizers/uniform.py:34-56 verbatim in behaviour; with a multi-element scale it is the per-channel form of
    quantizers/lsq_module.py:147-173,254-274,317-358 (qparams broadcast along ``ch_axis``, grad-scale
    (qmax * numel / C) ** -0.5, learnable zero-point rounded and clamped with a straight-through gradient).
    Used as the same-GPU checker of the model-scale parity tests and as benchmarks/yolo_qat.py's ``--quant-impl eager``
    yardstick.  A float64 scale is rounded to float32 before use, like ATen does with the reference's 0-dim float64
    Parameter."""

    def __init__(self, num_bits=8, symmetric=True, ch_axis=None, grad_boost=1.0):
        self.num_bits, self.symmetric = num_bits, symmetric
        self.qmin, self.qmax = (-(2 ** (num_bits - 1)), 2 ** (num_bits - 1) - 1) if symmetric else (0, 2 ** num_bits - 1)
        self.calib_grad_scale = 1
        self.ch_axis = ch_axis
        self.grad_boost = grad_boost

    @classmethod
    def like(cls, q):
        e = cls(q.num_bits, q.symmetric, getattr(q, "ch_axis", None), getattr(q, "grad_boost", 1.0))
        e.calib_grad_scale = q.calib_grad_scale
        return e

    def quantize(self, x, scale, zero_point, is_learning_scale=False):
        C = scale.numel() if isinstance(scale, torch.Tensor) else 1
        shape = [1] * x.dim()
        if C > 1:
            shape[self.ch_axis] = C
        if isinstance(scale, torch.Tensor):
            scale = scale.to(torch.float32).reshape(shape if C > 1 else ())
        if isinstance(zero_point, torch.Tensor):
            zero_point = zero_point.to(torch.float32).reshape(shape if C > 1 else ())
        if is_learning_scale and isinstance(scale, torch.Tensor):
            gs = (self.qmax * (x.numel() / C)) ** -0.5 * float(self.grad_boost)
            cgs = self.calib_grad_scale
            gs = gs * (cgs.sum() if isinstance(cgs, torch.Tensor) else cgs)
            scale = _ScaleGrad.apply(scale, gs)
            if not self.symmetric and isinstance(zero_point, torch.Tensor) and zero_point.is_floating_point():
                zero_point = torch.clamp(_RoundSTE.apply(zero_point), self.qmin, self.qmax)
                zero_point = _ScaleGrad.apply(zero_point, gs)
        x_int = torch.clamp(_RoundSTE.apply(x / scale + zero_point), self.qmin, self.qmax)
        return (x_int - zero_point) * scale
