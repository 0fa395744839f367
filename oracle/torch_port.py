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
