"""UniformQuantizer / LSQQuantizer: the reference's fake-quant plugins on the CUDA kernels.

Reference: quantizers/uniform.py:8-102 (UniformQuantizer), :242-271 (ScaleGradient, RoundStraightThrough),
:105-151 (FunLSQ, dead code -> ``mask_mode='funlsq'``), quantizers/lsq_module.py:147-173,254-274,317-358
(per-channel learnable scale / zero-point -> ``ch_axis``).  Same constructor ``(num_bits, symmetric)``, attributes
``num_bits, symmetric, qmin, qmax, calib_grad_scale`` and ``quantize(x, scale, zero_point, is_learning_scale)``.

The six eager forward kernels and ~14 backward kernels of the reference become ONE forward kernel and ONE backward
kernel (dx + per-channel dscale/dzero_point in the same pass); outputs are bit-identical to the reference on CPU.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib, ops
from ..utils.registry import register_class
from .base import BaseQuantizer


def _as_cuda(x: torch.Tensor) -> torch.Tensor:
    if x.is_cuda:
        return x
    if not torch.cuda.is_available():
        raise RuntimeError("vsiquantization_b200 needs a CUDA device: there is no CPU fallback")
    return x.cuda()  # differentiable: gradients flow back to the host tensor


@register_class
class UniformQuantizer(BaseQuantizer):
    """Uniform affine fake-quantiser with straight-through / LSQ gradients.

    Extension attributes (defaults reproduce the reference): ``ch_axis`` (None = per tensor; inferred from the
    shape of a multi-element scale), ``mask_mode`` ('rounded' = the reference's autograd semantics, 'funlsq' =
    quantizers/uniform.py:144-150), ``grad_boost`` (the x5000 activation hack of lsq_module.py:151-152; 1.0)."""

    def __init__(self, num_bits=8, symmetric=True, ch_axis: Optional[int] = None, mask_mode: str = "rounded",
                 grad_boost: float = 1.0):
        self.num_bits = num_bits
        self.symmetric = symmetric
        self.qmin = 0
        self.qmax = 2 ** self.num_bits - 1
        if self.symmetric:
            self.qmin = -(2 ** (self.num_bits - 1))
            self.qmax = 2 ** (self.num_bits - 1) - 1
        self.calib_grad_scale = 1
        self.ch_axis = ch_axis
        self.mask_mode = mask_mode
        self.grad_boost = grad_boost

    # -- helpers ------------------------------------------------------------------------------------
    def calculate_grad_scale(self, quant_tensor, channels: int = 1):
        """1/sqrt(Qp * numel) (uniform.py:58-71); per channel 1/sqrt(Qp * numel / C) (lsq_module.py:327-340)."""
        return ops.lsq_grad_scale(self.qmax, quant_tensor.numel(), channels)

    def _resolve_axis(self, x: torch.Tensor, scale) -> Optional[int]:
        if not isinstance(scale, torch.Tensor) or scale.numel() == 1:
            return None
        if self.ch_axis is not None:
            return self.ch_axis
        if scale.dim() == x.dim():  # broadcast shape such as [1, C, 1, 1]
            axes = [i for i, d in enumerate(scale.shape) if d != 1]
            if len(axes) == 1:
                return axes[0]
        raise ValueError("per-channel scale: set quantizer.ch_axis or pass a broadcast-shaped scale")

    supports_pre_relu = True  # quantize(..., pre_relu=True) fuses the preceding ReLU into the kernels
    supports_pre_silu = True  # quantize(..., pre_act="silu") fuses SiLU on channels_last tensors (ops.ci_supported)

    def _spec(self, ch_axis, zp_learned=False, pre_relu=False, pre_silu=False) -> ops.QSpec:
        mode = _lib.MASK_FUNLSQ if self.mask_mode == "funlsq" else _lib.MASK_ROUNDED
        return ops.QSpec(self.qmin, self.qmax, ch_axis=ch_axis, zp_learned=zp_learned, mask_mode=mode,
                         pre_relu=pre_relu, pre_silu=pre_silu)

    # -- the plugin entry point ------------------------------------------------------------------------
    def kernel_args(self, x, scale, zero_point, is_learning_scale=False):
        """(scale, zero_point, QSpec) exactly as quantize() hands them to the forward kernel, or None when this
        quantiser cannot ride along as the SECOND stage of a two-output epilogue over the channels_last tensor x."""
        if not ops.ci_supported(x):
            return None
        ch_axis = self._resolve_axis(x, scale)
        if ch_axis not in (None, 1):
            return None
        zp_tensor = isinstance(zero_point, torch.Tensor)
        zp_round = zp_tensor and is_learning_scale and not self.symmetric and zero_point.is_floating_point()
        s = scale.detach() if isinstance(scale, torch.Tensor) else scale
        z = zero_point.detach() if zp_tensor else zero_point
        return s, z, self._spec(ch_axis, zp_learned=zp_round)

    def quantize(self, x, scale, zero_point, is_learning_scale=False, pre_relu=False, bias=None, pre_act=None,
                 second=None, precomputed=None):
        """Fake-quantise x (uniform.py:34-56); with ``pre_relu`` quantise relu(x) in the same pass (the fused layer's
        F.relu, modules/fused.py:133, folded into the kernel; gradients include relu's mask).

        scale: Python int/float, np.float64, or a tensor / nn.Parameter with 1 or C entries (fp32 or the reference's
        0-dim fp64, on any device).  zero_point: Python int, or a float tensor / nn.Parameter (learnable: the forward
        uses clamp(round(z)), uniform.py:98-102).  The result is autograd-connected to x and to every qparam that
        requires grad."""
        if pre_act not in (None, "relu", "silu"):
            raise ValueError("pre_act must be None, 'relu' or 'silu'")
        pre_relu = bool(pre_relu) or pre_act == "relu"
        pre_silu = pre_act == "silu"
        if not x.is_cuda:
            if second is not None or precomputed is not None:
                raise ValueError("two-output epilogue / precomputed results are for CUDA tensors")
            y = self.quantize(_as_cuda(x), scale, zero_point, is_learning_scale, pre_relu, bias, "silu" if pre_silu else None)
            return y.to(x.device)
        ch_axis = self._resolve_axis(x, scale)
        if bias is not None or pre_silu or second is not None:
            # SiLU (x / (1 + exp(-x)), the fused layer's F.silu of modules/fused.py:133) exists in the channels_last
            # epilogue kernels only; ``second`` = (scale2, zero_point2, spec2, sink): the same pass also writes the next
            # layer's quantize_inp result (fake_quantize.py:44-45) into sink
            if precomputed is not None:
                raise ValueError("a precomputed result replaces a plain fake-quant launch, not an epilogue")
            return self._quantize_epilogue(x, bias, scale, zero_point, is_learning_scale, pre_relu, ch_axis, pre_silu,
                                           second=second)
        scale_learn = isinstance(scale, torch.Tensor) and scale.requires_grad and torch.is_grad_enabled()
        zp_tensor = isinstance(zero_point, torch.Tensor)
        # the reference rounds / clamps a tensor zero-point only on the asymmetric learning path (uniform.py:50-52)
        zp_round = zp_tensor and is_learning_scale and not self.symmetric and zero_point.is_floating_point()
        zp_learn = zp_round and zero_point.requires_grad and torch.is_grad_enabled()
        if not (scale_learn or zp_learn):
            s = scale.detach() if isinstance(scale, torch.Tensor) else scale
            z = zero_point.detach() if zp_tensor else zero_point
            return ops.FakeQuantFixed.apply(x, s, z, self._spec(ch_axis, zp_learned=zp_round, pre_relu=pre_relu),
                                            precomputed)
        if not isinstance(scale, torch.Tensor):
            raise TypeError("a learnable zero_point needs a tensor scale")
        C = scale.numel()
        gs_host, gs_dev = 1.0, None
        if is_learning_scale:  # ScaleGradient (uniform.py:48-49); otherwise plain autograd, factor 1
            gs_host = self.calculate_grad_scale(x, C) * float(self.grad_boost)
            cgs = self.calib_grad_scale
            if isinstance(cgs, torch.Tensor):
                # a [C] calib_grad_scale (utils/estimate_bn.py:136) is sum-reduced onto the scale by autograd
                gs_dev = cgs.detach().to(device=x.device, dtype=torch.float32).sum().reshape(1)
            else:
                gs_host *= float(cgs)
        zp_arg = zero_point if zp_learn else (zero_point.detach() if zp_tensor else zero_point)
        return ops.FakeQuantLearned.apply(x, scale, zp_arg, self._spec(ch_axis, zp_learned=zp_round, pre_relu=pre_relu),
                                          gs_host, gs_dev, precomputed)

    def _quantize_epilogue(self, x, bias, scale, zero_point, is_learning_scale, pre_relu, ch_axis, pre_silu=False,
                           second=None):
        """fq(act(x + bias)) on a channels_last conv output (ops.FakeQuantEpilogue); bias gradient from the same pass."""
        if not ops.ci_supported(x):
            raise ValueError("bias / SiLU fusion needs a channels_last float32 CUDA tensor with C % 4 == 0 and C <= 1024")
        if self.mask_mode == "funlsq":
            raise ValueError("mask_mode='funlsq' has no fused-bias form")
        zp_tensor = isinstance(zero_point, torch.Tensor)
        zp_round = zp_tensor and is_learning_scale and not self.symmetric and zero_point.is_floating_point()
        grad_on = torch.is_grad_enabled()
        scale_learn = isinstance(scale, torch.Tensor) and scale.requires_grad and grad_on
        zp_learn = zp_round and zero_point.requires_grad and grad_on
        gs_host, gs_dev = 1.0, None
        if is_learning_scale and isinstance(scale, torch.Tensor):
            gs_host = self.calculate_grad_scale(x, scale.numel()) * float(self.grad_boost)
            cgs = self.calib_grad_scale
            if isinstance(cgs, torch.Tensor):
                gs_dev = cgs.detach().to(device=x.device, dtype=torch.float32).sum().reshape(1)
            else:
                gs_host *= float(cgs)
        s_arg = scale if scale_learn else (scale.detach() if isinstance(scale, torch.Tensor) else scale)
        z_arg = zero_point if zp_learn else (zero_point.detach() if zp_tensor else zero_point)
        spec = self._spec(ch_axis, zp_learned=zp_round, pre_relu=pre_relu, pre_silu=pre_silu)
        return ops.FakeQuantEpilogue.apply(x, bias, s_arg, z_arg, spec, gs_host, gs_dev, second)

    def quantize_codes(self, x, scale, zero_point, code_bits: Optional[int] = None):
        """(fake-quantised tensor, integer codes) -- the reference keeps codes as floats (uniform.py:54).  int8 / uint8
        when the range fits 8 bits, int16 / uint16 otherwise; ``code_bits=4`` packs two codes per byte (W4 deployment)."""
        x = _as_cuda(x)
        ch_axis = self._resolve_axis(x, scale)
        zp_round = isinstance(zero_point, torch.Tensor) and not self.symmetric and zero_point.is_floating_point() \
            and zero_point.requires_grad
        s = scale.detach() if isinstance(scale, torch.Tensor) else scale
        z = zero_point.detach() if isinstance(zero_point, torch.Tensor) else zero_point
        return ops.quantize_codes(x.detach(), s, z, self._spec(ch_axis, zp_round), code_bits)


@register_class
class LSQQuantizer(UniformQuantizer):
    """Learned-step-size quantiser.  Named by the reference (README.md:70-71, modules/fuse_config.py:183) but never
    defined there: ``CLASS_REGISTRY['LSQQuantizer']`` is a KeyError in the reference.  Here it is the learnable path
    of UniformQuantizer (uniform.py:47-52) -- which is what actually runs in the reference -- with the per-channel
    form of quantizers/lsq_module.py available through ``ch_axis``; identical arithmetic, same kernels."""
