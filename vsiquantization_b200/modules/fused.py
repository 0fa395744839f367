"""Fused QAT layers: ConvBnReLU, ConvBn, ConvReLU, Conv, LinearBnReLU, LinearBn, LinearReLU, Linear.

Reference: modules/fused.py:32-412.  Same positional constructor order
``(layers..., observer_w, quantizer_w, observer_a, quantizer_a, w_symmetric, a_symmetric, [is_fuse_bn], bits_w, bits_a)``
and attributes (conv_fuse / linear_fuse, is_fuse_bn, is_relu, bn, weight_quantizer, activation_quantizer, bits_w, bits_a,
quantize_out, quantize_inp).  The construction-time BN fold (fused.py:98-108, :292-300) runs on the bn_fold CUDA kernel
(bit-identical); conv / linear stay on cuDNN / cuBLAS.  The reference's Linear / LinearReLU / bias-less LinearBnReLU do
not run (``bool(tensor)``, missing get_weight_bias -- SURVEY.md 0.9); they are fixed here."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..quantizers.fake_quantize import FakeQuantize


def _activation(x, relu_flag: Optional[bool]):
    if relu_flag is None:
        return x
    return F.relu(x) if relu_flag else F.silu(x)


def _fold_bn_into(core: nn.Module, weight: torch.Tensor, bias: Optional[torch.Tensor], bn: nn.Module) -> None:
    """W' = W*gamma/sqrt(var+eps), b' = beta + (b-mean)*gamma/sqrt(var+eps) written into ``core`` (kernel 5)."""
    if not torch.cuda.is_available():
        raise RuntimeError("vsiquantization_b200 needs a CUDA device to fold BatchNorm: there is no CPU fallback")
    home = weight.device
    dev = home if home.type == "cuda" else torch.device("cuda")
    to = lambda t: None if t is None else t.detach().to(dev, torch.float32)  # noqa: E731
    Wf, bf, _, _ = ops.bn_fold(to(weight), to(bias), to(bn.weight), to(bn.bias), to(bn.running_mean),
                               to(bn.running_var), bn.eps)
    with torch.no_grad():
        core.weight.copy_(Wf.to(home))
        if core.bias is not None:
            core.bias.copy_(bf.to(home))


class _ConvBase(FakeQuantize):
    def _make_conv(self, cv: nn.Conv2d, bias: bool) -> None:
        self.conv_fuse = nn.Conv2d(cv.in_channels, cv.out_channels, cv.kernel_size, cv.stride, cv.padding,
                                   cv.dilation, cv.groups, bias=bias).to(cv.weight.device)
        with torch.no_grad():
            self.conv_fuse.weight.copy_(cv.weight)
            if self.conv_fuse.bias is not None:
                if cv.bias is not None:
                    self.conv_fuse.bias.copy_(cv.bias)
                else:
                    self.conv_fuse.bias.zero_()

    def _conv(self, x, weights, bias):
        c = self.conv_fuse
        return F.conv2d(x, weights, bias, stride=c.stride, padding=c.padding, dilation=c.dilation, groups=c.groups)

    @staticmethod
    def _dense_input(x):
        """A channel slice of a channels_last tensor (``chunk`` / ``split`` along dim 1, as in C2f blocks: stride(1) == 1
        but pitched rows) is made dense ONCE, here.  cuDNN needs a dense tensor, so ATen would copy it inside the
        convolution anyway -- in the forward and AGAIN in the backward, because autograd saves the view -- and the layer
        would fall off the channels_last epilogue (separate bias-add pass, separate bias-gradient reduction, separate
        observer pass while calibrating).  Same values, so every result is unchanged."""
        if (x.dim() == 4 and x.is_cuda and x.shape[1] > 1 and x.stride(1) == 1 and not x.is_contiguous()
                and not x.is_contiguous(memory_format=torch.channels_last)):
            return x.contiguous(memory_format=torch.channels_last)
        return x


class ConvBnReLU(_ConvBase):
    """Conv2d + BatchNorm2d + ReLU/SiLU (fused.py:32-134).  ``relu`` may be nn.ReLU or nn.SiLU."""

    def __init__(self, cv, bn, relu, observer_w_name: str, quantizer_w_name: str, observer_a_name: str,
                 quantizer_a_name: str, w_symmetric: bool = True, a_symmetric: bool = True, is_fuse_bn=True,
                 bits_w: int = 8, bits_a: int = 8, **ext):
        super().__init__(observer_w_name, quantizer_w_name, observer_a_name, quantizer_a_name, w_symmetric, a_symmetric,
                         bits_w, bits_a, **ext)
        self._make_conv(cv, bias=bool(is_fuse_bn or cv.bias is not None))
        self.is_fuse_bn = is_fuse_bn
        self.is_relu = isinstance(relu, nn.ReLU)  # False means SiLU (fused.py:81)
        self._has_act = relu is not None
        if is_fuse_bn:
            _fold_bn_into(self.conv_fuse, cv.weight, cv.bias, bn)
        else:
            self.bn = bn
        self._bn_reestimate = None  # set by utils.estimate_bn while re-estimating

    fuse_relu_into_quant = True  # ReLU + output fake-quant as one kernel pass (same values, one read/write less)
    fuse_bias_into_quant = True  # channels_last only: conv bias add + its gradient ride in the quantiser kernels

    def _bn(self, x):
        if self._bn_reestimate is not None:
            return self._bn_reestimate(self, x)
        return self.bn(x)

    def _pre_activation(self, x, weights, bias):
        x = self._conv(x, weights, bias)
        if not self.is_fuse_bn:
            x = self._bn(x)
        return x

    fuse_eval_bn = True  # inference-mode BN + ReLU of a layer that kept its BN as ONE NHWC pass (no autograd: no_grad only)

    def run_forward_core(self, x, weights, bias):
        if not self.is_fuse_bn and self._bn_reestimate is not None:
            # BN re-estimation: the hook normalises with the batch moments and applies the activation in the same pass
            return self._bn_reestimate(self, self._conv(x, weights, bias), act="relu" if self.is_relu else "silu")
        if (self.fuse_eval_bn and not self.is_fuse_bn and self.is_relu and not torch.is_grad_enabled()
                and not self.bn.training and self.bn.running_mean is not None):
            # calibration / evaluation of an unfused layer (fused.py:131-134: bn, then relu): ATen spends two read + write
            # passes on it, vsiq_ci_bn_normalize one (x * a[c] + b[c] from the running moments, ReLU folded in)
            y = self._conv(x, weights, bias)
            if ops.ci_supported(y):
                return ops.ci_bn_normalize(y, self.bn.running_mean, self.bn.running_var, self.bn.weight, self.bn.bias,
                                           self.bn.eps, relu=True)
            return F.relu(self.bn(y))
        x = self._pre_activation(x, weights, bias)
        # the reference applies SiLU when relu is not an nn.ReLU -- including relu=None (ConvBn overrides this)
        return F.relu(x) if self.is_relu else F.silu(x)

    fuse_observer_into_epilogue = True  # calibration: bias / BN + activation + output observer as ONE NHWC pass

    def _calibration_forward(self, x, act):
        """The calibration forward (calibrate_qat_model: observers on, quantisation off) of a layer on
        channels_last memory: conv, then ONE pass that adds the bias (or applies the inference-mode BN the layer kept),
        applies the activation, writes the result and feeds the output observer -- instead of bias / BN + activation
        passes followed by an observer pass (fused.py:124-134, quantization_manager.py:55-71).  None = not applicable
        (nothing has been touched yet): the caller runs the ordinary forward."""
        aq = self.activation_quantizer
        collecting = (not aq.is_learning_scale) and aq.is_observer_qparam
        obs = aq.observer
        if (not collecting or aq.is_quantize or not hasattr(obs, "observe_epilogue")
                or getattr(obs, "ch_axis", None) is not None or x.dim() != 4 or not x.is_cuda or x.is_contiguous()
                or not x.is_contiguous(memory_format=torch.channels_last)):
            return None
        reestimating = self._bn_reestimate is not None and not self.is_fuse_bn
        if self._bn_reestimate is not None and self.is_fuse_bn:
            return None
        if not self.is_fuse_bn and not reestimating and (self.bn.training or self.bn.running_mean is None):
            return None
        if self.quantize_inp:
            x = self.quantize_input(x)
        weights, bias = self.get_weight_bias()
        weights = self.quantize_weights(weights)
        if reestimating:
            # reestimate_BN_stats while the output observer is still collecting: batch moments (one read), then
            # normalise + activation + observer as one pass (utils/estimate_bn.py hook)
            return self._bn_reestimate(self, self._conv(x, weights, bias), act=act, collect=aq)
        fn = F.relu if act == "relu" else F.silu
        if self.is_fuse_bn:
            pre = self._conv(x, weights, None)  # the epilogue adds the bias, as in the training step
            y = aq.collect_epilogue(pre, act, bias=bias)
            if y is None:
                y = aq.quantize(fn(pre if bias is None else pre + bias.view(1, -1, 1, 1)))
            return y
        pre = self._conv(x, weights, bias)
        bn = self.bn
        y = aq.collect_epilogue(pre, act, bn=(bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.eps))
        if y is None:
            y = aq.quantize(fn(bn(pre)))
        return y

    def forward(self, x):
        """fake_quantize.py:43-51 with the activation folded into the output quantiser when that is possible: ReLU in
        every layout, SiLU (what the reference applies whenever ``relu`` is not an nn.ReLU, fused.py:81,133) on
        channels_last tensors."""
        act = "relu" if self.is_relu else "silu"
        x = self._dense_input(x)
        if (self.fuse_observer_into_epilogue and self._has_act and self.quantize_out
                and type(self).run_forward_core is ConvBnReLU.run_forward_core):
            y = self._calibration_forward(x, act)
            if y is not None:
                return y
        if not (self.fuse_relu_into_quant and self._has_act and self.quantize_out
                and type(self).run_forward_core is ConvBnReLU.run_forward_core
                and self.activation_quantizer.can_fuse_relu()
                and not (self._bn_reestimate is not None and not self.is_fuse_bn)):
            return super().forward(x)
        if self.quantize_inp:
            x = self.quantize_input(x)
        weights, bias = self.get_weight_bias()
        weights = self.quantize_weights(weights)
        consumer = self.__dict__.get("_inp_consumer")
        if consumer is not None and not consumer.quantize_inp:
            consumer = None
        if (self.fuse_bias_into_quant and bias is not None and self.is_fuse_bn and x.dim() == 4 and not x.is_contiguous()
                and x.is_contiguous(memory_format=torch.channels_last)):
            pre = self._conv(x, weights, None)  # bias-free conv; the epilogue adds the bias and returns its gradient
            if ops.ci_supported(pre):
                return self._quantize_out(pre, act, bias, consumer)
            return self.activation_quantizer.quantize(pre + bias.view(1, -1, 1, 1), pre_act=act)
        pre = self._pre_activation(x, weights, bias)
        if consumer is not None and ops.ci_supported(pre):
            return self._quantize_out(pre, act, None, consumer)
        return self.activation_quantizer.quantize(pre, pre_act=act)

    def _quantize_out(self, pre, act, bias, consumer):
        """Output epilogue over a channels_last conv result; with a linked ``quantize_inp`` consumer
        (feed_input_quantizer_of) the same pass also writes that layer's input quantisation."""
        plan = consumer.activation_quantizer.prequant_plan(pre) if consumer is not None else None
        if plan is None or not self.activation_quantizer.can_emit_second(act, pre):
            return self.activation_quantizer.quantize(pre, pre_act=act, bias=bias)
        sink: list = []
        y = self.activation_quantizer.quantize(pre, pre_act=act, bias=bias, second=plan + (sink,))
        if sink:
            self._offer_prequant(y, sink[0], consumer)
        return y


class ConvBn(ConvBnReLU):
    """Conv2d + BatchNorm2d, no activation (fused.py:137-156)."""

    def __init__(self, cv, bn, observer_w_name: str, quantizer_w_name: str, observer_a_name: str, quantizer_a_name: str,
                 w_symmetric: bool = True, a_symmetric: bool = True, is_fuse_bn=True, bits_w: int = 8, bits_a: int = 8,
                 **ext):
        super().__init__(cv, bn, None, observer_w_name, quantizer_w_name, observer_a_name, quantizer_a_name, w_symmetric,
                         a_symmetric, is_fuse_bn, bits_w, bits_a, **ext)

    def run_forward_core(self, x, weights, bias):
        x = self._conv(x, weights, bias)
        if not self.is_fuse_bn:
            x = self._bn(x)
        return x


class ConvReLU(_ConvBase):
    """Conv2d + ReLU/SiLU (fused.py:159-203)."""

    def __init__(self, cv, relu, observer_w_name: str, quantizer_w_name: str, observer_a_name: str, quantizer_a_name: str,
                 w_symmetric: bool = True, a_symmetric: bool = True, bits_w: int = 8, bits_a: int = 8, **ext):
        super().__init__(observer_w_name, quantizer_w_name, observer_a_name, quantizer_a_name, w_symmetric, a_symmetric,
                         bits_w, bits_a, **ext)
        self._make_conv(cv, bias=cv.bias is not None)
        self.is_relu = isinstance(relu, nn.ReLU)

    def run_forward_core(self, x, weights, bias):
        return _activation(self._conv(x, weights, bias), self.is_relu)


class Conv(_ConvBase):
    """Quantisation-aware Conv2d (fused.py:206-248)."""

    def __init__(self, cv, observer_w_name: str, quantizer_w_name: str, observer_a_name: str, quantizer_a_name: str,
                 w_symmetric: bool = True, a_symmetric: bool = True, bits_w: int = 8, bits_a: int = 8, **ext):
        super().__init__(observer_w_name, quantizer_w_name, observer_a_name, quantizer_a_name, w_symmetric, a_symmetric,
                         bits_w, bits_a, **ext)
        self._make_conv(cv, bias=cv.bias is not None)

    def run_forward_core(self, x, weights, bias):
        return self._conv(x, weights, bias)


class _LinearBase(FakeQuantize):
    def _make_linear(self, linear: nn.Linear, bias: bool) -> None:
        self.linear_fuse = nn.Linear(linear.in_features, linear.out_features, bias=bias).to(linear.weight.device)
        with torch.no_grad():
            self.linear_fuse.weight.copy_(linear.weight)
            if self.linear_fuse.bias is not None:
                if linear.bias is not None:
                    self.linear_fuse.bias.copy_(linear.bias)
                else:
                    self.linear_fuse.bias.zero_()


class LinearBnReLU(_LinearBase):
    """Linear + BatchNorm1d + ReLU/SiLU (fused.py:251-317); also accepts a bias-less Linear (the reference crashes)."""

    def __init__(self, linear, bn, relu, observer_w_name: str, quantizer_w_name: str, observer_a_name: str,
                 quantizer_a_name: str, w_symmetric: bool = True, a_symmetric: bool = True, is_fuse_bn: bool = True,
                 bits_w: int = 8, bits_a: int = 8, **ext):
        super().__init__(observer_w_name, quantizer_w_name, observer_a_name, quantizer_a_name, w_symmetric, a_symmetric,
                         bits_w, bits_a, **ext)
        self._make_linear(linear, bias=bool(is_fuse_bn or linear.bias is not None))
        self.is_fuse_bn = is_fuse_bn
        self.is_relu = isinstance(relu, nn.ReLU)
        if is_fuse_bn:
            _fold_bn_into(self.linear_fuse, linear.weight, linear.bias, bn)
        else:
            self.bn = bn

    def run_forward_core(self, x, weights, bias):
        x = F.linear(x, weights, bias)
        if not self.is_fuse_bn:
            x = self.bn(x)
        return F.relu(x) if self.is_relu else F.silu(x)


class LinearBn(LinearBnReLU):
    """Linear + BatchNorm1d (fused.py:320-344)."""

    def __init__(self, linear, bn, observer_w_name: str, quantizer_w_name: str, observer_a_name: str,
                 quantizer_a_name: str, w_symmetric: bool = True, a_symmetric: bool = True, is_fuse_bn: bool = True,
                 bits_w: int = 8, bits_a: int = 8, **ext):
        super().__init__(linear, bn, None, observer_w_name, quantizer_w_name, observer_a_name, quantizer_a_name,
                         w_symmetric, a_symmetric, is_fuse_bn, bits_w, bits_a, **ext)

    def run_forward_core(self, x, weights, bias):
        x = F.linear(x, weights, bias)
        if not self.is_fuse_bn:
            x = self.bn(x)
        return x


class LinearReLU(_LinearBase):
    """Linear + ReLU/SiLU (fused.py:347-379)."""

    def __init__(self, linear, relu, observer_w_name: str, quantizer_w_name: str, observer_a_name: str,
                 quantizer_a_name: str, w_symmetric: bool = True, a_symmetric: bool = True, bits_w: int = 8,
                 bits_a: int = 8, **ext):
        super().__init__(observer_w_name, quantizer_w_name, observer_a_name, quantizer_a_name, w_symmetric, a_symmetric,
                         bits_w, bits_a, **ext)
        self._make_linear(linear, bias=linear.bias is not None)
        self.is_relu = isinstance(relu, nn.ReLU)

    def run_forward_core(self, x, weights, bias):
        return _activation(F.linear(x, weights, bias), self.is_relu)


class Linear(_LinearBase):
    """Quantisation-aware Linear (fused.py:382-412)."""

    def __init__(self, linear, observer_w_name: str, quantizer_w_name: str, observer_a_name: str, quantizer_a_name: str,
                 w_symmetric: bool = True, a_symmetric: bool = True, bits_w: int = 8, bits_a: int = 8, **ext):
        super().__init__(observer_w_name, quantizer_w_name, observer_a_name, quantizer_a_name, w_symmetric, a_symmetric,
                         bits_w, bits_a, **ext)
        self._make_linear(linear, bias=linear.bias is not None)

    def run_forward_core(self, x, weights, bias):
        return F.linear(x, weights, bias)


FUSED_CLASSES = (ConvBnReLU, ConvBn, ConvReLU, Conv, LinearBnReLU, LinearBn, LinearReLU, Linear)
