"""FakeQuantize: base of every fused QAT layer (reference: quantizers/fake_quantize.py:8-69).

Template-method forward: [quantise input] -> get_weight_bias -> quantise weights -> run_forward_core -> [quantise output].
Owns two QuantizationManagers (``weight_quantizer``, ``activation_quantizer``) -- the attributes the control API
(utils/quantize_manager.py) discovers layers by."""
from __future__ import annotations

from typing import Optional

import torch.nn as nn

from .quantization_manager import QuantizationManager


class FakeQuantize(nn.Module):
    def __init__(self, observer_w_name: str, quantizer_w_name: str, observer_a_name: str, quantizer_a_name: str,
                 w_symmetric: bool = True, a_symmetric: bool = True, bits_w: int = 4, bits_a: int = 8,
                 quantize_out: bool = True, quantize_inp: bool = False, w_ch_axis: Optional[int] = None,
                 a_ch_axis: Optional[int] = None):
        super().__init__()
        self.weight_quantizer = QuantizationManager(quantizer_w_name, observer_w_name, bits_w, w_symmetric,
                                                    is_learning_scale=True, ch_axis=w_ch_axis)
        self.activation_quantizer = QuantizationManager(quantizer_a_name, observer_a_name, bits_a, a_symmetric,
                                                        is_learning_scale=True, ch_axis=a_ch_axis)
        self.bits_w = bits_w
        self.bits_a = bits_a
        self.quantize_out = quantize_out
        self.quantize_inp = quantize_inp

    def forward(self, x):
        if self.quantize_inp:
            x = self.quantize_activation(x)
        weights, bias = self.get_weight_bias()
        weights = self.quantize_weights(weights)
        out = self.run_forward_core(x, weights, bias)
        if self.quantize_out:
            out = self.quantize_activation(out)
        return out

    def run_forward_core(self, x, weights, bias):
        raise NotImplementedError

    def _core(self):
        core = getattr(self, "conv_fuse", None)
        return core if core is not None else self.linear_fuse

    def get_weight_bias(self):
        core = self._core()
        return core.weight, core.bias  # the bias is not quantised (fake_quantize.py:56-60)

    def quantize_weights(self, weights):
        return self.weight_quantizer.quantize(weights)

    def quantize_activation(self, out):
        return self.activation_quantizer.quantize(out)
