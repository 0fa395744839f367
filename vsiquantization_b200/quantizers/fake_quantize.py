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
            x = self.quantize_input(x)
        weights, bias = self.get_weight_bias()
        weights = self.quantize_weights(weights)
        out = self.run_forward_core(x, weights, bias)
        if self.quantize_out:
            out = self.quantize_activation(out)
        return out

    def quantize_input(self, x):
        """The ``quantize_inp`` step (fake_quantize.py:44-45).  When the layer that produced x was linked to this one
        (feed_input_quantizer_of) its epilogue has already written this result in the same pass as its own output; the
        tensor is picked up here instead of launching a fake-quant over x."""
        pre = self.__dict__.get("_prequant")
        if pre is not None:
            self.__dict__["_prequant"] = None
            if pre[0]() is x:
                return pre[1]
        return self.quantize_activation(x)

    def feed_input_quantizer_of(self, consumer: Optional["FakeQuantize"]) -> None:
        """Link this layer to the (unique or not) layer that consumes its output with ``quantize_inp=True``: this
        layer's output epilogue then also writes the consumer's input quantisation (two outputs from one pass over the
        conv result: 12 bytes per element instead of 8 + 8, one launch less).  None removes the link.  Results and
        gradients are those of the unlinked model bit for bit; whenever the fused form does not apply (observing,
        quantisers off, NCHW tensors) both layers fall back to their own launches."""
        if consumer is not None and not hasattr(consumer, "activation_quantizer"):
            raise TypeError("the consumer must be a fused QAT layer")
        self.__dict__["_inp_consumer"] = consumer

    def _offer_prequant(self, y, y2_raw, consumer) -> None:
        import weakref
        y2 = consumer.activation_quantizer.quantize_precomputed(y, y2_raw)
        consumer.__dict__["_prequant"] = (weakref.ref(y), y2)

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
