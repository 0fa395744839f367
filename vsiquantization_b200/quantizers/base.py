"""Quantizer interface (reference: quantizers/base.py:4-32)."""
from abc import ABC, abstractmethod


class BaseQuantizer(ABC):
    """Maps a float tensor onto a uniform integer grid and back (fake quantisation)."""

    @abstractmethod
    def quantize(self, x, scale, zero_point, is_learning_scale=False):
        ...
