"""Observer interface (reference: observers/base.py:5-41)."""
from abc import ABC, abstractmethod


class BaseObserver(ABC):
    """Collects statistics of the tensors it is shown and turns them into (scale, zero_point)."""

    @abstractmethod
    def observe(self, x):
        ...

    @abstractmethod
    def get_scale_zero_point(self):
        ...
