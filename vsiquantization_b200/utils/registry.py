"""Plugin registry -- same contract as the reference's utils/registry.py:2-27: a global dict keyed by
``cls.__name__`` and a decorator that fills it; later registrations overwrite earlier ones, unknown names
are a plain ``KeyError`` at lookup time (quantizers/quantization_manager.py:41-42)."""
from typing import Dict

CLASS_REGISTRY: Dict[str, type] = {}


def register_class(cls):
    CLASS_REGISTRY[cls.__name__] = cls
    return cls
