"""Model surgery: replace conv/linear (+bn) (+relu) child sequences by fused QAT layers.

Reference: modules/fuse.py:9-277 (``fuse_modules_unified(model, fuse_patterns, is_trace=False, config_manager=None)``).
Behaviour kept: patterns are lists of 'conv' | 'linear' | 'bn' | 'relu' matched against CONSECUTIVE children of every
module (registration order); 'relu' matches nn.ReLU and nn.SiLU (fuse.py:99); already-fused modules are not descended
into (:77-81); the first child of a match is replaced by the fused layer, the others by nn.Identity (:145-148).
Differences, both supersets:
  * the per-layer config is looked up by the bare child name first (what the reference does, fuse.py:113-114) and, when
    no pattern matches that, by the fully qualified name -- so rules such as "backbone.*conv" / "head" work;
  * ``is_trace=True`` takes the same direct path: the reference's fx path builds fused layers inside a throw-away
    GraphModule and returns the original, un-fused model (fuse.py:241-252)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch.nn as nn

from .fuse_config import FuseConfig, FuseConfigManager
from .fused import (FUSED_CLASSES, Conv, ConvBn, ConvBnReLU, ConvReLU, Linear, LinearBn, LinearBnReLU, LinearReLU)

PATTERN_TO_FUSED = {
    ("conv", "bn", "relu"): ConvBnReLU,
    ("conv", "bn"): ConvBn,
    ("conv", "relu"): ConvReLU,
    ("linear", "bn", "relu"): LinearBnReLU,
    ("linear", "bn"): LinearBn,
    ("linear", "relu"): LinearReLU,
    ("conv",): Conv,
    ("linear",): Linear,
}

_KIND = {
    "conv": (nn.Conv2d,),
    "linear": (nn.Linear,),
    "bn": (nn.BatchNorm2d, nn.BatchNorm1d),
    "relu": (nn.ReLU, nn.SiLU),
}
_WITH_FUSE_BN_FLAG = (ConvBnReLU, ConvBn, LinearBnReLU, LinearBn)


def get_module_type_str(module) -> Optional[str]:
    """'conv' | 'linear' | 'bn' | 'relu' | None (fuse.py:29-43; only nn.ReLU counts as 'relu' there)."""
    for kind, types in (("conv", _KIND["conv"]), ("linear", _KIND["linear"]), ("bn", _KIND["bn"]), ("relu", (nn.ReLU,))):
        if isinstance(module, types):
            return kind
    return None


def find_fusable_sequences(model: nn.Module, pattern: Sequence[str]) -> List[Tuple[str, nn.Module, List[str]]]:
    """(qualified parent name, parent module, child names) for every non-overlapping run of consecutive children that
    matches ``pattern``.  Pure matching, no tensors touched (unit-testable without a GPU)."""
    kinds = [_KIND[k] for k in pattern]
    hits = []
    for parent_name, parent in model.named_modules():
        if isinstance(parent, FUSED_CLASSES) or len(parent._modules) < len(pattern):
            continue
        names = list(parent._modules.keys())
        i = 0
        while i + len(pattern) <= len(names):
            window = names[i:i + len(pattern)]
            if all(isinstance(parent._modules[n], k) for n, k in zip(window, kinds)):
                hits.append((parent_name, parent, window))
                i += len(pattern)
            else:
                i += 1
    return hits


def _config_for(config_manager: FuseConfigManager, child: str, qualified: str) -> FuseConfig:
    cfg = config_manager.find_config(child) if hasattr(config_manager, "find_config") else None
    if cfg is None and hasattr(config_manager, "find_config"):
        cfg = config_manager.find_config(qualified)
    if cfg is None:
        cfg = config_manager.get_config_for_layer(child)
    return cfg


def _build(fused_class, layers: Iterable[nn.Module], cfg: FuseConfig):
    args = list(layers) + [cfg.observer_w_name, cfg.quantizer_w_name, cfg.observer_a_name, cfg.quantizer_a_name,
                           cfg.w_symmetric, cfg.a_symmetric]
    if fused_class in _WITH_FUSE_BN_FLAG:
        args.append(cfg.is_fuse_bn)
    args += [cfg.bits_w, cfg.bits_a]
    ext = {}
    if getattr(cfg, "w_ch_axis", None) is not None:
        ext["w_ch_axis"] = cfg.w_ch_axis
    if getattr(cfg, "a_ch_axis", None) is not None:
        ext["a_ch_axis"] = cfg.a_ch_axis
    return fused_class(*args, **ext)


def _fuse_modules(model, fuse_patterns, config_manager=None):
    if config_manager is None:
        config_manager = FuseConfigManager()
    for pattern in fuse_patterns:
        fused_class = PATTERN_TO_FUSED.get(tuple(pattern))
        if fused_class is None:
            continue  # unknown patterns are skipped silently, like fuse.py:71-73
        for parent_name, parent, names in find_fusable_sequences(model, pattern):
            qualified = f"{parent_name}.{names[0]}" if parent_name else names[0]
            cfg = _config_for(config_manager, names[0], qualified)
            parent._modules[names[0]] = _build(fused_class, [parent._modules[n] for n in names], cfg)
            for n in names[1:]:
                parent._modules[n] = nn.Identity()
    return model


def _fuse_modules_trace(model, fuse_patterns, config_manager=None):
    return _fuse_modules(model, fuse_patterns, config_manager)


def fuse_modules_unified(model, fuse_patterns, is_trace=False, config_manager=None):
    """Fuse ``model`` in place and return it (fuse.py:254-277)."""
    if is_trace:
        return _fuse_modules_trace(model, fuse_patterns, config_manager)
    return _fuse_modules(model, fuse_patterns, config_manager)
