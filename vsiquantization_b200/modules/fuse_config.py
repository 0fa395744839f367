"""Per-layer fusion / quantisation configuration (reference: modules/fuse_config.py:6-243).

FuseConfig keeps the reference's nine fields, order and defaults; ``FuseConfig(**dict)`` raises TypeError on unknown
keys exactly like the reference.  Two extension fields default to the reference's behaviour: ``w_ch_axis`` /
``a_ch_axis`` (None = per tensor) select per-channel quantisation (quantizers/lsq_module.py's capability, which the
reference never wires to this config path)."""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import yaml


@dataclass
class FuseConfig:
    observer_w_name: str = "MinMaxObserver"
    quantizer_w_name: str = "UniformQuantizer"
    observer_a_name: str = "MinMaxObserver"
    quantizer_a_name: str = "UniformQuantizer"
    w_symmetric: bool = True
    a_symmetric: bool = True
    is_fuse_bn: bool = True
    bits_w: int = 8
    bits_a: int = 8
    w_ch_axis: Optional[int] = None  # extension
    a_ch_axis: Optional[int] = None  # extension


class FuseConfigManager:
    """Maps layer names to FuseConfigs: first registered pattern that matches wins, insertion order
    (fuse_config.py:96-99); a pattern is a regex (re.search) with substring fallback when it does not compile
    (:101-123)."""

    def __init__(self, default_config: Optional[FuseConfig] = None):
        self.default_config = default_config or FuseConfig()
        self.layer_configs: Dict[str, FuseConfig] = {}

    def add_layer_config(self, layer_pattern: str, config: FuseConfig) -> None:
        self.layer_configs[layer_pattern] = config

    @staticmethod
    def _match_pattern(layer_name: str, pattern: str) -> bool:
        try:
            return re.search(pattern, layer_name) is not None
        except re.error:
            return pattern in layer_name

    def find_config(self, layer_name: str) -> Optional[FuseConfig]:
        for pattern, config in self.layer_configs.items():
            if self._match_pattern(layer_name, pattern):
                return config
        return None

    def get_config_for_layer(self, layer_name: str) -> FuseConfig:
        found = self.find_config(layer_name)
        return found if found is not None else self.default_config

    def set_default_config(self, config: FuseConfig) -> None:
        self.default_config = config

    def clear_layer_configs(self) -> None:
        self.layer_configs.clear()

    def get_all_patterns(self) -> List[str]:
        return list(self.layer_configs)

    def __repr__(self):
        return f"FuseConfigManager(default={self.default_config}, patterns={list(self.layer_configs)})"


def load_fuse_config_from_yaml(yaml_path: str) -> FuseConfigManager:
    """YAML schema of the reference (fuse_config.py:152-209): optional ``default:`` mapping and ``layers:`` mapping of
    pattern -> FuseConfig fields.  FileNotFoundError for a missing file, ValueError for malformed YAML."""
    try:
        with open(yaml_path, "r", encoding="utf-8") as f:
            data = yaml.safe_load(f)
    except FileNotFoundError:
        raise FileNotFoundError(f"Configuration file not found: {yaml_path}")
    except yaml.YAMLError as e:
        raise ValueError(f"Invalid YAML format in {yaml_path}: {e}")
    manager = FuseConfigManager()
    data = data or {}
    if "default" in data:
        manager.default_config = FuseConfig(**data["default"])
    for pattern, fields in (data.get("layers") or {}).items():
        manager.add_layer_config(pattern, FuseConfig(**fields))
    return manager


def create_fuse_config_manager(default_config: Optional[FuseConfig] = None,
                               layer_configs: Optional[Dict[str, Union[FuseConfig, Dict]]] = None) -> FuseConfigManager:
    """Programmatic counterpart of the YAML loader (fuse_config.py:212-244)."""
    manager = FuseConfigManager(default_config)
    for pattern, config in (layer_configs or {}).items():
        if isinstance(config, dict):
            config = FuseConfig(**config)
        elif not isinstance(config, FuseConfig):
            raise ValueError(f"Config for pattern '{pattern}' must be FuseConfig or dict")
        manager.add_layer_config(pattern, config)
    return manager
