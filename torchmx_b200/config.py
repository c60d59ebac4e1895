"""Quantization configs, API-compatible with /root/reference/torchmx/config.py (MXConfig :24-95,
QLinearConfig :99-148, QAttentionConfig :152-262): frozen dataclasses that round-trip through
plain dicts.  Pure host-side data; nothing here touches the GPU path.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Any, Optional

from . import dtypes


class _DictConfig:
    """dict <-> dataclass plumbing shared by the three configs.  A field whose annotation names
    another config class is (de)serialised recursively; `None` sub-configs are omitted."""

    _NESTED: dict = {}

    @classmethod
    def load_from_dict(cls, config_dict: dict) -> Any:
        kwargs = {}
        for f in fields(cls):
            if f.name not in config_dict or config_dict[f.name] is None:
                continue
            sub = cls._NESTED.get(f.name)
            kwargs[f.name] = sub.load_from_dict(config_dict[f.name]) if sub else config_dict[f.name]
        return cls(**kwargs)

    def to_dict(self) -> dict:
        out = {}
        for f in fields(self):
            v = getattr(self, f.name)
            if v is None:
                continue
            out[f.name] = v.to_dict() if isinstance(v, _DictConfig) else v
        return out


@dataclass(frozen=True)
class MXConfig(_DictConfig):
    """One MX tensor format: element dtype (by name, see dtypes.STR_TO_SUPPORTED_ELEM_DTYPE) and
    block size (default 32).  Raises ValueError on unknown names / block_size < 1
    (reference: config.py:43-50)."""

    elem_dtype_name: str
    block_size: int = 32

    def __post_init__(self):
        if self.elem_dtype_name not in dtypes.STR_TO_SUPPORTED_ELEM_DTYPE:
            raise ValueError(
                f"Unsupported element dtype name: {self.elem_dtype_name}. "
                f"Supported names are: {tuple(dtypes.STR_TO_SUPPORTED_ELEM_DTYPE.keys())}")
        if self.block_size < 1:
            raise ValueError(f"Block size must be at least 1, got {self.block_size}")

    @property
    def elem_dtype(self) -> dtypes.DType:
        return dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[self.elem_dtype_name]


@dataclass(frozen=True)
class QLinearConfig(_DictConfig):
    """Weights + activations formats of one linear layer (reference: config.py:99-148)."""

    weights_config: MXConfig
    activations_config: MXConfig

    _NESTED = {"weights_config": MXConfig, "activations_config": MXConfig}


@dataclass(frozen=True)
class QAttentionConfig(_DictConfig):
    """Projection linears + optional q / k / v / attention-weights formats
    (reference: config.py:152-262).  The reference intends "all four or none" for the optional
    entries but its check (config.py:186-198) can never fire; like the reference we do not raise,
    `is_qkv_quantization_enabled` simply reports whether all four are present."""

    projection_config: QLinearConfig
    query_config: Optional[MXConfig] = None
    key_config: Optional[MXConfig] = None
    value_config: Optional[MXConfig] = None
    attention_weights_config: Optional[MXConfig] = None

    _NESTED = {"projection_config": QLinearConfig, "query_config": MXConfig, "key_config": MXConfig,
               "value_config": MXConfig, "attention_weights_config": MXConfig}

    @property
    def is_qkv_quantization_enabled(self) -> bool:
        return all((self.query_config, self.key_config, self.value_config, self.attention_weights_config))
