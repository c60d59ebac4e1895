"""torchmx_b200 -- B200-native (sm_100a) implementation of torchmx's MX quantize / dequantize /
MX-matmul hot path behind torchmx's own operator API.

Importing the package registers the `torchmx::quantize_mx` / `torchmx::dequantize_mx` custom ops
and the MXTensor aten overrides (the reference's `torchmx/__init__.py` does the same by importing
`ops`).  `MXTensor` is re-exported so the README example `from torchmx import MXTensor` works.
"""
from . import dtypes, env_variables, config, utils  # noqa: F401
from . import mx_tensor  # noqa: F401  (registers the custom ops)
from . import ops  # noqa: F401  (registers the aten overrides)
from .mx_tensor import MXTensor  # noqa: F401

__all__ = ["MXTensor", "dtypes", "config", "env_variables", "utils", "mx_tensor", "ops"]
