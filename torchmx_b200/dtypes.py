"""Element-format descriptors for the OCP MX formats, API-compatible with the reference's
`torchmx.dtypes` (/root/reference/torchmx/dtypes.py:9-183): same `DType` fields, same module-level
names, same lookup tables.  The numeric constants are the format definitions themselves (OCP MX
v1.0 tables); the CUDA side mirrors them in csrc/mxq_common.cuh `Fmt<>`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass(frozen=True, repr=False)
class DType:
    name: str
    max: float                # largest finite magnitude
    max_pow2: int             # exponent of the largest binade
    exponent_bias: int
    exponent_bits: int
    mantissa_bits: int
    has_nan: bool
    has_inf: bool
    torch_dtype: Optional[torch.dtype] = None

    def __repr__(self) -> str:
        return self.name


def _fp(name, ebits, mbits, bias, *, nan, inf, torch_dtype=None, top_exp_is_special=None, max_override=None):
    """Derive max / max_pow2 from the bit layout instead of listing them."""
    if top_exp_is_special is None:
        top_exp_is_special = inf  # IEEE-like formats reserve the top exponent
    top = (1 << ebits) - 1 - (1 if top_exp_is_special else 0)
    max_pow2 = top - bias
    frac = 2.0 - 2.0 ** (-mbits)
    mx = max_override if max_override is not None else (2.0 ** max_pow2) * frac
    return DType(name=name, max=mx, max_pow2=max_pow2, exponent_bias=bias, exponent_bits=ebits,
                 mantissa_bits=mbits, has_nan=nan, has_inf=inf, torch_dtype=torch_dtype)


# e4m3 "fn": top exponent is usable, only mantissa 0b111 there is NaN -> max = 1.75 * 2^8
float8_e4m3 = _fp("float8_e4m3", 4, 3, 7, nan=True, inf=False, torch_dtype=torch.float8_e4m3fn, max_override=448.0)
float6_e3m2 = _fp("float6_e3m2", 3, 2, 3, nan=False, inf=False)
float6_e2m3 = _fp("float6_e2m3", 2, 3, 1, nan=False, inf=False)
float4_e2m1 = _fp("float4_e2m1", 2, 1, 1, nan=False, inf=False)
# MX int8 as the reference models it: the element is the integer itself, shared exponent = maxE - 6
int8 = DType(name="int8", max=127.0, max_pow2=6, exponent_bias=0, exponent_bits=0, mantissa_bits=7,
             has_nan=False, has_inf=False, torch_dtype=torch.int8)

float64 = _fp("float64", 11, 52, 1023, nan=True, inf=True, torch_dtype=torch.float64,
              max_override=torch.finfo(torch.float64).max)
float32 = _fp("float32", 8, 23, 127, nan=True, inf=True, torch_dtype=torch.float32,
              max_override=torch.finfo(torch.float32).max)
bfloat16 = _fp("bfloat16", 8, 7, 127, nan=True, inf=True, torch_dtype=torch.bfloat16,
               max_override=torch.finfo(torch.bfloat16).max)
float22_e8m13 = _fp("float22_e8m13", 8, 13, 127, nan=True, inf=True)

# B200 extension (not a torchmx element type; parity unpinned): IEEE-like fp8 with inf/nan
float8_e5m2 = _fp("float8_e5m2", 5, 2, 15, nan=True, inf=True, torch_dtype=torch.float8_e5m2)

SUPPORTED_FP_ELEM_DTYPES = (float8_e4m3, float6_e3m2, float6_e2m3, float4_e2m1)
SUPPORTED_ELEM_DTYPES = SUPPORTED_FP_ELEM_DTYPES + (int8,)
STR_TO_SUPPORTED_ELEM_DTYPE = {d.name: d for d in SUPPORTED_ELEM_DTYPES}

# extension table: everything the CUDA library can encode (reference set + e5m2)
EXTENDED_ELEM_DTYPES = SUPPORTED_ELEM_DTYPES + (float8_e5m2,)
STR_TO_ELEM_DTYPE = {d.name: d for d in EXTENDED_ELEM_DTYPES}

# E8M0 shared scale: 8 exponent bits, bias 127, 0xFF = NaN, no zero / inf (OCP MX 5.4.1)
e8m0 = DType(name="e8m0", max=2.0 ** 127, max_pow2=127, exponent_bias=127, exponent_bits=8, mantissa_bits=0,
             has_nan=True, has_inf=False)
E8M0_EXPONENT_NAN_VAL = 255

# ids used across the C ABI (include/mxq.h mxq_elem_t)
ELEM_ID = {"float8_e4m3": 0, "float6_e3m2": 1, "float6_e2m3": 2, "float4_e2m1": 3, "int8": 4, "float8_e5m2": 5}
