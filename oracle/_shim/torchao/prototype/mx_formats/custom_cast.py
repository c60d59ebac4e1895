"""Restatement of torchao 0.6.1 `prototype/mx_formats/custom_cast.py` casts
used by the reference's `simulated` branch (call sites:
torchmx/mx_quantization_utils.py:483,485,487).  Test infrastructure only.

Published algorithm (fp32 -> sub-byte float, unpacked one code per uint8):
  * split sign / magnitude;
  * magnitude >= max_normal            -> all-ones magnitude code (saturate);
  * magnitude <  min_normal            -> add a "magic" float whose ulp equals the
                                          target subnormal step, read the low bits;
  * otherwise                          -> rebias exponent, integer round-to-nearest-
                                          even on the dropped mantissa bits, shift;
  * re-attach the sign above the magnitude bits.
"""
import torch

_F32_MBITS = 23
_F32_BIAS = 127


def _f32_to_floatx_unpacked(x: torch.Tensor, ebits: int, mbits: int) -> torch.Tensor:
    assert x.dtype == torch.float32
    bias = (1 << (ebits - 1)) - 1
    max_int = (1 << (ebits + mbits)) - 1
    sign_mask = 1 << (ebits + mbits)
    drop = _F32_MBITS - mbits
    # largest normal: all-ones exponent AND all-ones mantissa (no inf/nan codes)
    max_normal = 2.0 ** ((1 << ebits) - 1 - bias) * (2.0 - 2.0 ** (-mbits))
    min_normal = 2.0 ** (1 - bias)

    bits = x.view(torch.int32)
    sign = bits & -0x80000000
    mag_bits = bits ^ sign
    mag = mag_bits.view(torch.float32)

    saturate = mag >= max_normal
    denorm = (~saturate) & (mag < min_normal)
    normal = ~(saturate | denorm)

    # subnormal targets: float add does the RNE for us
    magic_exp = (_F32_BIAS - bias) + (_F32_MBITS - mbits) + 1
    magic_i = magic_exp << _F32_MBITS
    magic_f = torch.tensor(magic_i, dtype=torch.int32).view(torch.float32)
    den_code = (mag + magic_f).view(torch.int32) - magic_i

    # normal targets: integer RNE on the dropped bits
    odd = (mag_bits >> drop) & 1
    nrm = mag_bits + ((bias - _F32_BIAS) << _F32_MBITS) + ((1 << (drop - 1)) - 1) + odd
    nrm_code = nrm >> drop

    out = torch.full_like(mag_bits, max_int)
    out = torch.where(denorm, den_code, out)
    out = torch.where(normal, nrm_code, out)
    out = out.to(torch.uint8)

    sign_lp = ((sign >> (_F32_MBITS + 8 - mbits - ebits)) & sign_mask).to(torch.uint8)
    return out | sign_lp


def f32_to_f4_unpacked(x):
    return _f32_to_floatx_unpacked(x, 2, 1)


def f32_to_f6_e2m3_unpacked(x):
    return _f32_to_floatx_unpacked(x, 2, 3)


def f32_to_f6_e3m2_unpacked(x):
    return _f32_to_floatx_unpacked(x, 3, 2)


def unpack_uint4(u: torch.Tensor) -> torch.Tensor:
    shape = list(u.shape)
    shape[-1] *= 2
    return torch.stack([u >> 4, u & 0xF], dim=-1).view(shape)


_F4_VALUES = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def f4_unpacked_to_f32(u: torch.Tensor) -> torch.Tensor:
    lut = torch.tensor(_F4_VALUES + [-v for v in _F4_VALUES], dtype=torch.float32)
    return lut[u.long()]
