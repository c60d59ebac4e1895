"""Host helpers with the reference's names (/root/reference/torchmx/utils.py): logger factory,
seed helper, uint4 size arithmetic, and the fp4 pack / unpack layout ops.

Nibble order (pinned by the reference, utils.py:145 and tests/test_mx_tensor.py:526-533): byte j of
the flattened tensor holds element 2j in its HIGH nibble and element 2j+1 in its LOW nibble.  The
CUDA kernels fuse this packing into their stores/loads; the two functions below are plain tensor
layout ops for callers that hold unpacked nibbles (they run on whatever device the tensor is on and
do no MX arithmetic).
"""
from __future__ import annotations

import logging
import math
import random
from typing import Iterable, List

import numpy as np
import torch

from . import env_variables as _env

_CONFIGURED = set()


def get_logger(logger_name: str = "TORCHMX",
               format_string: str = "%(asctime)s - %(name)s - %(levelname)s - %(message)s",
               console_output: bool = True) -> logging.Logger:
    """reference: utils.py:12-41 (level from env LOG_LEVEL, optional LOG_FILE, no propagation)."""
    log = logging.getLogger(logger_name)
    log.setLevel(_env.TORCHMX_LOG_LEVEL)
    if logger_name not in _CONFIGURED:  # the reference re-adds handlers on every call; once is enough
        fmt = logging.Formatter(format_string)
        if console_output:
            h = logging.StreamHandler()
            h.setFormatter(fmt)
            log.addHandler(h)
        if _env.TORCHMX_LOG_FILE:
            fh = logging.FileHandler(_env.TORCHMX_LOG_FILE)
            fh.setFormatter(fmt)
            log.addHandler(fh)
        _CONFIGURED.add(logger_name)
    log.propagate = False
    return log


def get_uniform_random_number(min_val: int, max_val: int, shape: Iterable[int], dtype: torch.dtype) -> torch.Tensor:
    """U[min_val, max_val) (reference: utils.py:44-58)."""
    return torch.rand(*shape, dtype=dtype) * (max_val - min_val) + min_val


def set_seed(seed: int) -> None:
    """reference: utils.py:148-159."""
    random.seed(seed)
    np.random.seed(seed)
    torch.random.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def tensor_size_hp_to_fp4x2(orig_size, packing_dim: int) -> List[int]:
    """logical size -> packed-byte size along `packing_dim` (ceil: an odd tail still needs a byte)."""
    out = list(orig_size)
    out[packing_dim] = math.ceil(out[packing_dim] / 2)
    return out


def tensor_size_fp4x2_to_hp(orig_size, unpacking_dim: int) -> List[int]:
    out = list(orig_size)
    out[unpacking_dim] = out[unpacking_dim] * 2
    return out


def unpack_uint4(uint8_data: torch.Tensor, packing_dim: int = -1) -> torch.Tensor:
    """[.., n, ..] bytes -> [.., 2n, ..] nibbles along `packing_dim` (high nibble first)."""
    d = packing_dim % uint8_data.dim()
    hi = uint8_data >> 4
    lo = uint8_data & 0xF
    return torch.stack((hi, lo), dim=d + 1).reshape(tensor_size_fp4x2_to_hp(uint8_data.shape, d)).to(torch.uint8)


def pack_uint4(uint8_data: torch.Tensor, packing_dim: int = -1) -> torch.Tensor:
    """nibbles -> bytes over the FLATTENED contiguous tensor (pairs may straddle rows exactly as in
    the reference, utils.py:144-145); the result takes the size halved along `packing_dim`."""
    shape = uint8_data.shape
    assert shape[packing_dim] % 2 == 0
    flat = uint8_data.contiguous().reshape(-1)
    return ((flat[0::2] << 4) | flat[1::2]).reshape(tensor_size_hp_to_fp4x2(shape, packing_dim))
