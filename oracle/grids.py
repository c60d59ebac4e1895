"""Deterministic input grids shared by oracle/gen_golden.py and tests/ (test infrastructure).

Built with plain numpy integer arithmetic (no RNG), so the same bits come out on every box.
"""
from __future__ import annotations

import numpy as np

BLOCK = 32

# Lead elements: one per exponent field value (mantissa 0), plus the specials.
# 0x7F80 = +Inf is already (255 << 7); add -Inf, +NaN, -NaN.
LEADS = np.concatenate([
    (np.arange(256, dtype=np.uint32) << 7).astype(np.uint16),
    np.array([0xFF80, 0x7FC0, 0xFFC1], dtype=np.uint16),
])


def quant_grid(leads: np.ndarray = LEADS) -> np.ndarray:
    """uint16 [n_leads * 2115, 32]: every bf16 bit pattern appears as a payload element in a
    block together with every lead; the lead sits at position (block index % 32) so that all
    lanes of the amax reduction are exercised.  The 65,536 patterns are cut into groups of 31
    (the last group is padded with +0)."""
    pats = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    n_groups = -(-65536 // 31)
    pad = np.zeros(n_groups * 31 - 65536, dtype=np.uint16)
    groups = np.concatenate([pats, pad]).reshape(n_groups, 31)
    out = np.empty((len(leads), n_groups, BLOCK), dtype=np.uint16)
    pos = np.arange(n_groups) % BLOCK
    # column index of payload j in a block whose lead is at position p: j if j < p else j + 1
    pay_cols = np.where(np.arange(31)[None, :] < pos[:, None], np.arange(31)[None, :], np.arange(31)[None, :] + 1)
    rows = np.arange(n_groups)[:, None]
    for i, lead in enumerate(leads):
        blk = np.empty((n_groups, BLOCK), dtype=np.uint16)
        blk[rows, pay_cols] = groups
        blk[np.arange(n_groups), pos] = lead
        out[i] = blk
    return out.reshape(-1, BLOCK)


def quant_grid_small() -> np.ndarray:
    """A 1/8 subsample of the leads (every 8th exponent + the extremes + the specials) for the
    CPU-only test tier, which must stay within minutes."""
    idx = sorted(set(list(range(0, 256, 8)) + [1, 2, 3, 4, 5, 6, 7, 9, 126, 127, 128, 247, 249, 250, 251, 252, 253, 254, 255, 256, 257, 258]))
    return quant_grid(LEADS[idx])


def dequant_grid(elem: str):
    """(codes uint8 [256, n], scales uint8 [256, n/32]) covering every (code byte, scale) pair.
    Row r uses scale r for all of its blocks.  6-bit formats only use bytes 0..63 (bits 7:6 of a
    stored fp6 code are always zero: mx_quantization_utils.py:400-402)."""
    if elem in ("float6_e3m2", "float6_e2m3"):
        row = np.tile(np.arange(64, dtype=np.uint8), 4)
    else:
        row = np.arange(256, dtype=np.uint8)
    codes = np.tile(row[None, :], (256, 1))
    per = 2 if elem == "float4_e2m1" else 1
    n_blocks = codes.shape[1] * per // BLOCK
    scales = np.tile(np.arange(256, dtype=np.uint8)[:, None], (1, n_blocks))
    if elem == "int8":
        codes = codes.view(np.int8)
    return codes, scales


def structured_cases():
    """Small hand-shaped inputs in the spirit of /root/reference/tests/test_mx_quantization.py
    (normal->normal, ->saturation, ->subnormal, underflow, zeros, bf16 subnormal inputs) at the
    block sizes the reference tests use (3..6) plus 2, 8, 16.  Returns {name: (uint16 array, bs)}."""
    cases = {}
    man = np.array([0b1111111, 0b0001010, 0b1000001, 0b1, 0b0101010, 0, 0b0110011, 0b1110010], dtype=np.uint16)
    sgn = np.array([1, 0, 0, 1, 0, 1, 0, 1], dtype=np.uint16)
    for bs in (2, 3, 4, 5, 6, 8, 16):
        rows = []
        for base in (0, 1, 5, 12, 100, 118, 127, 200, 240, 250, 254):
            for bump in (0, 1, 2, 3, 4, 6, 8, 9, 10, 14, 19):
                e = np.full(bs, base, dtype=np.uint16)
                e[-1] = min(base + bump, 254)
                rows.append(np.resize(sgn, bs) << 15 | e << 7 | np.resize(man, bs))
        cases[f"bs{bs}"] = (np.stack(rows).astype(np.uint16), bs)
    return cases
