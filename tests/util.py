"""helpers shared by the tests: bf16 bit patterns <-> torch tensors, digests, NaN canonicalisation."""
import hashlib

import numpy as np
import torch

ELEMS = ["float8_e4m3", "float6_e3m2", "float6_e2m3", "float4_e2m1", "int8"]
FP_ELEMS = ELEMS[:4]
MODES = {"simulated": "False", "hw_exact": "True"}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def bf16_tensor(bits: np.ndarray, device="cpu") -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(bits).view(np.int16).copy()).view(torch.bfloat16).to(device)


def bits_of(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16).copy()
    if t.dtype == torch.float32:
        return t.numpy().view(np.uint32).copy()
    return t.numpy().copy()


def canon_nan(bits: np.ndarray) -> np.ndarray:
    """NaN payload/sign is not part of the contract: map every NaN to the canonical quiet NaN."""
    if bits.dtype == np.uint16:
        return np.where((bits & 0x7FFF) > 0x7F80, np.uint16(0x7FC0), bits)
    if bits.dtype == np.uint32:
        return np.where((bits & 0x7FFFFFFF) > 0x7F800000, np.uint32(0x7FC00000), bits)
    return bits


def assert_bits_equal(got: np.ndarray, want: np.ndarray, what=""):
    got, want = canon_nan(np.asarray(got)), canon_nan(np.asarray(want))
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    g, w = got.view(np.uint8) if got.dtype == np.int8 else got, want.view(np.uint8) if want.dtype == np.int8 else want
    bad = np.nonzero(g != w)
    if bad[0].size:
        first = tuple(int(b[0]) for b in bad)
        raise AssertionError(f"{what}: {bad[0].size} mismatching elements of {g.size}; first at {first}: got {g[first]} want {w[first]}")
