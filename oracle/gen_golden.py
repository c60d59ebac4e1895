"""Pin the oracle against the real reference and write the golden fixtures.

Run HERE (the container that has /root/reference), never on the GPU box:

    python oracle/gen_golden.py            # writes tests/golden/digests.json + fixtures.npz

What it does
  1. imports the unmodified reference package from /root/reference through the torchao shim
     (oracle/_shim; the reference depends on torchao==0.6.1 which is not installed here);
  2. runs the reference's torchmx::quantize_mx / dequantize_mx custom ops on the exhaustive
     grids of oracle/grids.py for every element type and both MX_HARDWARE_EXACT_QUANTIZATION
     values, and REQUIRES the C oracle (oracle/mx_oracle.c) to agree bit-for-bit;
  3. stores SHA-256 digests of the reference outputs (tests recompute them from the oracle and,
     on the GPU, from the CUDA kernels) and small raw fixtures produced by the reference's
     user-level API (MXTensor.to_mx / to_dtype / matmul / F.linear, padding, transposes).

Test infrastructure only.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
# the reference package is called `torchmx`, and so is this repo's drop-in alias package: make
# sure the reference wins in THIS process, and that the repo root is only used for `oracle.*`.
sys.path = [os.path.join(HERE, "_shim"), REF] + [p for p in sys.path if os.path.abspath(p or ".") != ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import torchmx  # noqa: E402  (the reference)
assert os.path.abspath(torchmx.__file__).startswith(REF), torchmx.__file__
from torchmx import dtypes as rdt  # noqa: E402
from torchmx import env_variables as renv  # noqa: E402
from torchmx.mx_tensor import MXTensor, dequantize_mx, quantize_mx  # noqa: E402

import importlib.util  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


grids = _load("oracle_grids", os.path.join(HERE, "grids.py"))
mxo = _load("oracle_mx", os.path.join(HERE, "mx_oracle.py"))

ELEMS = ["float8_e4m3", "float6_e3m2", "float6_e2m3", "float4_e2m1", "int8"]
MODES = {"simulated": "False", "hw_exact": "True"}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def t_bf16(bits: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16)


def bits_of(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()
    if t.dtype == torch.float32:
        return t.contiguous().numpy().view(np.uint32).copy()
    return t.contiguous().numpy().copy()


def ref_quantize(bits: np.ndarray, elem: str, bs: int, mode: str):
    renv.MX_EXACT_QUANTIZATION = MODES[mode]
    scale, data = quantize_mx(t_bf16(bits), elem, bs)
    return scale.numpy().copy(), data.numpy().copy()


def ref_dequantize(codes: np.ndarray, scales: np.ndarray, elem: str, bs: int, target: str, block_dim: int):
    td = torch.bfloat16 if target == "bf16" else torch.float32
    out = dequantize_mx(torch.from_numpy(codes), torch.from_numpy(scales), elem, bs, td, block_dim)
    return bits_of(out)


def main():
    t0 = time.time()
    torch.set_num_threads(os.cpu_count() or 1)
    digests = {"_meta": {
        "reference": "rain-neuromorphics/torchmx at /root/reference (imported, unmodified)",
        "torch": torch.__version__,
        "generator": "oracle/gen_golden.py",
        "note": "sha256 over the raw bytes of the reference's outputs on oracle/grids.py inputs",
    }}
    fx = {}

    # ---- 1. exhaustive quantize grid --------------------------------------------------
    grid = grids.quant_grid()
    digests["quant_grid_input"] = sha(grid)
    small = grids.quant_grid_small()
    digests["quant_grid_small_input"] = sha(small)
    for elem in ELEMS:
        for mode in MODES:
            if elem == "int8" and mode == "hw_exact":
                pass  # the reference ignores the toggle for int8 (mx_tensor.py:80-90); still recorded
            rs, rc = ref_quantize(grid, elem, 32, mode)
            os_, oc = mxo.quantize(grid, elem, 32, hw_exact=(mode == "hw_exact"), threads=8)
            assert np.array_equal(rs, os_), f"scale mismatch {elem} {mode}"
            bad = np.nonzero(rc.view(np.uint8) != oc.view(np.uint8))
            assert bad[0].size == 0, f"code mismatch {elem} {mode}: {bad[0].size} elements, first {bad[0][:5]}"
            digests[f"quant_grid/{elem}/{mode}/scales"] = sha(rs)
            digests[f"quant_grid/{elem}/{mode}/codes"] = sha(rc)
            ss, sc = ref_quantize(small, elem, 32, mode)
            digests[f"quant_grid_small/{elem}/{mode}/scales"] = sha(ss)
            digests[f"quant_grid_small/{elem}/{mode}/codes"] = sha(sc)
            print(f"[{time.time()-t0:6.1f}s] quant grid {elem:12s} {mode:9s} oracle == reference on {grid.size} elements", flush=True)
    # where hw_exact and simulated differ at all (expected: NaN-scale blocks only)
    for elem in ELEMS[:4]:
        s0, c0 = ref_quantize(grid, elem, 32, "simulated")
        s1, c1 = ref_quantize(grid, elem, 32, "hw_exact")
        if elem == "float4_e2m1":
            d = (c0 != c1).reshape(-1, 16).any(axis=1)
        else:
            d = (c0 != c1).reshape(-1, 32).any(axis=1)
        assert np.all(s0.reshape(-1)[d] == 255), "modes differ outside NaN-scale blocks"
        digests[f"mode_diff_blocks/{elem}"] = int(d.sum())

    # ---- 2. structured small cases at odd block sizes ---------------------------------
    for name, (bits, bs) in grids.structured_cases().items():
        fx[f"struct/{name}/x"] = bits
        for elem in ELEMS:
            if elem == "float4_e2m1" and (bits.size % 2):
                continue
            for mode in MODES:
                rs, rc = ref_quantize(bits, elem, bs, mode)
                os_, oc = mxo.quantize(bits, elem, bs, hw_exact=(mode == "hw_exact"))
                assert np.array_equal(rs, os_) and np.array_equal(rc.view(np.uint8), oc.view(np.uint8)), (name, elem, mode)
                fx[f"struct/{name}/{elem}/{mode}/scales"] = rs
                fx[f"struct/{name}/{elem}/{mode}/codes"] = rc

    # ---- 3. exhaustive dequantize grid -------------------------------------------------
    for elem in ELEMS:
        codes, scales = grids.dequant_grid(elem)
        for target in ("bf16", "f32"):
            r = ref_dequantize(codes, scales, elem, 32, target, 1)
            o = mxo.dequantize(codes, scales, elem, 32, target, 1)
            o = o.view(np.uint32) if target == "f32" else o
            # NaN payloads: compare NaN-ness, then canonicalise so the digest is payload-free
            if target == "bf16":
                rn, on = (r & 0x7FFF) > 0x7F80, (o & 0x7FFF) > 0x7F80
                r = np.where(rn, np.uint16(0x7FC0), r); o = np.where(on, np.uint16(0x7FC0), o)
            else:
                rn, on = (r & 0x7FFFFFFF) > 0x7F800000, (o & 0x7FFFFFFF) > 0x7F800000
                r = np.where(rn, np.uint32(0x7FC00000), r); o = np.where(on, np.uint32(0x7FC00000), o)
            bad = np.nonzero(r != o)
            assert bad[0].size == 0, f"dequant mismatch {elem} {target}: {bad[0].size}, first {[(int(a), int(b)) for a, b in zip(bad[0][:5], bad[1][:5])]}"
            digests[f"dequant_grid/{elem}/{target}"] = sha(r)
            print(f"[{time.time()-t0:6.1f}s] dequant grid {elem:12s} -> {target}: oracle == reference on {r.size} pairs (NaNs canonicalised)", flush=True)

    # ---- 4. user-level fixtures (MXTensor API) -----------------------------------------
    torch.manual_seed(0)
    x = torch.randn(128, 128, dtype=torch.bfloat16)
    fx["readme/x"] = bits_of(x)
    for elem in ELEMS:
        for mode in MODES:
            renv.MX_EXACT_QUANTIZATION = MODES[mode]
            m = MXTensor.to_mx(x, rdt.STR_TO_SUPPORTED_ELEM_DTYPE[elem], 32)
            fx[f"readme/{elem}/{mode}/scales"] = m._scale_e8m0.numpy().copy()
            fx[f"readme/{elem}/{mode}/codes"] = m._data.numpy().copy()
        fx[f"readme/{elem}/bf16"] = bits_of(m.to_dtype(torch.bfloat16))
        fx[f"readme/{elem}/f32"] = bits_of(m.to_dtype(torch.float32))
    renv.MX_EXACT_QUANTIZATION = "False"

    # special values (tests/conftest.py:51-63 shape) -- all codes zero, scales 255, dequant NaN
    sp = torch.randn(5, 4, dtype=torch.bfloat16)
    sp[0, 1] = float("inf"); sp[1, 1] = float("-inf"); sp[2, 1] = float("nan"); sp[3, 1] = -float("nan")
    sp[4, 1], sp[4, 2] = float("nan"), float("inf")
    fx["special/x"] = bits_of(sp)
    for elem in ELEMS:
        for mode in MODES:
            renv.MX_EXACT_QUANTIZATION = MODES[mode]
            m = MXTensor.to_mx(sp, rdt.STR_TO_SUPPORTED_ELEM_DTYPE[elem], 4)
            fx[f"special/{elem}/{mode}/scales"] = m._scale_e8m0.numpy().copy()
            fx[f"special/{elem}/{mode}/codes"] = m._data.numpy().copy()
    renv.MX_EXACT_QUANTIZATION = "False"

    # padding: last dim not a multiple of the block (mx_tensor.py:218-248, 288-321)
    g = torch.Generator().manual_seed(1)
    for shape, bs in (((3, 37), 32), ((2, 5, 45), 8), ((7, 33), 4), ((4, 70), 32), ((6, 31), 2)):
        xp = (torch.randn(*shape, generator=g) * 3).to(torch.bfloat16)
        key = "pad/" + "x".join(map(str, shape)) + f"_bs{bs}"
        fx[key + "/x"] = bits_of(xp)
        for elem in ELEMS:
            m = MXTensor.to_mx(xp, rdt.STR_TO_SUPPORTED_ELEM_DTYPE[elem], bs)
            fx[f"{key}/{elem}/scales"] = m._scale_e8m0.numpy().copy()
            fx[f"{key}/{elem}/codes"] = m._data.numpy().copy()
            fx[f"{key}/{elem}/padding"] = np.array([m._padding])
            fx[f"{key}/{elem}/shape"] = np.array(list(m.shape))
            fx[f"{key}/{elem}/bf16"] = bits_of(m.to_dtype(torch.bfloat16))

    # layout ops: t / transpose / 4-D view then to_dtype (tests/test_mx_tensor.py:195-356)
    xt = (torch.randn(4, 6, 64, 96, generator=g) * 2).to(torch.bfloat16)
    fx["layout/x"] = bits_of(xt)
    for elem in ELEMS:
        et = rdt.STR_TO_SUPPORTED_ELEM_DTYPE[elem]
        m = MXTensor.to_mx(xt, et, 32)
        fx[f"layout/{elem}/transpose23_bf16"] = bits_of(m.transpose(2, 3).to_dtype(torch.bfloat16))
        fx[f"layout/{elem}/transpose23_f32"] = bits_of(m.transpose(2, 3).to_dtype(torch.float32))
        m2 = MXTensor.to_mx(xt[0, 0], et, 32)
        fx[f"layout/{elem}/t_bf16"] = bits_of(m2.t().to_dtype(torch.bfloat16))
        fx[f"layout/{elem}/view_bf16"] = bits_of(m.view(24, 64, 96).to_dtype(torch.bfloat16))

    # matmul family: the reference = dequantize -> aten op in bf16 (ops.py:29-41,60-68,99-119)
    a = (torch.randn(64, 96, generator=g)).to(torch.bfloat16)
    w = (torch.randn(48, 96, generator=g)).to(torch.bfloat16)
    bias = torch.randn(48, generator=g).to(torch.bfloat16)
    q = torch.randn(2, 3, 32, 64, generator=g).to(torch.bfloat16)
    k = torch.randn(2, 3, 40, 64, generator=g).to(torch.bfloat16)
    for name, t in (("a", a), ("w", w), ("bias", bias), ("q", q), ("k", k)):
        fx[f"mm/{name}"] = bits_of(t)
    for ea, ew in (("float8_e4m3", "float6_e3m2"), ("float8_e4m3", "float4_e2m1"), ("float6_e2m3", "float6_e3m2"),
                   ("float4_e2m1", "float4_e2m1"), ("int8", "int8"), ("float8_e4m3", "float8_e4m3")):
        A = MXTensor.to_mx(a, rdt.STR_TO_SUPPORTED_ELEM_DTYPE[ea], 32)
        W = MXTensor.to_mx(w, rdt.STR_TO_SUPPORTED_ELEM_DTYPE[ew], 32)
        Q = MXTensor.to_mx(q, rdt.STR_TO_SUPPORTED_ELEM_DTYPE[ea], 32)
        K = MXTensor.to_mx(k, rdt.STR_TO_SUPPORTED_ELEM_DTYPE[ew], 32)
        tag = f"mm/{ea}x{ew}"
        with torch.no_grad():
            fx[tag + "/linear"] = bits_of(torch.nn.functional.linear(A, W))
            fx[tag + "/linear_bias"] = bits_of(torch.nn.functional.linear(A, W, bias))
            fx[tag + "/mm"] = bits_of(torch.matmul(A, W.t()))
            fx[tag + "/qk"] = bits_of(torch.matmul(Q, K.transpose(2, 3)))

    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fixtures.npz"), **fx)
    sz = os.path.getsize(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    print(f"[{time.time()-t0:6.1f}s] wrote {len(digests)} digests and {len(fx)} fixture arrays ({sz/1e6:.2f} MB)")


if __name__ == "__main__":
    main()
