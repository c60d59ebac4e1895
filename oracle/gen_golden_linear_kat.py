"""Golden vectors for the reference's exact linear-layer KAT (tests/layers/test_mx_linear.py:64-114 with the fixtures
tests/layers/conftest.py:10-20, 56-80): run the UNMODIFIED reference on CPU (through oracle/_shim) on the fixed 2x4 input and
6x4 weight, block size 2, for the eight GEMM_COMBINATIONS plus int8 x int8, and store the bf16 output bits and the SQNR the
reference's own test computes.  Run HERE (needs /root/reference), never on the GPU box:

    python oracle/gen_golden_linear_kat.py      # writes tests/golden/linear_kat.json

Test infrastructure only.
"""
import copy
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path = [os.path.join(HERE, "_shim"), REF] + [p for p in sys.path if os.path.abspath(p or ".") != ROOT]

import torch  # noqa: E402

import torchmx  # noqa: E402  (the reference)
assert os.path.abspath(torchmx.__file__).startswith(REF), torchmx.__file__
from torchao.quantization.utils import compute_error  # noqa: E402
from torchmx import env_variables as renv  # noqa: E402
from torchmx.config import MXConfig, QLinearConfig  # noqa: E402
from torchmx.quant_api import quantize_linear_  # noqa: E402

COMBOS = {"0": ("float8_e4m3", "float6_e3m2"), "1": ("float8_e4m3", "float4_e2m1"), "2": ("float6_e3m2", "float6_e3m2"),
          "3": ("float6_e3m2", "float4_e2m1"), "4": ("float6_e2m3", "float6_e3m2"), "5": ("float6_e2m3", "float4_e2m1"),
          "6": ("float4_e2m1", "float6_e3m2"), "7": ("float4_e2m1", "float4_e2m1"), "int8": ("int8", "int8")}
TABLE = {"0": 41.5, "1": 19.25, "2": 41.5, "3": 19.25, "4": 41.5, "5": 19.25, "6": 41.5, "7": 19.25, "int8": 47.5}  # conftest.py:10-20


def bits(t):
    return t.contiguous().view(torch.int16).flatten().tolist()


def main():
    w = torch.arange(4 * 6, dtype=torch.bfloat16).view(6, 4) + 0.123              # conftest.py:68-70
    x0 = torch.pow(2.0, -torch.arange(2 * 4).view(2, 4)).bfloat16()              # conftest.py:73-75
    w_pad = w.t().contiguous()                                                    # test_mx_linear.py:256-257 (Linear(6, 4))
    x_pad = torch.pow(2.0, -torch.arange(2 * 6).view(2, 6)).bfloat16()           # conftest.py:78-80
    out = {"weights": bits(w), "input": bits(x0), "input_padded": bits(x_pad), "cases": {}}
    for mode_name, mode in (("simulated", "False"), ("hw_exact", "True")):
        renv.MX_EXACT_QUANTIZATION = mode
        for key, (act, wt) in COMBOS.items():
            for block, (xin, wmat, tag) in ((2, (x0, w, "plain")), (4, (x_pad, w_pad, "padded"))):
                m = torch.nn.Sequential(torch.nn.Linear(wmat.shape[1], wmat.shape[0], bias=False, dtype=torch.bfloat16))
                m[0].weight.data = wmat.clone()
                m_mx = copy.deepcopy(m)
                quantize_linear_(m_mx, QLinearConfig(weights_config=MXConfig(wt, block), activations_config=MXConfig(act, block)))
                with torch.inference_mode():
                    y_ref, y_mx = m(xin), m_mx(xin)
                sqnr = compute_error(y_ref, y_mx).item()
                if tag == "plain":
                    assert abs(sqnr - TABLE[key]) <= 1e-5 * TABLE[key], (key, sqnr)
                out["cases"][f"{mode_name}/{key}/{tag}"] = {"act": act, "weight": wt, "block": block, "y_ref": bits(y_ref), "y_mx": bits(y_mx), "sqnr": sqnr}
    path = os.path.join(ROOT, "tests", "golden", "linear_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
