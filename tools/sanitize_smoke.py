"""Small pass over every CUDA kernel of libmxq.so for compute-sanitizer (one tool per run): quantize / dequantize fast and
generic paths, strided dequantize, transcode / pack, the three GEMM kernels (bias, ragged edges, split-K, packed operands)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import dtypes, mx_gemm
from torchmx_b200.mx_tensor import MXTensor
torch.manual_seed(0)
dev = "cuda"
with torch.no_grad():
    x = torch.randn(64, 256, device=dev, dtype=torch.bfloat16)
    for et in dtypes.SUPPORTED_ELEM_DTYPES:
        m = MXTensor.to_mx(x, et, 32)
        m.to_dtype(torch.bfloat16); m.to_dtype(torch.float32)
        MXTensor.to_mx(x[:, :48].contiguous(), et, 16).to_dtype(torch.bfloat16)      # generic block size
        m.t().to_dtype(torch.bfloat16)                                               # transposed dequantize
        MXTensor.to_mx(torch.randn(2, 3, 40, 64, device=dev, dtype=torch.bfloat16), et, 32).transpose(2, 3).to_dtype(torch.float32)
    for packed in (True, False):
        mx_gemm.set_packed_operands(packed)
        for (M, N, K, ea, eb, batch, bias) in [(300, 264, 384, "float8_e4m3", "float4_e2m1", 0, True), (130, 136, 256, "float6_e3m2", "float6_e2m3", 0, False),
                                               (17, 520, 1024, "float8_e4m3", "float6_e3m2", 0, True), (100, 130, 256, "float4_e2m1", "float4_e2m1", 2, False),
                                               (64, 100, 128, "float8_e4m3", "float8_e4m3", 0, False)]:
            lead = (batch,) if batch else ()
            A = MXTensor.to_mx(torch.randn(*lead, M, K, device=dev, dtype=torch.bfloat16), dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[ea], 32)
            B = MXTensor.to_mx(torch.randn(*lead, N, K, device=dev, dtype=torch.bfloat16), dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[eb], 32)
            b = torch.randn(N, device=dev, dtype=torch.bfloat16) if bias else None
            y = torch.bmm(A, B.transpose(1, 2)) if batch else torch.nn.functional.linear(A, B, b)
    mx_gemm.set_packed_operands(True)
    for s in ("1", "2", "4"):
        mx_gemm.overrides["split_k"] = int(s)
        A = MXTensor.to_mx(torch.randn(8, 2048, device=dev, dtype=torch.bfloat16), dtypes.float8_e4m3, 32)
        B = MXTensor.to_mx(torch.randn(200, 2048, device=dev, dtype=torch.bfloat16), dtypes.float6_e3m2, 32)
        torch.nn.functional.linear(A, B)
torch.cuda.synchronize()
print("sanitize_smoke done", mx_gemm.stats)
