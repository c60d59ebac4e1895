"""GPU tier: MX matmul (aten.mm / bmm / addmm / linear overrides).

Parity definition (DESIGN.md "matmul parity"): the reference computes dequantize(A) @ dequantize(B)^T
with a bf16 GEMM that accumulates in fp32 and rounds the result to bf16 once (torchmx/ops.py:18-19,
29-41).  The dequantized operands are exact, so the only freedom is the fp32 accumulation order.
Tolerance, written out: |out - ref| <= 2^-8 |ref| + 2^-18 * sum_k |a_k b_k|, with `ref` the same
contraction evaluated in fp64 (one bf16 ulp of the result + fp32 accumulation slack).  The fallback
path (operands that cannot use the tensor cores) must equal dequantize-then-aten-op EXACTLY, as the
reference's own test_bmm demands (tests/test_mx_tensor.py:266-289).
"""
import numpy as np
import pytest
import torch

from tests.util import bf16_tensor, bits_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def mx():
    import torchmx
    from torchmx_b200 import _C
    _C.lib()
    return torchmx


def _tol_check(out, A, B, bias=None, what=""):
    ad, bd = A.to_dtype(torch.float32).double(), B.to_dtype(torch.float32).double()
    ref = ad @ bd.transpose(-1, -2)
    S = ad.abs() @ bd.abs().transpose(-1, -2)
    if bias is not None:
        ref = ref + bias.double()
        S = S + bias.double().abs()
    err = (out.double() - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2.0 ** -18 * S + 1e-30
    bad = int((err > tol).sum())
    assert bad == 0, f"{what}: {bad}/{err.numel()} outside tolerance, worst err/S {(err / (S + 1e-30)).max().item():.3e}"
    assert not torch.isnan(out).any()


CASES = [
    # M, N, K, elem A, elem B, bias, batch, per-block exponent spread
    (128, 128, 128, "float8_e4m3", "float8_e4m3", False, 0, 0),
    (128, 256, 256, "float8_e4m3", "float6_e3m2", False, 0, 0),
    (256, 512, 512, "float8_e4m3", "float6_e3m2", False, 0, 8),
    (100, 200, 384, "float8_e4m3", "float4_e2m1", True, 0, 0),
    (1, 4096, 1024, "float8_e4m3", "float6_e3m2", False, 0, 0),
    (33, 130, 1024, "float6_e2m3", "float6_e3m2", True, 0, 20),
    (300, 77, 640, "float6_e3m2", "float6_e2m3", False, 0, 30),
    (512, 512, 128, "float8_e4m3", "float6_e3m2", False, 4, 0),
    (77, 300, 256, "float4_e2m1", "float4_e2m1", False, 3, 0),
    (1024, 2048, 2048, "float8_e4m3", "float6_e3m2", False, 0, 4),
    # CTA-pair kernel (M > 128 and N > 128): ragged M / N with the TMA-store epilogue (N % 8 == 0) and the direct-store
    # fallback (N % 8 != 0), bias, batches, a K that is not a multiple of the 4-K-block scale ring
    (300, 520, 640, "float8_e4m3", "float6_e3m2", True, 0, 6),
    (257, 129, 384, "float8_e4m3", "float4_e2m1", True, 0, 0),
    (640, 300, 256, "float6_e3m2", "float6_e2m3", False, 2, 0),
    (2048, 2048, 128, "float8_e4m3", "float6_e3m2", False, 3, 0),
    # skinny weight-streaming kernel (M <= 128, not batched): all token-tile widths, cluster split-K, ragged N
    (32, 4096, 4096, "float8_e4m3", "float6_e3m2", True, 0, 6),
    (17, 1000, 1024, "float8_e4m3", "float4_e2m1", False, 0, 0),
    (64, 1024, 4096, "float8_e4m3", "float6_e3m2", True, 0, 0),
    (100, 384, 2048, "float6_e2m3", "float6_e3m2", False, 0, 12),
    (128, 7168, 1024, "float8_e4m3", "float6_e3m2", False, 0, 0),
    (5, 2048, 7168, "float8_e4m3", "float6_e3m2", True, 0, 0),
    (8, 640, 384, "float8_e4m3", "float8_e4m3", False, 0, 0),
    # small grids with a long K loop (M > 128): 128 x 128 tiles of the single-CTA kernel (under-filled pair grid)
    (300, 384, 4096, "float8_e4m3", "float6_e3m2", True, 0, 6),
    (257, 129, 2048, "float8_e4m3", "float4_e2m1", True, 0, 0),
    (1024, 1024, 4096, "float8_e4m3", "float6_e3m2", False, 0, 10),
    (640, 200, 8192, "float6_e2m3", "float8_e5m2", False, 0, 0),
    # float8_e5m2 (labelled extension element type) as a native one-byte operand, all three kernels
    (300, 520, 640, "float8_e5m2", "float6_e3m2", True, 0, 10),
    (48, 1000, 512, "float8_e4m3", "float8_e5m2", False, 0, 0),
    (96, 200, 256, "float8_e5m2", "float8_e5m2", False, 2, 0),
]


# shapes whose 256x256 pair tiles would under-fill the GPU are dispatched to 128x128 tiles; "pair" switches that rule off so the
# CTA-pair kernel keeps its small-shape coverage (ragged edges, bias, batches)
TILE_RULES = [(*c, rule) for c in CASES for rule in (("auto", "pair") if c[0] > 128 and c[1] > 128 else ("auto",))]


@pytest.mark.parametrize("M,N,K,ea,eb,bias,batch,spread,rule", TILE_RULES)
def test_tensor_core_matmul(mx, monkeypatch, M, N, K, ea, eb, bias, batch, spread, rule):
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import mx_gemm
    monkeypatch.setitem(mx_gemm.overrides, "wide_tiles", rule == "pair")  # mxq_gemm_args_t.flags: MXQ_GEMM_WIDE_TILES
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    sa, sb = ((batch, M, K), (batch, N, K)) if batch else ((M, K), (N, K))
    a = torch.randn(*sa, device=DEV, dtype=torch.bfloat16, generator=g)
    b = torch.randn(*sb, device=DEV, dtype=torch.bfloat16, generator=g)
    if spread:
        for t in (a, b):
            e = torch.randint(-spread, spread, (*t.shape[:-1], K // 32), device=DEV, generator=g).float()
            t *= torch.exp2(e).repeat_interleave(32, -1).to(torch.bfloat16)
    A = MXTensor.to_mx(a, dtypes.STR_TO_ELEM_DTYPE[ea], 32)
    B = MXTensor.to_mx(b, dtypes.STR_TO_ELEM_DTYPE[eb], 32)
    bias_t = torch.randn(N, device=DEV, dtype=torch.bfloat16, generator=g) if bias else None
    before = dict(mx_gemm.stats)
    if batch:
        out = torch.bmm(A, B.transpose(1, 2))
    else:
        out = torch.nn.functional.linear(A, B, bias_t)
    assert mx_gemm.stats["tensor_core"] == before["tensor_core"] + 1, "expected the tcgen05 path"
    assert out.dtype == torch.bfloat16 and out.shape == ((batch, M, N) if batch else (M, N))
    _tol_check(out, A, B, bias_t, f"{M}x{N}x{K} {ea}x{eb}")


@pytest.mark.parametrize("M,N,K,ea,eb", [(64, 96, 256, "float8_e4m3", "float6_e3m2"), (200, 136, 384, "float8_e4m3", "float4_e2m1"),
                                         (16, 520, 1024, "float6_e2m3", "float8_e4m3"), (130, 260, 128, "float4_e2m1", "float4_e2m1")])
def test_mx_linear_against_the_oracle_end_to_end(mx, oracle, M, N, K, ea, eb):
    """The whole MX linear through the GPU path (K1 on both operands, tensor-core GEMM) against the CPU oracle end to end: the
    oracle quantizes the same bf16 bits, dequantizes them and contracts with double accumulation (oracle/mx_oracle.c
    mxo_gemm_nt_bf16) -- no GPU kernel of this repository on the checker's side."""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16, generator=g) * 2
    w = torch.randn(N, K, device=DEV, dtype=torch.bfloat16, generator=g)
    w *= torch.exp2(torch.randint(-6, 6, (N, K // 32), device=DEV, generator=g).float()).repeat_interleave(32, -1).to(torch.bfloat16)
    out = torch.nn.functional.linear(MXTensor.to_mx(x, dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[ea], 32), MXTensor.to_mx(w, dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[eb], 32))
    bits = lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16)
    deq = {}
    for name, t, el in (("a", x, ea), ("b", w, eb)):
        scales, codes = oracle.quantize(bits(t), el, 32)
        deq[name] = oracle.dequantize(codes, scales, el, 32, target="bf16")
    ref = oracle.gemm_nt(deq["a"], deq["b"]).astype(np.float64)
    S = np.abs(oracle.bf16_bits_to_f32(deq["a"]).astype(np.float64)) @ np.abs(oracle.bf16_bits_to_f32(deq["b"]).astype(np.float64)).T
    err = np.abs(out.double().cpu().numpy() - ref)
    tol = 2.0 ** -8 * np.abs(ref) + 2.0 ** -18 * S + 1e-30  # bf16 rounding of the output + fp32 accumulation order
    assert (err <= tol).all(), f"{int((err > tol).sum())} of {err.size} outside tolerance"


def test_matmul_entry_points_agree(mx):
    """mm, addmm, linear (2-D and 3-D input, inference_mode and no_grad) and 4-D matmul all reach the
    same kernel and must agree bit-for-bit with each other."""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(4, 48, 256, device=DEV, dtype=torch.bfloat16, generator=g)
    w = torch.randn(320, 256, device=DEV, dtype=torch.bfloat16, generator=g)
    bias = torch.randn(320, device=DEV, dtype=torch.bfloat16, generator=g)
    X = MXTensor.to_mx(x, dtypes.float8_e4m3, 32)
    X2 = MXTensor.to_mx(x.reshape(-1, 256), dtypes.float8_e4m3, 32)
    W = MXTensor.to_mx(w, dtypes.float6_e3m2, 32)
    with torch.no_grad():
        y_lin3 = torch.nn.functional.linear(X, W, bias)
        y_lin2 = torch.nn.functional.linear(X2, W, bias)
        y_mm = torch.mm(X2, W.t())
        y_addmm = torch.addmm(bias, X2, W.t())
        y_nobias = torch.nn.functional.linear(X2, W)
    with torch.inference_mode():
        y_inf = torch.nn.functional.linear(X, W, bias)
    assert torch.equal(y_lin3.reshape(-1, 320), y_lin2)
    assert torch.equal(y_inf, y_lin3)
    assert torch.equal(y_addmm, y_lin2)
    assert torch.equal(y_mm, y_nobias)
    _tol_check(y_lin2, X2, W, bias, "linear")
    # attention-shaped 4-D matmuls (layers/mx_llama_attention.py:215-217, 243)
    q = torch.randn(2, 4, 64, 128, device=DEV, dtype=torch.bfloat16, generator=g)
    k = torch.randn(2, 4, 96, 128, device=DEV, dtype=torch.bfloat16, generator=g)
    Q, K = MXTensor.to_mx(q, dtypes.float8_e4m3, 32), MXTensor.to_mx(k, dtypes.float6_e3m2, 32)
    s = torch.matmul(Q, K.transpose(2, 3))
    assert s.shape == (2, 4, 64, 96)
    _tol_check(s, Q, K, None, "q @ k^T")
    # P @ V with V quantized along the sequence (block dim second to last after the transpose)
    pm = torch.rand(2, 4, 64, 128, device=DEV, dtype=torch.bfloat16, generator=g)
    v = torch.randn(2, 4, 128, 64, device=DEV, dtype=torch.bfloat16, generator=g)
    P = MXTensor.to_mx(pm, dtypes.float8_e4m3, 32)
    Vt = MXTensor.to_mx(v.transpose(2, 3).contiguous(), dtypes.float6_e3m2, 32)  # [.., 64, 128] blocked along seq
    o = torch.matmul(P, Vt.transpose(2, 3))
    _tol_check(o, P, Vt, None, "p @ v")


@pytest.mark.parametrize("splits", ["1", "2", "4", "8"])
def test_skinny_split_k_is_deterministic_and_consistent(mx, monkeypatch, splits):
    """The K splits of the decode kernel are reduced through DSMEM in a fixed order: repeated launches are
    bit-identical, and every split count stays within the matmul tolerance."""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import mx_gemm
    monkeypatch.setitem(mx_gemm.overrides, "split_k", int(splits))  # mxq_gemm_args_t.split_k
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(32, 4096, device=DEV, dtype=torch.bfloat16, generator=g)
    w = torch.randn(1536, 4096, device=DEV, dtype=torch.bfloat16, generator=g)
    X, W = MXTensor.to_mx(x, dtypes.float8_e4m3, 32), MXTensor.to_mx(w, dtypes.float6_e3m2, 32)
    y0 = torch.nn.functional.linear(X, W)
    for _ in range(3):
        assert torch.equal(torch.nn.functional.linear(X, W), y0)
    _tol_check(y0, X, W, None, f"skinny splits={splits}")


def test_weight_shadow_is_cached_and_invalidated(mx):
    """fp6 / fp4 weights are transcoded to the E4M3 container once per weight, not once per call (F.linear reaches the
    kernel as aten.t + aten.mm on short-lived views), and again after the codes change in place."""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import mx_gemm
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(64, 256, device=DEV, dtype=torch.bfloat16, generator=g)
    w = torch.randn(512, 256, device=DEV, dtype=torch.bfloat16, generator=g)
    X, W = MXTensor.to_mx(x, dtypes.float8_e4m3, 32), MXTensor.to_mx(w, dtypes.float6_e3m2, 32)
    n0 = mx_gemm.stats["transcode"]
    y = [torch.nn.functional.linear(X, W) for _ in range(4)] + [torch.mm(X, W.t())]
    assert mx_gemm.stats["transcode"] == n0 + 1
    assert all(torch.equal(y[0], t) for t in y[1:])
    W._data.copy_(MXTensor.to_mx(-w, dtypes.float6_e3m2, 32)._data)  # in-place update bumps the version counter
    y2 = torch.nn.functional.linear(X, W)
    assert mx_gemm.stats["transcode"] == n0 + 2
    assert torch.equal(y2, -y[0])


@pytest.mark.parametrize("ea,ew", [("float8_e4m3", "float6_e3m2"), ("float8_e4m3", "float4_e2m1"), ("float6_e2m3", "float6_e3m2"),
                                   ("float4_e2m1", "float4_e2m1"), ("int8", "int8"), ("float8_e4m3", "float8_e4m3")])
def test_fallback_matches_reference_fixture(mx, fixtures, ea, ew):
    """K = 96 / 64 cannot use the block-scaled tensor-core path (K % 128 != 0): the fused dequantize GEMM (K3d), compared
    with the reference's CPU outputs.  Operands are bit-identical; the bf16 GEMM is tcgen05 kind::f16 here and MKL there, so
    equality is up to accumulation order (same tolerance as above)."""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    a, w = bf16_tensor(fixtures["mm/a"], DEV), bf16_tensor(fixtures["mm/w"], DEV)
    bias = bf16_tensor(fixtures["mm/bias"], DEV)
    q, k = bf16_tensor(fixtures["mm/q"], DEV), bf16_tensor(fixtures["mm/k"], DEV)
    A, W = MXTensor.to_mx(a, dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[ea], 32), MXTensor.to_mx(w, dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[ew], 32)
    Q, K = MXTensor.to_mx(q, dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[ea], 32), MXTensor.to_mx(k, dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[ew], 32)
    tag = f"mm/{ea}x{ew}"
    with torch.no_grad():
        outs = {"linear": torch.nn.functional.linear(A, W), "linear_bias": torch.nn.functional.linear(A, W, bias),
                "mm": torch.matmul(A, W.t()), "qk": torch.matmul(Q, K.transpose(2, 3))}
    for name, out in outs.items():
        ref = bf16_tensor(fixtures[f"{tag}/{name}"], DEV).float()
        # one bf16 ulp of the result plus fp32-accumulation slack on ~100 products of O(10) magnitude
        torch.testing.assert_close(out.float(), ref, rtol=2 ** -7, atol=2e-2 if "int8" not in ea else 0.5)
    # and equal to dequantize-then-matmul on this device (reference: tests/test_mx_tensor.py:289 demands exact equality of its
    # own two cuBLAS calls; two different fp32 summation orders agree to one bf16 ulp, and on almost every element exactly)
    for got, want in ((outs["mm"], torch.matmul(A.to_dtype(torch.bfloat16), W.to_dtype(torch.bfloat16).t())),
                      (outs["qk"], torch.matmul(Q.to_dtype(torch.bfloat16), K.to_dtype(torch.bfloat16).transpose(2, 3)))):
        diff = (got.float() - want.float()).abs()
        assert (diff <= 2.0 ** -7 * want.float().abs() + 1e-6).all()
        assert (got == want).float().mean().item() > 0.98


def test_mx_inference_linear_and_quantize_linear(mx):
    """quantize_linear_ swaps every nn.Linear; MXInferenceLinear.forward = activation quantize + MX matmul
    (reference: quant_api.py:188-215, layers/mx_linear.py:61-95)."""
    from torchmx.config import MXConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx.mx_tensor import MXTensor
    from torchmx.quant_api import quantize_linear_
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(256, 512, bias=True), torch.nn.GELU(), torch.nn.Linear(512, 128, bias=False)).to(DEV, torch.bfloat16)
    ref = [m for m in model if isinstance(m, torch.nn.Linear)]
    w0, b0, w1 = ref[0].weight.data.clone(), ref[0].bias.data.clone(), ref[1].weight.data.clone()
    qc = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    quantize_linear_(model, qc)
    assert type(model[0]) is MXInferenceLinear and type(model[2]) is MXInferenceLinear
    assert isinstance(model[0].weight.data, MXTensor) and model[0].weight.shape == (512, 256)
    x = torch.randn(3, 40, 256, device=DEV, dtype=torch.bfloat16)
    y = model(x)
    assert y.shape == (3, 40, 128) and y.dtype == torch.bfloat16
    # same computation spelled out with dequantized operands
    from torchmx import dtypes
    xq = MXTensor.to_mx(x, dtypes.float8_e4m3, 32).to_dtype(torch.float32)
    h = torch.nn.functional.gelu((xq @ MXTensor.to_mx(w0, dtypes.float6_e3m2, 32).to_dtype(torch.float32).t() + b0.float()).to(torch.bfloat16))
    hq = MXTensor.to_mx(h, dtypes.float8_e4m3, 32).to_dtype(torch.float32)
    want = (hq @ MXTensor.to_mx(w1, dtypes.float6_e3m2, 32).to_dtype(torch.float32).t())
    sqnr = 20 * torch.log10(want.norm() / (want - y.float()).norm())
    assert sqnr > 35, sqnr
    # state_dict round trip keeps the MXTensor storage (reference: mx_tensor.py:526-528)
    sd = model.state_dict()
    assert isinstance(sd["0.weight"], MXTensor)


def test_pack_operand_is_an_exact_bit_permutation(mx):
    """mxq_pack_operand: fp4 = nibble swap of every byte (reference keeps the even element in the high nibble, the TMA /
    tensor-core order is low nibble first), fp6 = 16 one-byte codes -> 12 bytes, element i in bits [6i, 6i+6)."""
    from torchmx import dtypes
    from torchmx_b200 import _C
    from torchmx_b200.mx_tensor import _stream_ptr
    g = torch.Generator(device=DEV).manual_seed(21)
    n = 128 * 96
    c6 = torch.randint(0, 64, (n,), device=DEV, dtype=torch.uint8, generator=g)
    out6 = torch.empty(n * 6 // 8, device=DEV, dtype=torch.uint8)
    _C.check(_C.lib().mxq_pack_operand(c6.data_ptr(), dtypes.ELEM_ID["float6_e3m2"], n, out6.data_ptr(), 0, _stream_ptr(c6)), "pack6")
    bits = ((c6.cpu().numpy()[:, None] >> np.arange(6)) & 1).astype(np.uint8).reshape(-1)      # LSB-first bit stream
    want6 = np.packbits(bits, bitorder="little")
    assert np.array_equal(out6.cpu().numpy(), want6)
    c4 = torch.randint(0, 256, (n // 2,), device=DEV, dtype=torch.uint8, generator=g)
    out4 = torch.empty_like(c4)
    _C.check(_C.lib().mxq_pack_operand(c4.data_ptr(), dtypes.ELEM_ID["float4_e2m1"], n, out4.data_ptr(), 0, _stream_ptr(c4)), "pack4")
    h = c4.cpu().numpy()
    assert np.array_equal(out4.cpu().numpy(), ((h << 4) | (h >> 4)).astype(np.uint8))


@pytest.mark.parametrize("wdt", ["float4_e2m1", "float6_e3m2", "float6_e2m3"])
def test_packed_and_container_operands_agree_bit_for_bit(mx, wdt):
    """The packed 4 / 6-bit operand path and the E4M3-container path feed the tensor core the same values: identical outputs
    (same kernel, same accumulation order) for the pair, single-CTA and skinny kernels."""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import mx_gemm
    g = torch.Generator(device=DEV).manual_seed(8)
    for (M, N, K, batch) in [(300, 520, 640, 0), (48, 1000, 1024, 0), (96, 200, 256, 3)]:
        sa, sb = ((batch, M, K), (batch, N, K)) if batch else ((M, K), (N, K))
        X = MXTensor.to_mx(torch.randn(*sa, device=DEV, dtype=torch.bfloat16, generator=g), dtypes.float8_e4m3, 32)
        outs = []
        for packed in (True, False):
            mx_gemm.set_packed_operands(packed)
            try:
                W = MXTensor.to_mx(torch.randn(*sb, device=DEV, dtype=torch.bfloat16, generator=torch.Generator(device=DEV).manual_seed(9)),
                                   dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[wdt], 32)
                outs.append(torch.bmm(X, W.transpose(1, 2)) if batch else torch.nn.functional.linear(X, W))
            finally:
                mx_gemm.set_packed_operands(True)
        assert torch.equal(outs[0], outs[1]), (M, N, K, batch)


@pytest.mark.parametrize("rows,N,K,wdt,bias", [(32, 1536, 4096, "float6_e3m2", True), (1, 640, 1024, "float4_e2m1", False), (64, 4096, 2048, "float8_e4m3", True),
                                              (17, 1000, 512, "float6_e2m3", False), (48, 256, 14336, "float6_e3m2", False)])
@pytest.mark.parametrize("mode", ["False", "True"])
def test_fused_activation_quantization_is_bit_identical(mx, rows, N, K, wdt, bias, mode):
    """MXInferenceLinear.forward with the activation quantized inside the decode GEMM == quantize_mx (K1) followed by the MX
    matmul, bit for bit -- including blocks that hold Inf / NaN (scale 255) under both values of the hw_exact toggle."""
    from torchmx import env_variables as env
    from torchmx.config import MXConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx_b200 import mx_gemm
    env.MX_EXACT_QUANTIZATION = mode
    g = torch.Generator(device=DEV).manual_seed(rows * 31 + N)
    lin = torch.nn.Linear(K, N, bias=bias).to(DEV, torch.bfloat16)
    layer = MXInferenceLinear.from_float(lin, QLinearConfig(weights_config=MXConfig(wdt, 32), activations_config=MXConfig("float8_e4m3", 32)))
    x = torch.randn(rows, K, device=DEV, dtype=torch.bfloat16, generator=g)
    x *= torch.exp2(torch.randint(-12, 12, (rows, K // 32), device=DEV, generator=g).float()).repeat_interleave(32, -1).to(torch.bfloat16)
    x[0, 5] = float("inf")
    x[-1, K - 3] = float("nan")
    n0 = mx_gemm.stats.get("fused_act_quant", 0)
    y_fused = layer(x)
    assert mx_gemm.stats.get("fused_act_quant", 0) == n0 + 1, "expected the fused path"
    old = mx_gemm._FUSED_ACT
    mx_gemm._FUSED_ACT = False
    try:
        y_two = layer(x)
    finally:
        mx_gemm._FUSED_ACT = old
    assert mx_gemm.stats.get("fused_act_quant", 0) == n0 + 1
    assert torch.equal(torch.nan_to_num(y_fused.float(), nan=12345.0), torch.nan_to_num(y_two.float(), nan=12345.0))
    assert torch.isnan(y_fused[0]).all() and torch.isnan(y_fused[-1]).all() and (rows < 3 or not torch.isnan(y_fused[1]).any())


@pytest.mark.parametrize("adt", ["float6_e3m2", "float6_e2m3", "float4_e2m1"])
@pytest.mark.parametrize("mode", ["False", "True"])
def test_quantize_straight_into_the_operand_layout(mx, adt, mode):
    """mxq_quantize with MXQ_FLAG_OPERAND_LAYOUT == mxq_quantize followed by mxq_pack_operand, bit for bit (codes and scales),
    NaN / Inf blocks included, under both values of the hw_exact toggle; refused for one-byte element types"""
    from torchmx import dtypes
    from torchmx import env_variables as env
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import _C, mx_gemm
    from torchmx_b200.mx_tensor import _stream_ptr
    env.MX_EXACT_QUANTIZATION = mode
    et = dtypes.STR_TO_SUPPORTED_ELEM_DTYPE[adt]
    g = torch.Generator(device=DEV).manual_seed(77)
    rows, K = 333, 1024
    x = torch.randn(rows, K, device=DEV, dtype=torch.bfloat16, generator=g)
    x *= torch.exp2(torch.randint(-30, 30, (rows, K // 32), device=DEV, generator=g).float()).repeat_interleave(32, -1).to(torch.bfloat16)
    x[3, 40] = float("inf"); x[7, 999] = float("nan"); x[9, :64] = 0
    ref = MXTensor.to_mx(x, et, 32)
    want, fmt = mx_gemm._operand_rows(ref._data, et, None, packed=True)
    bits = mx_gemm._PACKED_BITS[fmt]
    got = torch.empty(rows, K * bits // 8, device=DEV, dtype=torch.uint8)
    sc = torch.empty(rows, K // 32, device=DEV, dtype=torch.uint8)
    flags = _C.FLAG_OPERAND_LAYOUT | (_C.FLAG_HW_EXACT if mode == "True" else 0)
    _C.check(_C.lib().mxq_quantize(x.data_ptr(), _C.HP_BF16, rows * K // 32, 32, dtypes.ELEM_ID[adt], flags, got.data_ptr(), sc.data_ptr(), 0, _stream_ptr(x)), "quantize")
    assert torch.equal(sc, ref._scale_e8m0) and torch.equal(got, want.view(rows, -1))
    rc = _C.lib().mxq_quantize(x.data_ptr(), _C.HP_BF16, rows * K // 32, 32, dtypes.ELEM_ID["float8_e4m3"], flags, got.data_ptr(), sc.data_ptr(), 0, _stream_ptr(x))
    assert rc == _C.ERR_UNSUPPORTED_SHAPE


@pytest.mark.parametrize("adt", ["float6_e3m2", "float6_e2m3", "float4_e2m1"])
@pytest.mark.parametrize("wdt", ["float8_e4m3", "float6_e3m2", "float4_e2m1"])
@pytest.mark.parametrize("rows,N,K,bias", [(300, 520, 640, True), (2048, 1024, 512, False), (40, 384, 256, True)])
def test_linear_with_4_and_6_bit_activations_is_two_launches_and_bit_identical(mx, adt, wdt, rows, N, K, bias):
    """MXInferenceLinear.forward with a 4 / 6-bit activation config (the reference's GEMM_COMBINATIONS, tests/layers/conftest.py:
    56-65): K1 writes the packed operand stream, the GEMM reads it -- no mxq_pack_operand launch for the activation -- and the
    result equals quantize -> pack -> GEMM bit for bit"""
    from torchmx.config import MXConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx_b200 import mx_gemm
    g = torch.Generator(device=DEV).manual_seed(rows + N)
    lin = torch.nn.Linear(K, N, bias=bias).to(DEV, torch.bfloat16)
    layer = MXInferenceLinear.from_float(lin, QLinearConfig(weights_config=MXConfig(wdt, 32), activations_config=MXConfig(adt, 32)))
    x = torch.randn(2, rows // 2, K, device=DEV, dtype=torch.bfloat16, generator=g) * 3
    layer(x)  # (the weight's operand shadow is made on first use)
    n0, t0 = mx_gemm.stats.get("packed_act_quant", 0), mx_gemm.stats["transcode"]
    y = layer(x)
    assert mx_gemm.stats.get("packed_act_quant", 0) == n0 + 1 and mx_gemm.stats["transcode"] == t0, "expected K1 -> GEMM without a pack launch"
    real = mx_gemm.linear_packed_act_quant
    mx_gemm.linear_packed_act_quant = lambda *a, **k: None
    try:
        y3 = layer(x)
    finally:
        mx_gemm.linear_packed_act_quant = real
    assert mx_gemm.stats["transcode"] == t0 + 1  # the three-launch path packs the activation
    assert y.shape == (2, rows // 2, N) and torch.equal(y, y3)


# ---- packed-only weights: PackedMXLinear / pack_linear_ / mxq_unpack_operand -----------------------------------
@pytest.mark.parametrize("wdt", ["float6_e3m2", "float6_e2m3", "float4_e2m1", "float8_e4m3"])
@pytest.mark.parametrize("rows", [8, 300])
def test_packed_linear_is_bit_identical_and_round_trips(wdt, rows):
    import io
    import torchmx  # noqa: F401
    from torchmx import mx_gemm
    from torchmx.config import MXConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx.layers.packed_linear import PackedMXLinear
    torch.manual_seed(1)
    lin = torch.nn.Linear(512, 384, bias=True, device=DEV, dtype=torch.bfloat16)
    qc = QLinearConfig(weights_config=MXConfig(wdt, 32), activations_config=MXConfig("float8_e4m3", 32))
    ref = MXInferenceLinear.from_float(lin, qc)
    x = torch.randn(rows, 512, device=DEV, dtype=torch.bfloat16)
    want = ref(x)
    codes, scales = ref.weight._data.clone(), ref.weight._scale_e8m0.clone()
    packed = PackedMXLinear.from_mx_linear(ref, keep_source=True)
    assert packed is not None
    bits = {"float6_e3m2": 6, "float6_e2m3": 6, "float4_e2m1": 4, "float8_e4m3": 8}[wdt]
    assert packed.weight_packed.shape == (384, 512 * bits // 8) and packed.weight_scale.shape == (384, 16)
    n0 = mx_gemm.stats["tensor_core"]
    got = packed(x)
    assert mx_gemm.stats["tensor_core"] == n0 + 1
    assert torch.equal(got, want)                      # same kernel, same operand bytes
    assert torch.equal(packed(ref.prepare_input(x)), want)
    # reference layout back, bit for bit (through the state_dict of the packed module)
    buf = io.BytesIO()
    torch.save(packed.state_dict(), buf)
    buf.seek(0)
    sd = torch.load(buf, weights_only=True)
    assert set(sd) == {"weight_packed", "weight_scale", "bias"} and sd["weight_packed"].numel() == 384 * 512 * bits // 8
    fresh = PackedMXLinear(512, 384, qc, packed.operand_format, bias=torch.nn.Parameter(torch.empty(384, device=DEV, dtype=torch.bfloat16)), device=DEV)
    fresh.load_state_dict(sd)
    back = fresh.to_mx_linear()
    assert torch.equal(back.weight._data, codes) and torch.equal(back.weight._scale_e8m0, scales)
    assert torch.equal(back(x), want)


@pytest.mark.parametrize("adt", ["float6_e3m2", "float4_e2m1"])
@pytest.mark.parametrize("wdt", ["float6_e3m2", "float4_e2m1"])
def test_packed_linear_with_4_and_6_bit_activations(adt, wdt):
    """a weight held only as its packed operand, activations quantized to 4 / 6 bits: K1 writes the packed activation stream
    (no pack launch), and the result equals the reference-layout layer's, bit for bit"""
    import torchmx  # noqa: F401
    from torchmx import mx_gemm
    from torchmx.config import MXConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx.layers.packed_linear import PackedMXLinear
    torch.manual_seed(2)
    lin = torch.nn.Linear(512, 384, bias=True, device=DEV, dtype=torch.bfloat16)
    qc = QLinearConfig(weights_config=MXConfig(wdt, 32), activations_config=MXConfig(adt, 32))
    ref = MXInferenceLinear.from_float(lin, qc)
    x = torch.randn(3, 100, 512, device=DEV, dtype=torch.bfloat16) * 2
    real = mx_gemm.linear_packed_act_quant
    mx_gemm.linear_packed_act_quant = lambda *a, **k: None  # the reference-layout layer on the three-launch path
    try:
        want = ref(x)
    finally:
        mx_gemm.linear_packed_act_quant = real
    packed = PackedMXLinear.from_mx_linear(ref, keep_source=True)
    n0, t0 = mx_gemm.stats.get("packed_act_quant", 0), mx_gemm.stats["transcode"]
    got = packed(x)
    assert mx_gemm.stats.get("packed_act_quant", 0) == n0 + 1 and mx_gemm.stats["transcode"] == t0
    assert torch.equal(got, want)


def test_pack_linear_on_a_model_and_what_it_leaves_alone():
    import torchmx  # noqa: F401
    from torchmx.config import MXConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx.layers.packed_linear import PackedMXLinear
    from torchmx.quant_api import pack_linear_, quantize_linear_, unpack_linear_
    torch.manual_seed(2)
    model = torch.nn.Sequential(torch.nn.Linear(256, 512, bias=False), torch.nn.GELU(), torch.nn.Linear(512, 96), torch.nn.Linear(96, 64)).to(DEV, torch.bfloat16)
    quantize_linear_(model, QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32)))
    x = torch.randn(40, 256, device=DEV, dtype=torch.bfloat16)
    want = model(x)
    sd_before = {k: (v._data.clone(), v._scale_e8m0.clone()) for k, v in model.state_dict().items() if k.endswith("weight")}
    mem0 = sum(m.weight._data.numel() + m.weight._scale_e8m0.numel() for m in model if isinstance(m, MXInferenceLinear))
    assert pack_linear_(model) == 2                                    # in_features 96 is not a multiple of 128: stays as it is
    assert [type(m) for m in model if not isinstance(m, torch.nn.GELU)] == [PackedMXLinear, PackedMXLinear, MXInferenceLinear]
    mem1 = sum(b.numel() for m in model if isinstance(m, PackedMXLinear) for b in m.buffers()) + model[3].weight._data.numel() + model[3].weight._scale_e8m0.numel()
    assert mem1 < 0.78 * mem0
    assert torch.equal(model(x), want)
    assert unpack_linear_(model) == 2
    for k, (codes, scales) in sd_before.items():
        w = model.state_dict()[k]
        assert torch.equal(w._data, codes) and torch.equal(w._scale_e8m0, scales)
    assert torch.equal(model(x), want)


# ---- special values (reference: tests/layers/test_mx_linear.py:117-175, tests/test_mx_tensor.py:102-160, :243-263) --------------
@pytest.mark.parametrize("M,N,K,rule", [(4, 384, 256, "auto"), (64, 1000, 1024, "auto"), (200, 300, 256, "auto"), (300, 520, 384, "pair"),
                                        (300, 520, 384, "auto")])
@pytest.mark.parametrize("bias", [False, True])
def test_nan_and_inf_blocks_poison_the_same_outputs_as_the_dequantize_path(mx, monkeypatch, M, N, K, rule, bias):
    """an Inf / NaN anywhere in a 32-block gives that block the NaN scale (255); the reference then dequantizes the whole block to
    NaN and every output that contracts over it is NaN.  The block-scaled MMA must agree (E8M0 0xFF is NaN in hardware too)."""
    from torchmx import dtypes, mx_gemm
    from torchmx.mx_tensor import MXTensor
    monkeypatch.setitem(mx_gemm.overrides, "wide_tiles", rule == "pair")
    g = torch.Generator(device=DEV).manual_seed(5)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16, generator=g)
    w = torch.randn(N, K, device=DEV, dtype=torch.bfloat16, generator=g)
    a[1, 3] = float("inf")
    a[2, K - 1] = float("-inf")
    a[3, 40] = float("nan")
    w[5, 70] = float("nan")
    w[N - 1, 0] = float("inf")
    b = torch.randn(N, device=DEV, dtype=torch.bfloat16, generator=g) if bias else None
    A, W = MXTensor.to_mx(a, dtypes.float8_e4m3, 32), MXTensor.to_mx(w, dtypes.float6_e3m2, 32)
    n0 = mx_gemm.stats["tensor_core"]
    out = torch.nn.functional.linear(A, W, b)
    assert mx_gemm.stats["tensor_core"] == n0 + 1
    mx_gemm.set_enabled(False)
    try:
        ref = torch.nn.functional.linear(A, W, b)
    finally:
        mx_gemm.set_enabled(True)
    want_nan = torch.zeros(M, N, dtype=torch.bool, device=DEV)
    want_nan[1:4] = True
    want_nan[:, 5] = True
    want_nan[:, N - 1] = True
    assert torch.equal(torch.isnan(ref), want_nan)
    assert torch.equal(torch.isnan(out), want_nan)
    ok = ~want_nan
    assert (out[ok].float() - ref[ok].float()).abs().max().item() <= 2.0 ** -6 * ref[ok].float().abs().max().item()


@pytest.mark.parametrize("elem", ["float8_e4m3", "float6_e3m2", "float6_e2m3", "float4_e2m1", "int8"])
@pytest.mark.parametrize("hp_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("padding", [0, 3])
def test_cast_autograd_is_a_pass_through(mx, elem, hp_dtype, padding):
    """reference tests/test_mx_tensor.py:243-263: to_mx / to_dtype are identity in the backward pass"""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    et = dtypes.STR_TO_ELEM_DTYPE[elem]
    x = torch.arange(8 + padding, device=DEV, dtype=torch.bfloat16).requires_grad_()
    grad = torch.arange(8 + padding, device=DEV, dtype=torch.bfloat16) * 0.5
    x_dq = MXTensor.to_mx(x, et, 8).to_dtype(hp_dtype)
    if et == dtypes.float4_e2m1 and padding > 0:
        with pytest.raises(ValueError):
            x_dq.backward(gradient=grad)
    else:
        x_dq.backward(gradient=grad)
        assert torch.equal(grad, x.grad)


@pytest.mark.parametrize("elem,floor", [("float8_e4m3", 19.0), ("int8", 38.0), ("float6_e3m2", 14.0), ("float6_e2m3", 14.0), ("float4_e2m1", 14.0)])
@pytest.mark.parametrize("shape,block", [((128, 128), 32), ((4, 6, 64), 16), ((2, 2, 8, 96), 8), ((1024,), 32), ((3, 70), 7)])
def test_round_trip_sqnr_floors(mx, elem, floor, shape, block):
    """reference tests/test_mx_tensor.py:59-100: SQNR of to_mx -> to_dtype per element type"""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    torch.manual_seed(1234)
    x = torch.randn(*shape, device=DEV, dtype=torch.bfloat16)
    y = MXTensor.to_mx(x, dtypes.STR_TO_ELEM_DTYPE[elem], block).to_dtype(torch.bfloat16)
    sqnr = float(20 * torch.log10(x.float().norm() / (x.float() - y.float()).norm()))
    assert sqnr >= floor, sqnr
    z = MXTensor.to_mx(torch.zeros_like(x), dtypes.STR_TO_ELEM_DTYPE[elem], block).to_dtype(torch.bfloat16)
    assert torch.equal(z, torch.zeros_like(x))


@pytest.mark.parametrize("ew", ["float8_e4m3", "float6_e3m2"])
def test_decode_linear_right_after_weight_quantization_is_race_free(mx, ew):
    """The decode kernel is launched with programmatic stream serialization and the quantize kernel lets its dependents start
    early: a weight quantized IMMEDIATELY before a decode-sized linear (F.linear(to_mx(x), to_mx(w)), or a meta-weight layer
    that quantizes per forward, torchmx/layers/mx_linear.py:68-92) must be complete before the GEMM reads it.  Only a weight
    the caller declares static (MXQ_GEMM_B_STATIC; MXInferenceLinear does, once its quantization has finished) is prefetched
    before the grid dependency resolves."""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import mx_gemm
    g = torch.Generator(device=DEV).manual_seed(31)
    x = torch.randn(16, 4096, device=DEV, dtype=torch.bfloat16, generator=g)
    ws = [torch.randn(16384, 4096, device=DEV, dtype=torch.bfloat16, generator=g) for _ in range(3)]
    X = MXTensor.to_mx(x, dtypes.float8_e4m3, 32)
    et = dtypes.STR_TO_ELEM_DTYPE[ew]
    torch.cuda.synchronize()
    got = []
    for rep in range(4):
        for w in ws:  # quantize (128 MB written) and use at once, no synchronisation in between
            got.append(torch.nn.functional.linear(X, MXTensor.to_mx(w, et, 32)))
    torch.cuda.synchronize()
    mx_gemm.overrides["no_pdl"] = True
    try:
        want = []
        for w in ws:
            W = MXTensor.to_mx(w, et, 32)
            torch.cuda.synchronize()
            want.append(torch.nn.functional.linear(X, W))
    finally:
        mx_gemm.overrides["no_pdl"] = False
    for i, y in enumerate(got):
        assert torch.equal(y, want[i % 3]), f"launch {i}: the GEMM read a weight that was still being written"
