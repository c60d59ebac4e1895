"""GPU tier: the MX matmul at BASELINE.json configs[2] sizes -- 8192 x 8192 x 8192 (fp8_e4m3 activations x fp6_e3m2 weights,
also fp4 weights and fp4 x fp4 on kind::mxf4), the lm_head-sized 2048 x 128256 x 4096, and the 4-D attention contractions
Q @ K^T and P @ V at [1, 32, 2048, 128] -- through the public ops (F.linear / torch.matmul on MXTensors) with the default
dispatch, i.e. the persistent multi-tile loop of the CTA-pair kernel with a long K ring (accumulator slots alternating, scale
ring wrapping across tiles).

Parity (DESIGN.md "matmul parity"): against the fp64 contraction of the dequantized operands (reference recipe:
torchmx/ops.py:29-41, 60-68, 99-119), |out - ref| <= 2^-8 |ref| + 2^-18 * sum_k |a_k b_k|.  For the big GEMMs 96 output rows
(tile edges + random) x all columns are checked; the attention contractions are checked in full.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def mx():
    import torchmx
    from torchmx_b200 import _C
    _C.lib()
    return torchmx


def _check_rows(out, A, B, rows, what):
    ad = A.to_dtype(torch.float32)[rows].double()
    bd = B.to_dtype(torch.float32)
    ref = torch.empty(len(rows), bd.shape[0], dtype=torch.float64, device=out.device)
    S = torch.empty_like(ref)
    for lo in range(0, bd.shape[0], 16384):  # column panels keep the fp64 copy of B small
        panel = bd[lo:lo + 16384].double()
        ref[:, lo:lo + 16384] = ad @ panel.t()
        S[:, lo:lo + 16384] = ad.abs() @ panel.abs().t()
    err = (out[rows].double() - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2.0 ** -18 * S + 1e-30
    bad = int((err > tol).sum())
    assert bad == 0, f"{what}: {bad}/{err.numel()} outside tolerance, worst err/S {(err / (S + 1e-30)).max().item():.3e}"
    assert not torch.isnan(out).any()


def _sample_rows(M, g):
    edges = [0, 1, 127, 128, 255, 256, 257, 383, 384, 511, 512, M // 2 - 1, M // 2, M - 257, M - 256, M - 129, M - 128, M - 1]
    rnd = torch.randint(0, M, (96 - len(edges),), generator=g).tolist()
    return torch.tensor(sorted(set(e for e in edges if 0 <= e < M) | set(rnd)), device=DEV)


@pytest.mark.parametrize("M,N,K,ea,eb,spread", [
    (8192, 8192, 8192, "float8_e4m3", "float6_e3m2", 0),   # BASELINE configs[2]
    (8192, 8192, 8192, "float8_e4m3", "float4_e2m1", 6),
    (8192, 8192, 8192, "float4_e2m1", "float4_e2m1", 4),   # kind::mxf4
    (2048, 128256, 4096, "float8_e4m3", "float6_e3m2", 0),  # Llama-3 lm_head at prefill 2048
    (4096, 14336, 4096, "float4_e2m1", "float4_e2m1", 0),   # kind::mxf4, N not a multiple of the wave
])
def test_config3_gemm_sampled_rows(mx, M, N, K, ea, eb, spread):
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import mx_gemm
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16, generator=g)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16, generator=g)
    if spread:
        for t in (a, b):
            e = torch.randint(-spread, spread, (t.shape[0], K // 32), device=DEV, generator=g).float()
            t *= torch.exp2(e).repeat_interleave(32, -1).to(torch.bfloat16)
    A = MXTensor.to_mx(a, dtypes.STR_TO_ELEM_DTYPE[ea], 32)
    B = MXTensor.to_mx(b, dtypes.STR_TO_ELEM_DTYPE[eb], 32)
    del a, b
    before = dict(mx_gemm.stats)
    out = torch.nn.functional.linear(A, B)
    assert mx_gemm.stats["tensor_core"] == before["tensor_core"] + 1, "expected the tcgen05 block-scaled path"
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    cg = torch.Generator().manual_seed(5)
    _check_rows(out, A, B, _sample_rows(M, cg), f"{M}x{N}x{K} {ea} x {eb}")
    # a second launch over the same operands is bit-identical (persistent tile order / slot alternation is deterministic)
    assert torch.equal(torch.nn.functional.linear(A, B), out)


@pytest.mark.parametrize("eq,ek,ep,ev", [("float8_e4m3", "float6_e3m2", "float8_e4m3", "float6_e3m2"),
                                         ("float8_e4m3", "float8_e4m3", "float8_e4m3", "float8_e4m3"),
                                         ("float4_e2m1", "float4_e2m1", "float4_e2m1", "float4_e2m1")])
def test_config3_attention_bmm_full(mx, eq, ek, ep, ev):
    """Q @ K^T: [1, 32, 2048, 128] x [1, 32, 2048, 128]^T -> bf16 [1, 32, 2048, 2048]; P @ V with V quantized along the
    key/value sequence (torchmx/layers/mx_llama_attention.py:209-213): [1, 32, 2048, 2048] x [1, 32, 2048, 128]"""
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import mx_gemm
    g = torch.Generator(device=DEV).manual_seed(21)
    q = torch.randn(1, 32, 2048, 128, device=DEV, dtype=torch.bfloat16, generator=g)
    k = torch.randn(1, 32, 2048, 128, device=DEV, dtype=torch.bfloat16, generator=g)
    v = torch.randn(1, 32, 2048, 128, device=DEV, dtype=torch.bfloat16, generator=g)
    E = dtypes.STR_TO_ELEM_DTYPE
    Q, Kx = MXTensor.to_mx(q, E[eq], 32), MXTensor.to_mx(k, E[ek], 32)
    before = dict(mx_gemm.stats)
    scores = torch.matmul(Q, Kx.transpose(2, 3))
    assert mx_gemm.stats["tensor_core"] == before["tensor_core"] + 1
    assert scores.shape == (1, 32, 2048, 2048)
    qd, kd = Q.to_dtype(torch.float32).double(), Kx.to_dtype(torch.float32).double()
    ref = qd @ kd.transpose(2, 3)
    S = qd.abs() @ kd.abs().transpose(2, 3)
    err = (scores.double() - ref).abs()
    assert int((err > 2.0 ** -8 * ref.abs() + 2.0 ** -18 * S + 1e-30).sum()) == 0
    del ref, S, err
    p = torch.softmax(scores.float() * 128 ** -0.5, -1).to(torch.bfloat16)
    P = MXTensor.to_mx(p, E[ep], 32)
    V = MXTensor.to_mx(v.transpose(2, 3).contiguous(), E[ev], 32).transpose(2, 3)
    before = dict(mx_gemm.stats)
    out = torch.matmul(P, V)
    assert mx_gemm.stats["tensor_core"] == before["tensor_core"] + 1
    assert out.shape == (1, 32, 2048, 128)
    pd, vd = P.to_dtype(torch.float32).double(), V.to_dtype(torch.float32).double()
    ref = pd @ vd
    S = pd.abs() @ vd.abs()
    err = (out.double() - ref).abs()
    assert int((err > 2.0 ** -8 * ref.abs() + 2.0 ** -18 * S + 1e-30).sum()) == 0


# kind::mxf4 (fp4 x fp4, K % 256 == 0, CTA-pair kernel): ragged edges, bias, batches, every scale-ring phase (K / 128 = 2, 4, 6, 10, 32)
MXF4_CASES = [(300, 520, 256, True, 0), (257, 264, 512, False, 0), (640, 300, 768, True, 2), (384, 512, 1280, False, 0),
              (1024, 2048, 4096, False, 0), (2048, 2048, 256, False, 3)]


@pytest.mark.parametrize("M,N,K,bias,batch", MXF4_CASES)
def test_mxf4_kernel_matches_contraction_and_the_mxf8f6f4_path(mx, monkeypatch, M, N, K, bias, batch):
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200 import mx_gemm
    monkeypatch.setitem(mx_gemm.overrides, "wide_tiles", True)  # keep the CTA-pair kernel on these small grids
    g = torch.Generator(device=DEV).manual_seed(M + 3 * N + K)
    sa, sb = ((batch, M, K), (batch, N, K)) if batch else ((M, K), (N, K))
    a = torch.randn(*sa, device=DEV, dtype=torch.bfloat16, generator=g)
    b = torch.randn(*sb, device=DEV, dtype=torch.bfloat16, generator=g)
    for t in (a, b):
        e = torch.randint(-10, 10, (*t.shape[:-1], K // 32), device=DEV, generator=g).float()
        t *= torch.exp2(e).repeat_interleave(32, -1).to(torch.bfloat16)
    A, B = MXTensor.to_mx(a, dtypes.float4_e2m1, 32), MXTensor.to_mx(b, dtypes.float4_e2m1, 32)
    bias = bias and not batch  # (bmm takes no bias)
    bias_t = torch.randn(N, device=DEV, dtype=torch.bfloat16, generator=g) if bias else None

    def run():
        return torch.bmm(A, B.transpose(1, 2)) if batch else torch.nn.functional.linear(A, B, bias_t)

    before = dict(mx_gemm.stats)
    out = run()
    assert mx_gemm.stats["tensor_core"] == before["tensor_core"] + 1
    ad, bd = A.to_dtype(torch.float32).double(), B.to_dtype(torch.float32).double()
    ref = ad @ bd.transpose(-1, -2)
    S = ad.abs() @ bd.abs().transpose(-1, -2)
    if bias:
        ref, S = ref + bias_t.double(), S + bias_t.double().abs()
    err = (out.double() - ref).abs()
    assert int((err > 2.0 ** -8 * ref.abs() + 2.0 ** -18 * S + 1e-30).sum()) == 0
    # same operands through kind::mxf8f6f4 (MXQ_GEMM_NO_MXF4): same products, at most the accumulation order differs
    monkeypatch.setitem(mx_gemm.overrides, "no_mxf4", True)
    out8 = run()
    assert (out.float() - out8.float()).abs().max().item() <= 2.0 ** -7 * out8.float().abs().max().item()
    assert (out == out8).float().mean().item() > 0.97
