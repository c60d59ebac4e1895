"""GPU tier: the operator-level boundary under torch.compile / torch.library -- the reference's own checks
(tests/test_mx_tensor.py:359-523): `torch.compile(MXTensor.to_mx, fullgraph=True)` and the compiled `dequantize_mx` equal the
eager ops bit for bit, `to_mx` -> `to_dtype` traces as ONE graph without breaks, and both custom ops pass
`torch.library.opcheck` (schema, fake kernel, autograd registration, AOT dispatch).  The backend is `aot_eager`: dynamo, the fake
kernels and AOT autograd are exercised, no code generator stands between the graph and the hand-written kernels."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ELEMS = ["float8_e4m3", "float6_e3m2", "float6_e2m3", "float4_e2m1", "int8"]


@pytest.fixture(autouse=True)
def _fresh_dynamo():
    torch._dynamo.reset()
    yield
    torch._dynamo.reset()


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("hp_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("padding,block_size", [(0, 2), (1, 2), (0, 32)])
def test_compiled_ops_equal_eager(elem, hp_dtype, padding, block_size):
    import torchmx  # noqa: F401
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor, dequantize_mx
    et = dtypes.STR_TO_ELEM_DTYPE[elem]
    x = torch.randn(4, 4 * block_size + padding, dtype=torch.bfloat16, device=DEV)
    to_mx_c = torch.compile(MXTensor.to_mx, fullgraph=True, backend="aot_eager")
    x_mx = MXTensor.to_mx(x, et, block_size)
    x_mx_c = to_mx_c(x, et, block_size)
    assert torch.equal(x_mx._scale_e8m0, x_mx_c._scale_e8m0) and torch.equal(x_mx._data, x_mx_c._data)
    assert x_mx_c.shape == x.shape and x_mx_c._padding == x_mx._padding
    to_dtype_c = torch.compile(dequantize_mx, fullgraph=True, backend="aot_eager")
    data, data_c = x_mx._data, x_mx_c._data
    if padding > 0 and et != dtypes.float4_e2m1:
        data, data_c = F.pad(data, (0, padding), value=0), F.pad(data_c, (0, padding), value=0)
    a = dequantize_mx(data, x_mx._scale_e8m0, et.name, block_size, hp_dtype, x_mx._block_dim)
    b = to_dtype_c(data_c, x_mx_c._scale_e8m0, et.name, block_size, hp_dtype, x_mx_c._block_dim)
    assert a.dtype == hp_dtype and torch.equal(a, b)


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("shape", [(2, 4), (1, 4, 8), (1, 1, 8, 16), (2, 5), (1, 4, 9), (1, 1, 8, 17), (3, 128)])
def test_no_graph_breaks(elem, shape):
    import torchmx  # noqa: F401
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor
    et = dtypes.STR_TO_ELEM_DTYPE[elem]

    def there_and_back(x, elem_dtype, block_size):
        return MXTensor.to_mx(x, elem_dtype, block_size).to_dtype(x.dtype)

    x = torch.randn(*shape, dtype=torch.bfloat16, device=DEV)
    block_size = 32 if shape[-1] % 32 == 0 else 2
    explanation = torch._dynamo.explain(there_and_back)(x, et, block_size)
    assert explanation.graph_break_count == 0, f"Graph breaks: {explanation.graph_break_count} {explanation.break_reasons}"
    assert explanation.graph_count == 1, f"Graphs: {explanation.graph_count}"


@pytest.mark.parametrize("elem", ELEMS)
def test_quantize_mx_registered(elem):
    import torchmx  # noqa: F401
    from torchmx.mx_tensor import quantize_mx
    x = torch.randn(4, 64, dtype=torch.bfloat16, device=DEV)
    for block_size in (2, 32):
        result = torch.library.opcheck(quantize_mx, (x, elem, block_size))
        assert all(v == "SUCCESS" for v in result.values()), result


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("hp_dtype", [torch.float32, torch.bfloat16])
def test_dequantize_mx_registered(elem, hp_dtype):
    import torchmx  # noqa: F401
    from torchmx import dtypes
    from torchmx.mx_tensor import MXTensor, dequantize_mx
    x = torch.randn(4, 64, dtype=torch.bfloat16, device=DEV)
    for block_size in (2, 32):
        m = MXTensor.to_mx(x, dtypes.STR_TO_ELEM_DTYPE[elem], block_size)
        result = torch.library.opcheck(dequantize_mx, (m._data, m._scale_e8m0, elem, block_size, hp_dtype, m._block_dim))
        assert all(v == "SUCCESS" for v in result.values()), result


def test_compiled_linear_layer_matches_eager():
    """a quantized layer inside a compiled function: the MXTensor weight parameter, the activation quantization and the MX
    linear override trace and give the eager result"""
    import torchmx  # noqa: F401
    from torchmx.config import MXConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 384, bias=True, device=DEV, dtype=torch.bfloat16)
    q = MXInferenceLinear.from_float(lin, QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32)))
    x = torch.randn(160, 256, device=DEV, dtype=torch.bfloat16)
    want = q(x)
    got = torch.compile(q, backend="aot_eager")(x)
    assert got.shape == want.shape
    err = (got.float() - want.float()).abs().max().item()
    assert err <= 2.0 ** -6 * want.float().abs().max().item(), err
