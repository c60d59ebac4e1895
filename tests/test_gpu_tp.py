"""GPU tier: tensor-parallel MX linears (torchmx_b200/layers/tp_linear.py), the ranks of a world-size-2/4 group evaluated
one after the other on ONE device (the NCCL all-reduce itself is exercised by tools/tp_llama_bench.py on a multi-GPU box
and by the gloo test in tests/test_host_logic.py).

Parity: a shard's MX codes / scales are bit-identical to the corresponding slice of the unsharded MX tensor (blocks never
straddle a shard boundary); column-parallel outputs concatenate to exactly the unsharded output; row-parallel partials sum
to the unsharded contraction up to the bf16 rounding of each partial: |sum - ref| <= 2^-8 |ref| + 2^-8 sum_r |partial_r| +
2^-18 sum_k |a_k b_k| against the fp64 contraction of the dequantized operands.
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


def _qc(w="float4_e2m1", a="float8_e4m3"):
    from torchmx.config import MXConfig, QLinearConfig
    return QLinearConfig(weights_config=MXConfig(w, 32), activations_config=MXConfig(a, 32))


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("wdt", ["float4_e2m1", "float6_e3m2"])
def test_column_parallel_is_exact(mx, world, wdt):
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx_b200.layers.tp_linear import ColumnParallelMXLinear
    torch.manual_seed(0)
    lin = torch.nn.Linear(512, 1024, bias=True).to(DEV, torch.bfloat16)
    x = torch.randn(3, 40, 512, device=DEV, dtype=torch.bfloat16)
    qc = _qc(wdt)
    full = MXInferenceLinear.from_float(lin, qc)
    shards = [ColumnParallelMXLinear.from_float(lin, qc, world_rank=(world, r)) for r in range(world)]
    per = 1024 // world
    for r, s in enumerate(shards):
        assert s.weight.shape == (per, 512)
        assert torch.equal(s.weight._data, full.weight._data[r * per:(r + 1) * per])
        assert torch.equal(s.weight._scale_e8m0, full.weight._scale_e8m0[r * per:(r + 1) * per])
    y = torch.cat([s(x) for s in shards], dim=-1)
    assert torch.equal(y, full(x))


@pytest.mark.parametrize("world", [2, 4])
def test_row_parallel_partials_sum_to_the_full_layer(mx, world):
    from torchmx import dtypes
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx.mx_tensor import MXTensor
    from torchmx_b200.layers.tp_linear import RowParallelMXLinear, shard_bounds
    torch.manual_seed(1)
    K, N = 2048, 384
    lin = torch.nn.Linear(K, N, bias=True).to(DEV, torch.bfloat16)
    x = torch.randn(64, K, device=DEV, dtype=torch.bfloat16)
    qc = _qc("float4_e2m1")
    full = MXInferenceLinear.from_float(lin, qc)
    shards = [RowParallelMXLinear.from_float(lin, qc, world_rank=(world, r)) for r in range(world)]
    partials = []
    for r, s in enumerate(shards):
        lo, hi = shard_bounds(K, world, r, 32)
        # fp4 codes are packed two per byte along K
        assert torch.equal(s.weight._data, full.weight._data[:, lo // 2:hi // 2])
        assert torch.equal(s.weight._scale_e8m0, full.weight._scale_e8m0[:, lo // 32:hi // 32])
        assert (s.bias is not None) == (r == 0)
        s.tp_world = 1  # evaluate the local partial only; the reduction is done below (NCCL does it across GPUs)
        partials.append(s(x[:, lo:hi].contiguous()))
    got = torch.stack([p.float() for p in partials]).sum(0)
    X = MXTensor.to_mx(x, dtypes.float8_e4m3, 32).to_dtype(torch.float32).double()
    W = full.weight.to_dtype(torch.float32).double()
    ref = X @ W.t() + lin.bias.double()
    S = X.abs() @ W.abs().t() + lin.bias.double().abs()
    tol = 2.0 ** -8 * ref.abs() + 2.0 ** -8 * torch.stack([p.double().abs() for p in partials]).sum(0) + 2.0 ** -18 * S
    assert ((got.double() - ref).abs() <= tol).all()


def test_shard_bounds_rejects_uneven_splits():
    from torchmx_b200.layers.tp_linear import shard_bounds
    assert shard_bounds(8192, 8, 3, 32) == (3072, 4096)
    with pytest.raises(ValueError):
        shard_bounds(8192 + 32, 8, 0, 32)
    with pytest.raises(ValueError):
        shard_bounds(100, 3, 0)
