"""GPU tier: K5 (csrc/mxq_glue.cu) -- the RMSNorm (+ residual add, + MX quantization) and rotary-embedding launches of a
decoder layer, against the transformers modules the reference's layers call (torchmx/layers/mx_llama_attention.py:171-187) and
against K1 on the kernel's own bf16 output."""
import numpy as np
import pytest
import torch

from tests.util import ELEMS, assert_bits_equal, bits_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ulps(a, b):
    """distance in bf16 steps between two bf16 tensors of the same sign pattern (sign-magnitude -> monotone integer)"""
    def key(t):
        v = t.view(torch.int16).int()
        return torch.where(v < 0, -(v & 0x7FFF), v)
    return (key(a) - key(b)).abs()


@pytest.mark.parametrize("shape", [(1, 7, 512), (3, 4096), (2, 33, 8192), (5, 96), (2, 16384), (2048, 4096), (1, 1100, 512), (1024, 8192), (1030, 8224)])
@pytest.mark.parametrize("residual", [False, True])
def test_rmsnorm_matches_the_transformers_module(shape, residual):
    from transformers.models.llama.modeling_llama import LlamaRMSNorm
    import torchmx  # noqa: F401
    from torchmx_b200 import glue_ops
    g = torch.Generator(device=DEV).manual_seed(sum(shape))
    x = torch.randn(*shape, device=DEV, dtype=torch.bfloat16, generator=g) * 3
    res = torch.randn(*shape, device=DEV, dtype=torch.bfloat16, generator=g) if residual else None
    mod = LlamaRMSNorm(shape[-1], eps=1e-5).to(DEV, torch.bfloat16)
    mod.weight.data = torch.randn(shape[-1], device=DEV, dtype=torch.bfloat16, generator=g)
    y, mx_none, h = glue_ops.rmsnorm(x, mod.weight, 1e-5, residual=res)
    assert mx_none is None
    if residual:
        assert torch.equal(h, x + res)  # the residual stream: one bf16 add, exact
    else:
        assert h is None
    want = mod(x + res if residual else x)
    # the row statistic is summed in another order than torch's reduction: rsqrt may differ in its last fp32 bit, which moves a
    # bf16 rounding of the normalised value by one step for a few elements -- never more
    d = _ulps(y, want)
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 2e-3, (int(d.max()), float((d > 0).float().mean()))


@pytest.mark.parametrize("elem", ELEMS + ["float8_e5m2"])
@pytest.mark.parametrize("mode", ["False", "True"])
@pytest.mark.parametrize("rows,hidden", [(67, 4096), (1500, 4096), (1024, 3072)])
def test_rmsnorm_quantized_output_is_k1_of_its_own_bf16_output(elem, mode, rows, hidden):
    import torchmx  # noqa: F401
    from torchmx_b200 import dtypes, glue_ops
    from torchmx_b200 import env_variables as env
    from torchmx_b200.mx_tensor import MXTensor
    env.MX_EXACT_QUANTIZATION = mode
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(rows, hidden, device=DEV, dtype=torch.bfloat16, generator=g)
    x *= torch.exp2(torch.randint(-30, 30, (rows, hidden // 32), device=DEV, generator=g).float()).repeat_interleave(32, -1).to(torch.bfloat16)
    x[5, 100] = float("inf")   # the row statistic becomes inf -> rsqrt 0 -> NaN from 0 * inf: NaN blocks through the quantizer
    x[9, 7] = float("nan")
    w = torch.randn(hidden, device=DEV, dtype=torch.bfloat16, generator=g)
    res = torch.randn(rows, hidden, device=DEV, dtype=torch.bfloat16, generator=g)
    dt = dtypes.STR_TO_ELEM_DTYPE[elem]
    y, mx, h = glue_ops.rmsnorm(x, w, 1e-6, residual=res, to_mx=dt)
    ref = MXTensor.to_mx(y, dt, 32)
    assert_bits_equal(bits_of(mx._scale_e8m0), bits_of(ref._scale_e8m0), "scales")
    assert_bits_equal(bits_of(mx._data), bits_of(ref._data), "codes")
    # codes only (no bf16 output requested): same bytes
    _, mx2, _ = glue_ops.rmsnorm(x, w, 1e-6, residual=res, to_mx=dt, want_y=False)
    assert torch.equal(mx2._data, mx._data) and torch.equal(mx2._scale_e8m0, mx._scale_e8m0)


@pytest.mark.parametrize("b,t,hq,hk,d,stacked", [(1, 2048, 32, 8, 128, False), (32, 1, 32, 8, 128, True), (2, 77, 4, 2, 64, True), (3, 5, 6, 6, 32, False)])
def test_rope_is_bit_identical_to_apply_rotary_pos_emb(b, t, hq, hk, d, stacked):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    import torchmx  # noqa: F401
    from torchmx_b200 import glue_ops
    g = torch.Generator(device=DEV).manual_seed(b + t)
    if stacked:  # q, k are column slices of one stacked projection output
        qkv = torch.randn(b, t, (hq + 2 * hk) * d, device=DEV, dtype=torch.bfloat16, generator=g)
        q2, k2, _ = qkv.split([hq * d, hk * d, hk * d], dim=-1)
    else:
        q2 = torch.randn(b, t, hq * d, device=DEV, dtype=torch.bfloat16, generator=g)
        k2 = torch.randn(b, t, hk * d, device=DEV, dtype=torch.bfloat16, generator=g)
    q, k = q2.view(b, t, hq, d).transpose(1, 2), k2.view(b, t, hk, d).transpose(1, 2)
    ang = torch.rand(b, t, d // 2, device=DEV, generator=g) * 100
    emb = torch.cat([ang, ang], -1)
    cos, sin = emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)
    for cs in ((cos, sin), (cos[:1].expand(1, t, d), sin[:1].expand(1, t, d))):
        got = glue_ops.rope(q, k, *cs)
        assert got is not None
        want = apply_rotary_pos_emb(q, k, *cs)
        assert got[0].is_contiguous() and got[1].is_contiguous()
        assert_bits_equal(bits_of(got[0]), bits_of(want[0]), "q")
        assert_bits_equal(bits_of(got[1]), bits_of(want[1]), "k")


@pytest.mark.parametrize("b,t,hq,hk,d,pos,cache_len", [(32, 1, 32, 8, 128, 77, 160), (1, 300, 8, 2, 128, 0, 300), (2, 5, 4, 4, 64, 9, 40)])
def test_rope_writes_keys_and_values_straight_into_a_cache(b, t, hq, hk, d, pos, cache_len):
    """K5b with `k_out` / `v` / `v_out`: the rotated keys land in place in a cache slice and the value heads are copied beside
    them by the same launch == rope, then `index_copy_` of k and v (the cache update of the reference's call site,
    torchmx/layers/mx_llama_attention.py:189-193); everything else in the caches is left alone"""
    import torchmx  # noqa: F401
    from torchmx_b200 import glue_ops
    g = torch.Generator(device=DEV).manual_seed(b * t + pos)
    qkv = torch.randn(b, t, (hq + 2 * hk) * d, device=DEV, dtype=torch.bfloat16, generator=g)
    q2, k2, v2 = qkv.split([hq * d, hk * d, hk * d], dim=-1)
    q, k, v = (x.view(b, t, -1, d).transpose(1, 2) for x in (q2, k2, v2))
    ang = torch.rand(b, t, d // 2, device=DEV, generator=g) * 100
    emb = torch.cat([ang, ang], -1)
    cos, sin = emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)
    kc = torch.randn(b, hk, cache_len, d, device=DEV, dtype=torch.bfloat16, generator=g)
    vc = torch.randn(b, hk, cache_len, d, device=DEV, dtype=torch.bfloat16, generator=g)
    kc_want, vc_want = kc.clone(), vc.clone()
    q_want, k_want = glue_ops.rope(q, k, cos, sin)
    idx = torch.arange(pos, pos + t, device=DEV)
    kc_want.index_copy_(2, idx, k_want)
    vc_want.index_copy_(2, idx, v)
    got = glue_ops.rope(q, k, cos, sin, k_out=kc[:, :, pos:pos + t], v=v, v_out=vc[:, :, pos:pos + t])
    assert got is not None and got[1].data_ptr() == kc[:, :, pos:pos + t].data_ptr()
    assert_bits_equal(bits_of(got[0]), bits_of(q_want), "q")
    assert_bits_equal(bits_of(kc), bits_of(kc_want), "key cache")
    assert_bits_equal(bits_of(vc), bits_of(vc_want), "value cache")
    assert glue_ops.rope(q, k, cos, sin, v=v) is None  # v without v_out


@pytest.mark.parametrize("elem", ELEMS + ["float8_e5m2"])
@pytest.mark.parametrize("shape", [(1, 32, 2048, 128), (32, 8, 1, 128), (2, 3, 77, 64), (1, 2, 5, 32)])
def test_quantize_heads_is_k1_of_the_transposed_tensor(elem, shape):
    import torchmx  # noqa: F401
    from torchmx_b200 import dtypes, glue_ops
    from torchmx_b200.mx_tensor import MXTensor
    b, h, t, d = shape
    g = torch.Generator(device=DEV).manual_seed(t)
    x = torch.randn(*shape, device=DEV, dtype=torch.bfloat16, generator=g)
    x *= torch.exp2(torch.randint(-20, 20, (b, h, t, d // 32), device=DEV, generator=g).float()).repeat_interleave(32, -1).to(torch.bfloat16)
    x[0, 0, 0, 3] = float("nan")
    x[-1, -1, -1, -1] = float("inf")
    dt = dtypes.STR_TO_ELEM_DTYPE[elem]
    got = glue_ops.quantize_heads(x, dt)
    want = MXTensor.to_mx(x.transpose(1, 2).reshape(b, t, h * d).contiguous(), dt, 32)
    assert got.shape == want.shape
    assert_bits_equal(bits_of(got._scale_e8m0), bits_of(want._scale_e8m0), "scales")
    assert_bits_equal(bits_of(got._data), bits_of(want._data), "codes")


@pytest.mark.parametrize("elem", ELEMS + ["float8_e5m2"])
@pytest.mark.parametrize("shape,cache_len", [((32, 8, 256, 128), 0), ((1, 8, 2048, 128), 0), ((2, 3, 96, 64), 160), ((1, 2, 32, 40), 0), ((3, 1, 416, 72), 512)])
@pytest.mark.parametrize("mode", ["False", "True"])
def test_quantize_transposed_is_k1_of_the_transposed_tensor(elem, shape, cache_len, mode):
    """K5d: V quantized along the sequence axis straight from [b, h, kv, d] (also from a slice of a longer cache) == the strided
    copy + K1 of the reference's recipe (torchmx/layers/mx_llama_attention.py:205-212), bit for bit"""
    import torchmx  # noqa: F401
    from torchmx_b200 import dtypes, glue_ops
    from torchmx_b200 import env_variables as env
    from torchmx_b200.mx_tensor import MXTensor
    env.MX_EXACT_QUANTIZATION = mode
    b, h, kv, d = shape
    g = torch.Generator(device=DEV).manual_seed(kv + d)
    full = torch.randn(b, h, max(cache_len, kv), d, device=DEV, dtype=torch.bfloat16, generator=g)
    full *= torch.exp2(torch.randint(-20, 20, (b, h, full.shape[2] // 32, 1, d), device=DEV, generator=g).float()).expand(-1, -1, -1, 32, -1).reshape(full.shape).to(torch.bfloat16)
    x = full[:, :, :kv]  # (a view with the cache's strides when cache_len > kv)
    x[0, 0, 3, 1] = float("nan")
    x[-1, -1, -1, -1] = float("inf")
    x[0, -1, :32, 2] = 0
    dt = dtypes.STR_TO_ELEM_DTYPE[elem]
    n0 = glue_ops.stats["quantize_transposed"]
    got = glue_ops.quantize_transposed(x, dt)
    assert got is not None and glue_ops.stats["quantize_transposed"] == n0 + 1
    want = MXTensor.to_mx(x.transpose(2, 3).contiguous(), dt, 32)
    assert got.shape == want.shape == (b, h, d, kv)
    assert_bits_equal(bits_of(got._scale_e8m0), bits_of(want._scale_e8m0), "scales")
    assert_bits_equal(bits_of(got._data), bits_of(want._data), "codes")


def test_glue_kernels_decline_what_they_cannot_take():
    import torchmx  # noqa: F401
    from torchmx_b200 import glue_ops
    x = torch.randn(4, 48, device=DEV, dtype=torch.bfloat16)
    assert glue_ops.rmsnorm(x, torch.ones(48, device=DEV, dtype=torch.bfloat16), 1e-5) is None      # hidden % 32
    x = torch.randn(4, 64, device=DEV, dtype=torch.float32)
    assert glue_ops.rmsnorm(x, torch.ones(64, device=DEV, dtype=torch.float32), 1e-5) is None      # bf16 only
    q = torch.randn(1, 2, 4, 24, device=DEV, dtype=torch.bfloat16)
    cs = torch.randn(1, 4, 24, device=DEV, dtype=torch.bfloat16)
    assert glue_ops.rope(q, q, cs, cs) is None                                                      # head_dim % 16
    from torchmx_b200 import dtypes
    assert glue_ops.quantize_transposed(torch.randn(1, 2, 48, 64, device=DEV, dtype=torch.bfloat16), dtypes.float8_e4m3) is None   # rows % 32
    assert glue_ops.quantize_transposed(torch.randn(1, 2, 64, 64, device=DEV, dtype=torch.bfloat16).transpose(2, 3), dtypes.float8_e4m3) is None  # cols strided


def test_quantize_llm_with_fused_norms_matches_the_unfused_model():
    """`quantize_llm_(..., fuse_rmsnorm=True)`: the norms of a decoder layer hand MXTensors to the MX blocks (prefill and decode sizes
    alike); logits agree with the unfused quantized model to the last-bit effects of the norm statistic"""
    import copy
    from transformers import LlamaConfig, LlamaForCausalLM
    import torchmx  # noqa: F401
    from torchmx_b200 import glue_ops
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.quant_api import FusedRMSNorm, quantize_llm_
    cfg = LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                      max_position_embeddings=512)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).to(DEV, torch.bfloat16).eval()
    lin = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    a, b = copy.deepcopy(model), copy.deepcopy(model)
    quantize_llm_(a, QAttentionConfig(projection_config=lin), lin)
    quantize_llm_(b, QAttentionConfig(projection_config=lin), lin, fuse_rmsnorm=True)
    l0 = b.model.layers[0]
    assert type(l0.input_layernorm) is FusedRMSNorm and l0.input_layernorm.to_mx is not None and l0.post_attention_layernorm.to_mx is not None
    assert type(b.model.norm) is FusedRMSNorm and b.model.norm.to_mx is None  # lm_head slices its input: stays bf16

    def sqnr(r, x):
        return float(20 * torch.log10(r.float().norm() / (r.float() - x.float()).norm()))

    for shape in ((2, 128), (4, 1)):  # prefill and decode: the decoder-layer norms hand MXTensors to the MX blocks
        ids = torch.randint(0, cfg.vocab_size, shape, device=DEV)
        before = dict(glue_ops.stats)
        with torch.no_grad():
            la, lb = a(input_ids=ids).logits, b(input_ids=ids).logits
        assert sqnr(la, lb) > 35, sqnr(la, lb)
        key = "rmsnorm_to_mx"
        assert glue_ops.stats[key] - before[key] >= 4 and glue_ops.stats["rope"] - before["rope"] == 4  # (rope: both models, two layers each)
