"""GPU tier: K4b (csrc/mxq_flash_attention.cu) -- the MX attention of the reference's blocks (torchmx/layers/mx_llama_attention.py:
195-243) as one kernel -- against the chain it replaces, built from the separately pinned kernels: K3 bmm (scores), K4a (softmax +
quantization of P, bit-exact against its own restated PyTorch chain in tests/test_gpu_attention.py) and K3 bmm again (P @ V).

The bar: the codes and scales of P are K4a's BIT FOR BIT (same rounding steps, same order of the fp32 row sum), and the output is
the bmm's bit for bit up to the sign of a zero (a chunk of keys no row of a tile can see is skipped instead of adding +-0)."""
import numpy as np
import pytest
import torch

from tests.util import assert_bits_equal, bits_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _repeat(t, n):
    """repeat_kv on the codes and scales, as contiguous copies (a single key / value head would otherwise come back as a stride-0
    view, which the bmm serves through the dequantize GEMM: same products, another accumulation order)"""
    from torchmx_b200.mx_tensor import MXTensor
    if n == 1:
        return t
    rep = lambda x: x.repeat_interleave(n, dim=1).contiguous()  # noqa: E731
    return MXTensor(rep(t._scale_e8m0), rep(t._data), t._elem_dtype, t._block_size, t._orig_dtype, t._padding, t._block_dim)


def _chain(q_mx, k_mx, vt_mx, scaling, mask, causal, p_dt):
    """the unfused path of layers/mx_llama_attention.py: bmm -> K4a -> bmm, output transposed to [b, q, h, d]"""
    from torchmx_b200 import attention_ops
    groups = q_mx.shape[1] // k_mx.shape[1]
    scores = torch.matmul(q_mx, _repeat(k_mx, groups).transpose(2, 3))
    p_mx = attention_ops.softmax_to_mx(scores, scaling, mask, causal, p_dt, 32)
    assert p_mx is not None
    out = torch.matmul(p_mx, _repeat(vt_mx, groups).transpose(2, 3))
    return out.transpose(1, 2).contiguous(), p_mx


def _inputs(b, h, hk, q_len, kv_len, seed, spread=2.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    q = torch.randn(b, h, q_len, 128, device=DEV, dtype=torch.bfloat16, generator=g) * spread
    k = torch.randn(b, hk, kv_len, 128, device=DEV, dtype=torch.bfloat16, generator=g) * spread
    v = torch.randn(b, hk, kv_len, 128, device=DEV, dtype=torch.bfloat16, generator=g)
    return q, k, v


def _quantize(q, k, v, qd, kd, vd):
    from torchmx_b200 import dtypes
    from torchmx_b200.mx_tensor import MXTensor
    E = dtypes.STR_TO_ELEM_DTYPE
    return (MXTensor.to_mx(q.contiguous(), E[qd], 32), MXTensor.to_mx(k.contiguous(), E[kd], 32),
            MXTensor.to_mx(v.transpose(2, 3).contiguous(), E[vd], 32))


def _same_up_to_zero_sign(got, want, what):
    g, w = bits_of(got).astype(np.int64), bits_of(want).astype(np.int64)
    zero = ((g & 0x7FFF) == 0) & ((w & 0x7FFF) == 0)
    assert_bits_equal(np.where(zero, 0, g), np.where(zero, 0, w), what)


def _additive_causal(q_len, kv_len, pad_from=None):
    m = torch.full((1, 1, q_len, kv_len), torch.finfo(torch.bfloat16).min, device=DEV, dtype=torch.bfloat16).triu_(kv_len - q_len + 1)
    if pad_from is not None:
        m[..., pad_from:] = torch.finfo(torch.bfloat16).min
    return m


CASES = [
    # b, h, hk, q_len, kv_len, masking
    (1, 2, 2, 128, 128, "causal"),      # one chunk: rows of 4 blocks (sequential sum order)
    (1, 4, 1, 256, 256, "causal"),      # both warpgroups, grouped-query heads, butterfly sum order
    (2, 4, 2, 512, 512, "causal"),
    (1, 2, 2, 2048, 2048, "causal"),    # the prefill size of BASELINE configs[3]
    (1, 2, 1, 200, 256, "causal"),      # ragged query tile, kv_len > q_len (cached prefix)
    (1, 2, 2, 1, 384, "none"),          # decode-like: one query row, unmasked
    (1, 2, 2, 130, 1024, "none"),       # unmasked, 32 blocks per row
    (2, 2, 2, 384, 384, "mask"),        # explicit additive mask (finfo.min, the eager / sdpa mask interface), odd number of chunks
    (1, 2, 2, 256, 640, "mask+pad"),    # + padded keys at the end
    (1, 1, 1, 512, 512, "mask+causal"),
    # kv_len a multiple of 32 but not of 128: the last chunk holds 1..3 blocks
    (4, 4, 2, 1, 160, "mask"),          # a decode step against a 160-slot cache, additive mask (grouped-query heads as tile rows)
    (3, 8, 2, 1, 256, "none"),          # a decode step, four query heads per key / value head, no mask
    (2, 4, 2, 200, 224, "causal"),
    (1, 2, 1, 96, 96, "causal"),        # less than one chunk
    (1, 2, 2, 300, 416, "mask+pad"),
    (1, 2, 2, 5, 352, "none"),
    (1, 2, 2, 64, 1000 // 32 * 32, "causal"),
    # at most 128 query rows per (batch, head): the one-tile variant of the kernel (two CTAs per SM)
    (2, 4, 2, 100, 640, "mask+pad"),
    (1, 2, 2, 128, 512, "mask+causal"),
    (2, 8, 2, 1, 1024, "mask"),         # decode against a long cache
    (32, 32, 8, 1, 256, "mask"),        # the decode step of BASELINE configs[3]: batch 32, four query heads per key / value head
]


@pytest.mark.parametrize("b,h,hk,q_len,kv_len,masking", CASES)
def test_flash_attention_matches_the_bmm_softmax_bmm_chain_bit_for_bit(b, h, hk, q_len, kv_len, masking):
    import torchmx  # noqa: F401
    from torchmx_b200 import attention_ops, dtypes
    q, k, v = _inputs(b, h, hk, q_len, kv_len, seed=q_len + kv_len)
    q_mx, k_mx, vt_mx = _quantize(q, k, v, "float8_e4m3", "float8_e4m3", "float8_e4m3")
    mask = None
    if masking.startswith("mask"):
        mask = _additive_causal(q_len, kv_len, pad_from=kv_len - 70 if "pad" in masking else None)
    causal = "causal" in masking
    scaling = 128 ** -0.5
    before = attention_ops.stats["flash_attention"]
    res = attention_ops.flash_attention(q_mx, k_mx, vt_mx, scaling, mask, causal, dtypes.float8_e4m3, 32, return_probs=True)
    assert res is not None and attention_ops.stats["flash_attention"] == before + 1
    out, probs = res
    want, p_ref = _chain(q_mx, k_mx, vt_mx, scaling, mask, causal, dtypes.float8_e4m3)
    assert_bits_equal(bits_of(probs._scale_e8m0), bits_of(p_ref._scale_e8m0), "scales of P")
    assert_bits_equal(bits_of(probs._data), bits_of(p_ref._data), "codes of P")
    assert out.shape == (b, q_len, h, 128) and out.is_contiguous()
    if kv_len % 128 == 0:
        _same_up_to_zero_sign(out, want, "attention output")
    else:
        # a key axis that is not a multiple of 128 sends the chain's P @ V through the dequantize-GEMM (bf16 MMAs, K = 16 per
        # step): the same exact products in another accumulation order -- the GEMM tolerance instead of equality
        pd, vd = p_ref.to_dtype(torch.float32).double(), _repeat(vt_mx, h // hk).to_dtype(torch.float32).double()
        ref, S = (pd @ vd.transpose(2, 3)).transpose(1, 2), (pd.abs() @ vd.abs().transpose(2, 3)).transpose(1, 2)
        for o in (out, want):
            assert int(((o.double() - ref).abs() > 2.0 ** -8 * ref.abs() + 2.0 ** -18 * S + 1e-30).sum()) == 0
    # and without the dump of P: same output
    out2 = attention_ops.flash_attention(q_mx, k_mx, vt_mx, scaling, mask, causal, dtypes.float8_e4m3, 32)
    assert torch.equal(out2.view(torch.int16), out.view(torch.int16))


@pytest.mark.parametrize("qd,kd,vd,pd", [("float6_e3m2", "float6_e2m3", "float4_e2m1", "float6_e3m2"), ("float4_e2m1", "float4_e2m1", "float6_e3m2", "float4_e2m1"),
                                        ("float8_e4m3", "float6_e3m2", "float8_e4m3", "float6_e2m3"), ("float8_e5m2", "float8_e4m3", "float8_e5m2", "float8_e5m2")])
@pytest.mark.parametrize("mode", ["False", "True"])
def test_flash_attention_every_element_type(qd, kd, vd, pd, mode):
    import torchmx  # noqa: F401
    from torchmx_b200 import attention_ops, dtypes
    from torchmx_b200 import env_variables as env
    env.MX_EXACT_QUANTIZATION = mode
    q, k, v = _inputs(1, 4, 2, 384, 384, seed=7)
    q_mx, k_mx, vt_mx = _quantize(q, k, v, qd, kd, vd)
    p_dt = dtypes.STR_TO_ELEM_DTYPE[pd]
    out, probs = attention_ops.flash_attention(q_mx, k_mx, vt_mx, 0.09, None, True, p_dt, 32, return_probs=True)
    want, p_ref = _chain(q_mx, k_mx, vt_mx, 0.09, None, True, p_dt)
    assert_bits_equal(bits_of(probs._scale_e8m0), bits_of(p_ref._scale_e8m0), "scales of P")
    assert_bits_equal(bits_of(probs._data), bits_of(p_ref._data), "codes of P")
    _same_up_to_zero_sign(out, want, "attention output")


def test_flash_attention_nan_and_extreme_rows():
    """a NaN query element poisons its row (scale 255 for every block of the row, NaN output), huge scores saturate one
    probability to 1 and push the rest below every representable value: same bytes as the chain"""
    import torchmx  # noqa: F401
    from torchmx_b200 import attention_ops, dtypes
    q, k, v = _inputs(1, 2, 2, 256, 256, seed=3)
    q[0, 0, 5, 17] = float("nan")
    q[0, 1, 100] *= 64
    k[0, 1, 40] *= 64
    q_mx, k_mx, vt_mx = _quantize(q, k, v, "float8_e4m3", "float8_e4m3", "float8_e4m3")
    out, probs = attention_ops.flash_attention(q_mx, k_mx, vt_mx, 0.09, None, True, dtypes.float8_e4m3, 32, return_probs=True)
    want, p_ref = _chain(q_mx, k_mx, vt_mx, 0.09, None, True, dtypes.float8_e4m3)
    assert_bits_equal(bits_of(probs._scale_e8m0), bits_of(p_ref._scale_e8m0), "scales of P")
    assert_bits_equal(bits_of(probs._data), bits_of(p_ref._data), "codes of P")
    # the whole block row of Q that holds the NaN has scale 255 -> NaN scores -> NaN output for that query row only
    assert bool(torch.isnan(out[0, 5, 0]).all()) and bool(torch.isnan(want[0, 5, 0]).all())
    _same_up_to_zero_sign(out, want, "attention output")


def test_flash_attention_close_to_fp32_attention_of_the_dequantized_operands():
    """independent of K3 / K4a: fp32 attention of the dequantized Q, K, V.  P is quantized to e4m3 per 32 keys (3 mantissa bits),
    so the bound is the quantization noise of P, not an ulp"""
    import torchmx  # noqa: F401
    from torchmx_b200 import attention_ops, dtypes
    q, k, v = _inputs(1, 4, 2, 512, 512, seed=11, spread=1.0)
    q_mx, k_mx, vt_mx = _quantize(q, k, v, "float8_e4m3", "float8_e4m3", "float8_e4m3")
    out = attention_ops.flash_attention(q_mx, k_mx, vt_mx, 128 ** -0.5, None, True, dtypes.float8_e4m3, 32)
    qf, kf, vf = q_mx.to_dtype(torch.float32), k_mx.to_dtype(torch.float32), vt_mx.to_dtype(torch.float32).transpose(2, 3)
    kf, vf = kf.repeat_interleave(2, 1), vf.repeat_interleave(2, 1)
    s = (qf @ kf.transpose(2, 3)) * 128 ** -0.5
    s = s.masked_fill(torch.ones(512, 512, dtype=torch.bool, device=DEV).triu_(1), float("-inf"))
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2)
    sqnr = float(20 * torch.log10(ref.norm() / (ref - out.float()).norm()))
    assert sqnr > 25, sqnr


def test_flash_attention_declines_what_it_cannot_take():
    import torchmx  # noqa: F401
    from torchmx_b200 import attention_ops, dtypes
    g = torch.Generator(device=DEV).manual_seed(1)
    q64, k64, v64 = (torch.randn(1, 2, 128, 64, device=DEV, dtype=torch.bfloat16, generator=g) for _ in range(3))
    assert attention_ops.flash_attention(*_quantize(q64, k64, v64, "float8_e4m3", "float8_e4m3", "float8_e4m3"), 0.1, None, True, dtypes.float8_e4m3, 32) is None  # head_dim 64
    q, k, v = _inputs(1, 2, 2, 128, 128, seed=1)
    q_mx, k_mx, vt_mx = _quantize(q, k, v, "int8", "float8_e4m3", "float8_e4m3")
    assert attention_ops.flash_attention(q_mx, k_mx, vt_mx, 0.1, None, True, dtypes.float8_e4m3, 32) is None   # int8 has no MMA form
    q_mx, k_mx, vt_mx = _quantize(q, k, v, "float8_e4m3", "float8_e4m3", "float8_e4m3")
    assert attention_ops.flash_attention(q_mx, k_mx, vt_mx, 0.1, None, True, dtypes.int8, 32) is None
    prev = attention_ops.set_flash_attention(False)
    try:
        assert attention_ops.flash_attention(q_mx, k_mx, vt_mx, 0.1, None, True, dtypes.float8_e4m3, 32) is None
    finally:
        attention_ops.set_flash_attention(prev)


def test_mx_attention_block_uses_the_flash_kernel_and_matches_the_chain():
    """`quantize_llm_` with Q / K / V / attention-weights configs on a Llama with head_dim 128: the block's attention is one K4b
    launch per layer and the logits equal the K3 -> K4a -> K3 chain's (up to the sign of zeros)"""
    import copy
    from transformers import LlamaConfig, LlamaForCausalLM
    import torchmx  # noqa: F401
    from torchmx_b200 import attention_ops
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.quant_api import quantize_llm_
    cfg = LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                      max_position_embeddings=512)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).to(DEV, torch.bfloat16).eval()
    lin = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    e = MXConfig("float8_e4m3", 32)
    quantize_llm_(model, QAttentionConfig(projection_config=lin, query_config=e, key_config=e, value_config=e, attention_weights_config=e), lin)
    ids = torch.randint(0, cfg.vocab_size, (2, 256), device=DEV)
    before = dict(attention_ops.stats)
    with torch.no_grad():
        a = model(input_ids=ids).logits
    assert attention_ops.stats["flash_attention"] - before["flash_attention"] == 2 and attention_ops.stats["fused_softmax"] == before["fused_softmax"]
    prev = attention_ops.set_flash_attention(False)
    try:
        with torch.no_grad():
            c = model(input_ids=ids).logits
    finally:
        attention_ops.set_flash_attention(prev)
    assert attention_ops.stats["fused_softmax"] - before["fused_softmax"] == 2
    _same_up_to_zero_sign(a, c, "logits")
