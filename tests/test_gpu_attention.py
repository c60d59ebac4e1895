"""GPU tier: K4a `mxq_softmax_quantize` -- scale + mask + fp32 softmax + bf16 rounding + MX quantization of the attention
probabilities in one pass (reference chain: torchmx/layers/mx_llama_attention.py:214-239).

Parity is checked two ways:
* bit-exact against a PyTorch restatement of the chain that adds the fp32 row sum in the kernel's order (everything else in
  the chain is order-independent, so codes and scales must be identical);
* against the chain exactly as the reference spells it (torch.softmax picks its own summation order): the probabilities may
  differ in the last ulp of the denominator, so a handful of bf16 values round the other way -- bounded here.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ordered_row_sum(e: torch.Tensor, masked: bool) -> torch.Tensor:
    """fp32 sum over the last dim in the order K4a uses: 32 sequential adds per MX block, then over the row's blocks
    * fewer than 8 blocks, or at most 32 without any mask: sequentially;
    * 8 .. 256 blocks under a mask (causal or additive): a 3-step butterfly over each group of 8 consecutive blocks (one row's
      lanes of a warp), then sequentially over the groups;
    * otherwise: a 5-step butterfly over each group of 32 blocks (a warp), then sequentially over the groups."""
    kv = e.shape[-1]
    tpr = kv // 32
    eb = e.reshape(*e.shape[:-1], tpr, 32)
    s = torch.zeros_like(eb[..., 0])
    for i in range(32):
        s = s + eb[..., i]
    if tpr < 8 or (not masked and tpr <= 32):
        acc = torch.zeros_like(s[..., 0])
        for j in range(tpr):
            acc = acc + s[..., j]
        return acc
    width = 8 if (masked and tpr <= 256) else 32
    groups = (tpr + width - 1) // width
    v = torch.nn.functional.pad(s, (0, groups * width - tpr)).reshape(*s.shape[:-1], groups, width)
    d = width // 2
    while d >= 1:
        v = v[..., :d] + v[..., d:2 * d]
        d //= 2
    v = v[..., 0]
    acc = v[..., 0]
    for j in range(1, groups):
        acc = acc + v[..., j]
    return acc


def _chain(scores, scaling, mask, causal, ordered: bool):
    """the reference chain up to the bf16 probabilities"""
    q_len, kv_len = scores.shape[-2:]
    w = scores * scaling
    if mask is not None:
        w = w + mask
    if causal:
        w = w.masked_fill(torch.ones(q_len, kv_len, dtype=torch.bool, device=w.device).triu_(kv_len - q_len + 1), float("-inf"))
    if not ordered:
        return torch.softmax(w, dim=-1, dtype=torch.float32).to(torch.bfloat16)
    x = w.float()
    e = torch.exp(x - x.amax(-1, keepdim=True))
    return (e / _ordered_row_sum(e, causal or mask is not None).unsqueeze(-1)).to(torch.bfloat16)


def _make(b, h, q, kv, mode, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    scores = (torch.randn(b, h, q, kv, device=DEV, generator=g) * 24).to(torch.bfloat16)
    mask, causal = None, False
    if mode == "causal":
        causal = True
    elif mode in ("mask", "mask_bcast", "mask_sliced"):
        mb, mh = (1, 1) if mode == "mask_bcast" else (b, h)
        width = kv + 5 if mode == "mask_sliced" else kv  # a slice of a wider mask: rows lose their 16-byte alignment
        m = torch.zeros(mb, mh, q, width, device=DEV, dtype=torch.bfloat16)
        m.masked_fill_(torch.rand(mb, mh, q, width, device=DEV, generator=g) < 0.3, torch.finfo(torch.bfloat16).min)
        m[..., 0] = 0  # no fully hidden row
        mask = m[..., :kv]
    return scores, mask, causal


ELEMS = ["float8_e4m3", "float6_e3m2", "float6_e2m3", "float4_e2m1", "int8"]


@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("shape,mode", [
    ((2, 3, 64, 32), "none"), ((1, 2, 40, 96), "causal"), ((2, 2, 17, 160), "mask"), ((1, 4, 128, 1024), "causal"),
    ((1, 2, 64, 1056), "mask_bcast"), ((2, 2, 256, 2048), "causal"), ((1, 1, 8, 4096), "mask_sliced"), ((1, 1, 3, 32768), "none"),
    ((1, 2, 2048, 2048), "mask_bcast"), ((1, 3, 70, 2176), "causal"), ((1, 1, 9, 8192), "mask"), ((1, 1, 5, 8224), "causal"),
    ((2, 1, 33, 224), "causal"), ((1, 1, 2, 256), "none"), ((1, 2, 50, 1024), "none"), ((1, 1, 20, 2080), "none"), ((1, 1, 7, 512), "mask"),
])
def test_softmax_to_mx_bit_exact_against_the_ordered_chain(elem, shape, mode):
    import torchmx  # noqa: F401
    from torchmx import attention_ops, dtypes
    from torchmx.mx_tensor import MXTensor
    et = dtypes.STR_TO_ELEM_DTYPE[elem]
    scores, mask, causal = _make(*shape, mode)
    scaling = 128 ** -0.5
    n0 = attention_ops.stats["fused_softmax"]
    got = attention_ops.softmax_to_mx(scores, scaling, mask, causal, et, 32)
    assert got is not None and attention_ops.stats["fused_softmax"] == n0 + 1
    want = MXTensor.to_mx(_chain(scores, scaling, mask, causal, ordered=True), et, 32)
    assert got.shape == want.shape and got.dtype == want.dtype and got._data.dtype == want._data.dtype
    assert torch.equal(got._scale_e8m0, want._scale_e8m0)
    assert torch.equal(got._data, want._data)


@pytest.mark.parametrize("elem", ["float8_e4m3", "float4_e2m1"])
def test_softmax_to_mx_against_torch_softmax(elem):
    """torch.softmax sums in its own order: identical up to the rare bf16 rounding flip"""
    import torchmx  # noqa: F401
    from torchmx import attention_ops, dtypes
    from torchmx.mx_tensor import MXTensor
    et = dtypes.STR_TO_ELEM_DTYPE[elem]
    scores, mask, causal = _make(1, 8, 1024, 2048, "causal", seed=3)
    scaling = 128 ** -0.5
    got = attention_ops.softmax_to_mx(scores, scaling, mask, causal, et, 32)
    p_ref = _chain(scores, scaling, mask, causal, ordered=False)
    want = MXTensor.to_mx(p_ref, et, 32)
    differing = (got._data != want._data).float().mean().item()
    assert differing < 2e-3, differing
    d = (got.to_dtype(torch.float32) - want.to_dtype(torch.float32)).abs()
    step = 0.34 if elem == "float4_e2m1" else 0.125  # never more than one code step of the largest block
    assert d.max().item() <= step * p_ref.float().max().item()
    assert (got._scale_e8m0 != want._scale_e8m0).float().mean().item() < 1e-3


def test_softmax_to_mx_nan_and_hidden_rows():
    """a NaN score poisons its row (scale 255, codes zero) and a fully hidden row is NaN too, exactly like the unfused chain"""
    import torchmx  # noqa: F401
    from torchmx import attention_ops, dtypes
    from torchmx.mx_tensor import MXTensor
    scores, _, _ = _make(1, 2, 16, 256, "none", seed=5)
    scores[0, 0, 3, 77] = float("nan")
    mask = torch.zeros(1, 1, 16, 256, device=DEV, dtype=torch.bfloat16)
    mask[0, 0, 9, :] = float("-inf")
    for hw_exact in ("False", "True"):
        from torchmx import env_variables as env
        prev, env.MX_EXACT_QUANTIZATION = env.MX_EXACT_QUANTIZATION, hw_exact
        try:
            got = attention_ops.softmax_to_mx(scores, 0.125, mask, False, dtypes.float8_e4m3, 32)
            want = MXTensor.to_mx(_chain(scores, 0.125, mask, False, ordered=True), dtypes.float8_e4m3, 32)
        finally:
            env.MX_EXACT_QUANTIZATION = prev
        assert torch.equal(got._scale_e8m0, want._scale_e8m0) and torch.equal(got._data, want._data)
        assert (got._scale_e8m0[0, 0, 3] == 255).all() and (got._scale_e8m0[0, :, 9] == 255).all()
        assert (got._scale_e8m0[0, 1, 3] != 255).all()


def test_softmax_to_mx_declines_what_it_cannot_do():
    import torchmx  # noqa: F401
    from torchmx import attention_ops, dtypes
    s = torch.randn(1, 1, 4, 48, device=DEV).to(torch.bfloat16)
    assert attention_ops.softmax_to_mx(s, 1.0, None, False, dtypes.float8_e4m3, 32) is None          # kv % 32
    s = torch.randn(1, 1, 4, 64, device=DEV).to(torch.bfloat16)
    assert attention_ops.softmax_to_mx(s, 1.0, None, False, dtypes.float8_e4m3, 16) is None          # block size
    assert attention_ops.softmax_to_mx(s.float(), 1.0, None, False, dtypes.float8_e4m3, 32) is None  # dtype
    assert attention_ops.softmax_to_mx(s.transpose(1, 2), 1.0, None, False, dtypes.float8_e4m3, 32) is not None  # [1,4,1,64] is contiguous
    assert attention_ops.softmax_to_mx(s, 1.0, torch.zeros(1, 1, 4, 64, device=DEV), False, dtypes.float8_e4m3, 32) is None  # fp32 mask


def test_mx_attention_block_fused_softmax_and_implied_causal_mask(monkeypatch):
    """the MX attention block with the fused kernel agrees with the unfused chain, and a model configured for sdpa (no mask
    tensor: the attention function is expected to apply is_causal) attends causally just like the eager configuration"""
    import copy
    from transformers import LlamaConfig, LlamaForCausalLM
    import torchmx  # noqa: F401
    from torchmx import attention_ops
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.quant_api import quantize_llm_
    cfg = LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                      max_position_embeddings=512)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).to(DEV, torch.bfloat16).eval()
    lin = QLinearConfig(weights_config=MXConfig("float8_e4m3", 32), activations_config=MXConfig("float8_e4m3", 32))
    e = MXConfig("float8_e4m3", 32)
    qm = copy.deepcopy(model)
    quantize_llm_(qm, QAttentionConfig(projection_config=lin, query_config=e, key_config=e, value_config=e, attention_weights_config=e), lin)
    ids = torch.randint(0, cfg.vocab_size, (2, 128), device=DEV)
    monkeypatch.setattr(attention_ops, "_FLASH", False)  # (this test is about K4a inside the bmm -> softmax -> bmm chain; K4b: test_gpu_flash_attention.py)
    n0 = dict(attention_ops.stats)
    with torch.no_grad():
        fused = qm(input_ids=ids).logits
    assert attention_ops.stats["fused_softmax"] == n0["fused_softmax"] + 2 and attention_ops.stats["unfused_softmax"] == n0["unfused_softmax"]
    prev = attention_ops.set_fused_softmax(False)
    try:
        with torch.no_grad():
            unfused = qm(input_ids=ids).logits
    finally:
        attention_ops.set_fused_softmax(prev)
    assert attention_ops.stats["unfused_softmax"] == n0["unfused_softmax"] + 2
    sqnr = float(20 * torch.log10(unfused.float().norm() / (unfused.float() - fused.float()).norm()))
    assert sqnr > 40, sqnr
    # sdpa configuration: HF skips the mask tensor for a plain causal prefill
    qm.config._attn_implementation = "sdpa"
    for layer in qm.model.layers:
        layer.self_attn.config._attn_implementation = "sdpa"
    with torch.no_grad():
        implied = qm(input_ids=ids).logits
    sqnr = float(20 * torch.log10(fused.float().norm() / (fused.float() - implied.float()).norm()))
    assert sqnr > 40, sqnr
    # and the last token must not see the future: changing later tokens leaves earlier logits untouched
    ids2 = ids.clone()
    # (V is quantized along the sequence, so a 32-token block is the granularity at which nothing may leak)
    ids2[:, 96:] = (ids2[:, 96:] + 1) % cfg.vocab_size
    with torch.no_grad():
        implied2 = qm(input_ids=ids2).logits
    assert torch.equal(implied[:, :96], implied2[:, :96])
    assert not torch.equal(implied[:, 96:], implied2[:, 96:])


@pytest.mark.parametrize("elem", ["float8_e4m3", "float6_e3m2", "float4_e2m1"])
def test_quantizing_kv_heads_before_repeating_them_is_bit_identical(elem):
    """the reference repeats K / V to the query heads and then quantizes (mx_llama_attention.py:189-213); the block quantizes
    the key/value heads once and repeats codes + scales"""
    from transformers.models.llama.modeling_llama import repeat_kv
    import torchmx  # noqa: F401
    from torchmx import dtypes
    from torchmx.layers.mx_llama_attention import _repeat_heads
    from torchmx.mx_tensor import MXTensor
    et = dtypes.STR_TO_ELEM_DTYPE[elem]
    torch.manual_seed(4)
    k = torch.randn(2, 2, 160, 128, device=DEV, dtype=torch.bfloat16)
    v = torch.randn(2, 2, 160, 128, device=DEV, dtype=torch.bfloat16)
    want_k = MXTensor.to_mx(repeat_kv(k, 4).contiguous(), et, 32)
    got_k = _repeat_heads(MXTensor.to_mx(k.contiguous(), et, 32), 4)
    want_v = MXTensor.to_mx(repeat_kv(v, 4).transpose(2, 3).contiguous(), et, 32).transpose(2, 3)
    got_v = _repeat_heads(MXTensor.to_mx(v.transpose(2, 3).contiguous(), et, 32), 4).transpose(2, 3)
    for got, want in ((got_k, want_k), (got_v, want_v)):
        assert got.shape == want.shape and got._block_dim == want._block_dim
        assert torch.equal(got._data, want._data) and torch.equal(got._scale_e8m0, want._scale_e8m0)
    q = MXTensor.to_mx(torch.randn(2, 8, 128, 128, device=DEV, dtype=torch.bfloat16), dtypes.float8_e4m3, 32)
    assert torch.equal(torch.matmul(q, got_k.transpose(2, 3)), torch.matmul(q, want_k.transpose(2, 3)))


# ---- K1b: SwiGLU gating + quantization (mxq_silu_mul_quantize) ---------------------------------------------------------------
@pytest.mark.parametrize("elem", ELEMS)
@pytest.mark.parametrize("shape", [(1, 64), (3, 7, 96), (300, 1024), (2, 128, 14336)])
def test_silu_mul_to_mx_is_bit_identical_to_the_chain(elem, shape):
    import torchmx  # noqa: F401
    from torchmx import dtypes, mlp_ops
    from torchmx.mx_tensor import MXTensor
    et = dtypes.STR_TO_ELEM_DTYPE[elem]
    g = torch.Generator(device=DEV).manual_seed(11)
    gate = (torch.randn(*shape, device=DEV, generator=g) * 4).to(torch.bfloat16)
    up = (torch.randn(*shape, device=DEV, generator=g) * 2).to(torch.bfloat16)
    gate.view(-1)[:8] = torch.tensor([0.0, -0.0, 88.0, -88.0, -120.0, 3e38, -3e38, 1e-30], device=DEV).to(torch.bfloat16)
    got = mlp_ops.silu_mul_to_mx(gate, up, et, 32)
    want = MXTensor.to_mx(torch.nn.functional.silu(gate) * up, et, 32)
    assert got is not None and got.shape == want.shape
    assert torch.equal(got._scale_e8m0, want._scale_e8m0) and torch.equal(got._data, want._data)


@pytest.mark.parametrize("elem", ["int8", "float8_e4m3"])
def test_silu_mul_to_mx_every_bf16_gate_value(elem):
    """K1b evaluates silu by a five-instruction sequence (csrc/mxq_silu.cuh); the gate has only 65536 possible values, all of them
    go through here -- each alone in its block beside zeros (so that it sets the block scale and keeps as many of its bits as the
    element type has) and among random neighbours, with up = 1, a power of two and random multipliers."""
    import torchmx  # noqa: F401
    from torchmx import dtypes, mlp_ops
    from torchmx.mx_tensor import MXTensor
    et = dtypes.STR_TO_ELEM_DTYPE[elem]
    every = torch.arange(65536, device=DEV, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    g = torch.Generator(device=DEV).manual_seed(13)
    alone = torch.zeros(65536, 32, device=DEV, dtype=torch.bfloat16)
    alone[:, 5] = every
    among = (torch.randn(65536, 32, device=DEV, generator=g) * 3).to(torch.bfloat16)
    among[:, 17] = every
    gate = torch.cat([alone, among], dim=1).reshape(512, 128 * 64).contiguous()
    for up in (torch.ones_like(gate), torch.full_like(gate, -0.125), (torch.randn(gate.shape, device=DEV, generator=g) * 2).to(torch.bfloat16)):
        got = mlp_ops.silu_mul_to_mx(gate, up, et, 32)
        want = MXTensor.to_mx(torch.nn.functional.silu(gate) * up, et, 32)
        assert torch.equal(got._scale_e8m0, want._scale_e8m0) and torch.equal(got._data, want._data)


def test_silu_sequence_equals_the_plain_formula_for_every_bf16_input(tmp_path):
    """tools/silu_check.cu compares the bf16 rounding of the shipped sequence with that of g / (1 + expf(-g)) for all 65536 inputs
    on this GPU and exits non-zero on any difference (recorded run: profiles/r2_k1b_silu_check.json)"""
    import json, os, shutil, subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "silu_check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-I", os.path.join(root, "torchmx_b200", "csrc"),
                    os.path.join(root, "tools", "silu_check.cu"), "-o", exe], check=True, timeout=300)
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    report = json.loads(run.stdout)
    assert report["cuda"] == "no error" and report["inputs"] == 65536
    assert report["mismatches"]["shipped"]["count"] == 0 and run.returncode == 0


def test_silu_mul_to_mx_on_column_slices_nan_blocks_and_refusals():
    import torchmx  # noqa: F401
    from torchmx import dtypes, mlp_ops
    from torchmx import env_variables as env
    from torchmx.mx_tensor import MXTensor
    torch.manual_seed(12)
    both = torch.randn(2, 40, 2 * 1024, device=DEV).to(torch.bfloat16)  # a stacked gate+up projection output
    gate, up = both.split([1024, 1024], dim=-1)
    gate[0, 3, 40] = float("nan")
    up[1, 5, 100] = float("inf")
    for hw_exact in ("False", "True"):
        prev, env.MX_EXACT_QUANTIZATION = env.MX_EXACT_QUANTIZATION, hw_exact
        try:
            got = mlp_ops.silu_mul_to_mx(gate, up, dtypes.float8_e4m3, 32)
            want = MXTensor.to_mx(torch.nn.functional.silu(gate) * up, dtypes.float8_e4m3, 32)
        finally:
            env.MX_EXACT_QUANTIZATION = prev
        assert torch.equal(got._scale_e8m0, want._scale_e8m0) and torch.equal(got._data, want._data)
        assert got._scale_e8m0[0, 3, 1] == 255 and got._scale_e8m0[1, 5, 3] == 255
    assert mlp_ops.silu_mul_to_mx(gate[..., :48], up[..., :48], dtypes.float8_e4m3, 32) is None          # cols % 32
    assert mlp_ops.silu_mul_to_mx(gate.float(), up.float(), dtypes.float8_e4m3, 32) is None              # dtype
    assert mlp_ops.silu_mul_to_mx(gate[..., 8:72], up[..., 8:72], dtypes.float8_e4m3, 32) is None         # rows not 32-byte aligned
    assert mlp_ops.silu_mul_to_mx(gate.transpose(0, 1), up.transpose(0, 1), dtypes.float8_e4m3, 32) is None  # no single row stride
