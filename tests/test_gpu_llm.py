"""GPU tier: `quantize_llm_` (reference: torchmx/quant_api.py:218-271) on a tiny random-init HF Llama -- attention and MLP
blocks swapped for their MX versions, Q / K / V / attention-weights quantized, both attention contractions on MX operands
(reference: layers/mx_llama_attention.py:195-243).

Parity: the tensor-core path must agree with the dequantize-then-bf16 path the reference itself executes (same quantized
operands; only fp32 accumulation order and one extra bf16 rounding of the scores differ), and the quantized model must stay
close to the bf16 model (SQNR, cf. the reference's linear-layer thresholds in tests/layers/conftest.py:10-20)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sqnr(ref, x):
    return float(20 * torch.log10(ref.float().norm() / (ref.float() - x.float()).norm()))


@pytest.fixture(scope="module")
def tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                      max_position_embeddings=512)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).to(DEV, torch.bfloat16).eval(), cfg


@pytest.mark.parametrize("qkv", [True, False])
def test_quantize_llm_swaps_blocks_and_matches_the_dequantize_path(tiny_llama, qkv):
    import copy
    import torchmx
    from torchmx import mx_gemm
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx.layers.mx_llama_attention import MXInferenceLlamaAttention, MXInferenceLlamaMLP
    from torchmx.quant_api import quantize_llm_
    model, cfg = tiny_llama
    ids = torch.randint(0, cfg.vocab_size, (2, 128), device=DEV)
    with torch.no_grad():
        ref = model(input_ids=ids).logits
    qm = copy.deepcopy(model)
    lin = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    e = MXConfig("float8_e4m3", 32)
    qa = QAttentionConfig(projection_config=lin, query_config=e, key_config=e, value_config=e, attention_weights_config=e) if qkv \
        else QAttentionConfig(projection_config=lin)
    quantize_llm_(qm, qa, lin)
    layer = qm.model.layers[0]
    assert type(layer.self_attn) is MXInferenceLlamaAttention and type(layer.mlp) is MXInferenceLlamaMLP
    assert all(type(getattr(layer.self_attn, n)) is MXInferenceLinear for n in ("q_proj", "k_proj", "v_proj", "o_proj"))
    assert type(qm.lm_head) is MXInferenceLinear and "qconfig" in repr(layer.self_attn)
    assert layer.self_attn.qconfig.is_qkv_quantization_enabled == qkv
    from torchmx import attention_ops
    before, before_attn = dict(mx_gemm.stats), dict(attention_ops.stats)
    with torch.no_grad():
        out_tc = qm(input_ids=ids).logits
    n_mm = mx_gemm.stats["tensor_core"] - before["tensor_core"]
    # 2 layers x (stacked q/k/v + o + stacked gate/up + down) + lm_head; when Q/K/V are quantized both attention contractions of a
    # layer run inside one K4b launch (head_dim 128, 128 keys)
    assert n_mm == 9 and mx_gemm.stats["fallback"] == before["fallback"]
    assert attention_ops.stats["flash_attention"] - before_attn["flash_attention"] == (2 if qkv else 0)
    mx_gemm.set_enabled(False)
    try:
        with torch.no_grad():
            out_deq = qm(input_ids=ids).logits
    finally:
        mx_gemm.set_enabled(True)
    assert _sqnr(out_deq, out_tc) > 30, _sqnr(out_deq, out_tc)
    assert _sqnr(ref, out_tc) > 12, _sqnr(ref, out_tc)


def test_mx_attention_decode_with_cache(tiny_llama):
    """prefill + a few decode steps through the MX attention block with HF's DynamicCache: same tokens as one full forward"""
    import copy
    import torchmx  # noqa: F401
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.quant_api import quantize_llm_
    model, cfg = tiny_llama
    qm = copy.deepcopy(model)
    lin = QLinearConfig(weights_config=MXConfig("float8_e4m3", 32), activations_config=MXConfig("float8_e4m3", 32))
    quantize_llm_(qm, QAttentionConfig(projection_config=lin), lin)
    ids = torch.randint(0, cfg.vocab_size, (1, 40), device=DEV)
    with torch.no_grad():
        full = qm(input_ids=ids).logits
        out = qm(input_ids=ids[:, :32], use_cache=True)
        steps = [out.logits[:, -1]]
        past = out.past_key_values
        for t in range(32, 39):
            out = qm(input_ids=ids[:, t:t + 1], past_key_values=past, use_cache=True)
            past = out.past_key_values
            steps.append(out.logits[:, -1])
    got = torch.stack(steps, 1)
    assert _sqnr(full[:, 31:39], got) > 25, _sqnr(full[:, 31:39], got)


def test_quantize_llm_qwen2_with_projection_biases():
    """Qwen2 (reference: layers/mx_qwen2_attention.py): q / k / v projections carry biases -> the bias epilogue of K3, plus the
    MX attention contractions and the fused softmax; tensor-core path vs the dequantize path the reference executes"""
    import copy
    from transformers import Qwen2Config, Qwen2ForCausalLM
    import torchmx  # noqa: F401
    from torchmx import attention_ops, mx_gemm
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.layers.mx_llama_attention import MXInferenceQwen2Attention, MXInferenceQwen2MLP
    from torchmx.quant_api import quantize_llm_
    cfg = Qwen2Config(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
                      max_position_embeddings=512, use_sliding_window=False)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    model = Qwen2ForCausalLM(cfg).to(DEV, torch.bfloat16).eval()
    for layer in model.model.layers:  # random-init biases are zero: make them matter
        for n in ("q_proj", "k_proj", "v_proj"):
            torch.nn.init.normal_(getattr(layer.self_attn, n).bias, std=0.5)
    ids = torch.randint(0, cfg.vocab_size, (2, 128), device=DEV)
    with torch.no_grad():
        ref = model(input_ids=ids).logits
    qm = copy.deepcopy(model)
    lin = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    e = MXConfig("float8_e4m3", 32)
    quantize_llm_(qm, QAttentionConfig(projection_config=lin, query_config=e, key_config=e, value_config=e, attention_weights_config=e), lin)
    layer = qm.model.layers[0]
    assert type(layer.self_attn) is MXInferenceQwen2Attention and type(layer.mlp) is MXInferenceQwen2MLP
    assert layer.self_attn.q_proj.bias is not None
    before, soft0, flash0 = dict(mx_gemm.stats), attention_ops.stats["fused_softmax"], attention_ops.stats["flash_attention"]
    with torch.no_grad():
        out_tc = qm(input_ids=ids).logits
    # 13 linears on the tensor cores; the attention of each layer (both contractions + softmax + quantization of P) is one K4b launch
    assert mx_gemm.stats["tensor_core"] - before["tensor_core"] == 9 and mx_gemm.stats["fallback"] == before["fallback"]  # q/k/v stacked
    assert attention_ops.stats["flash_attention"] == flash0 + 2 and attention_ops.stats["fused_softmax"] == soft0
    prev = attention_ops.set_flash_attention(False)  # the chain K4b replaces: two bmm launches + K4a per layer
    try:
        with torch.no_grad():
            out_chain = qm(input_ids=ids).logits
    finally:
        attention_ops.set_flash_attention(prev)
    assert mx_gemm.stats["tensor_core"] - before["tensor_core"] == 9 + 13 and attention_ops.stats["fused_softmax"] == soft0 + 2
    assert _sqnr(out_chain, out_tc) > 60, _sqnr(out_chain, out_tc)
    mx_gemm.set_enabled(False)
    try:
        with torch.no_grad():
            out_deq = qm(input_ids=ids).logits
    finally:
        mx_gemm.set_enabled(True)
    assert _sqnr(out_deq, out_tc) > 30, _sqnr(out_deq, out_tc)
    assert _sqnr(ref, out_tc) > 12, _sqnr(ref, out_tc)
    # decode-sized input: q/k/v (with their biases) and gate/up run as one stacked launch each; same result as separate launches
    from torchmx.layers import mx_llama_attention as mla
    att = layer.self_attn
    assert att.__dict__["_qkv"] is not None and att.__dict__["_qkv"].bias is not None
    assert att.k_proj.bias.data_ptr() == att.__dict__["_qkv"].bias[att.q_proj.out_features:].data_ptr()
    ids8 = ids[:, :4]
    n0 = mx_gemm.stats["tensor_core"]
    with torch.no_grad():
        stacked = qm(input_ids=ids8).logits
    # per layer qkv, o, gate_up, down + the Q.K^T bmm (P.V contracts over 4 padded positions: dequantize path), then lm_head
    assert mx_gemm.stats["tensor_core"] - n0 == 2 * 4 + 2 + 1
    prev, mla.STACKED_MAX_ROWS, mla.MLP_STACKED_MAX_ROWS = (mla.STACKED_MAX_ROWS, mla.MLP_STACKED_MAX_ROWS), -1, -1
    try:
        with torch.no_grad():
            separate = qm(input_ids=ids8).logits
    finally:
        mla.STACKED_MAX_ROWS, mla.MLP_STACKED_MAX_ROWS = prev
    assert _sqnr(separate, stacked) > 35, _sqnr(separate, stacked)


def test_stacked_projections_match_separate_launches(tiny_llama):
    """q/k/v and gate/up stacked into one launch each for decode-sized activations: same storage (no second copy, state_dict
    unchanged), outputs within accumulation noise of the separate launches (the split-K factor of the weight-streaming kernel
    depends on the number of output tiles); above the row cap only gate/up stack (its consumer K1b reads the column slices in
    place) and the outputs are bit-identical"""
    import copy
    import torchmx  # noqa: F401
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.layers import mx_llama_attention as mla
    from torchmx.quant_api import quantize_llm_
    model, cfg = tiny_llama
    lin = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    built = {}
    for fuse in (True, False):
        prev, mla.FUSE_PROJECTIONS = mla.FUSE_PROJECTIONS, fuse
        try:
            qm = copy.deepcopy(model)
            quantize_llm_(qm, QAttentionConfig(projection_config=lin), lin)
        finally:
            mla.FUSE_PROJECTIONS = prev
        built[fuse] = qm
    fused, plain = built[True], built[False]
    att, mlp = fused.model.layers[0].self_attn, fused.model.layers[0].mlp
    assert att.__dict__["_qkv"] is not None and mlp.__dict__["_gate_up"] is not None and plain.model.layers[0].self_attn.__dict__["_qkv"] is None
    assert att.k_proj.weight._data.data_ptr() == att.__dict__["_qkv"].weight._data[att.q_proj.out_features:].data_ptr()  # aliases, not copies
    sd_f, sd_p = fused.state_dict(), plain.state_dict()
    assert list(sd_f) == list(sd_p)
    for k in sd_f:
        a, b = sd_f[k], sd_p[k]
        if hasattr(a, "_data"):
            assert torch.equal(a._data, b._data) and torch.equal(a._scale_e8m0, b._scale_e8m0), k
    ids = torch.randint(0, cfg.vocab_size, (2, 128), device=DEV)
    from torchmx import mx_gemm
    n0 = mx_gemm.stats["tensor_core"]
    with torch.no_grad():
        out_f = fused(input_ids=ids).logits
    n1 = mx_gemm.stats["tensor_core"]
    with torch.no_grad():
        out_p = plain(input_ids=ids).logits
    assert n1 - n0 == 2 * 4 + 1 and mx_gemm.stats["tensor_core"] - n1 == 2 * 7 + 1   # 256 tokens: q/k/v and gate/up stacked
    assert _sqnr(out_p, out_f) > 40, _sqnr(out_p, out_f)
    prev_rows, mla.STACKED_MAX_ROWS = mla.STACKED_MAX_ROWS, 128  # the cap: above it q/k/v stay separate launches, bit-identical to `plain`
    try:
        n0 = mx_gemm.stats["tensor_core"]
        with torch.no_grad():
            out_c = fused(input_ids=ids).logits
        assert mx_gemm.stats["tensor_core"] - n0 == 2 * 6 + 1
        assert torch.equal(out_c, out_p)
    finally:
        mla.STACKED_MAX_ROWS = prev_rows
    ids8 = ids[:, :4]
    n0 = mx_gemm.stats["tensor_core"]
    with torch.no_grad():
        d_f = fused(input_ids=ids8).logits
    n1 = mx_gemm.stats["tensor_core"]
    with torch.no_grad():
        d_p = plain(input_ids=ids8).logits
    assert n1 - n0 == 2 * 4 + 1 and mx_gemm.stats["tensor_core"] - n1 == 2 * 7 + 1   # 8 tokens: q/k/v and gate/up stacked
    assert _sqnr(d_p, d_f) > 40, _sqnr(d_p, d_f)
    # a layer whose weight is replaced afterwards must not keep using the stale stacked copy
    att.k_proj.weight = torch.nn.Parameter(plain.model.layers[1].self_attn.k_proj.weight, requires_grad=False)
    plain.model.layers[0].self_attn.k_proj.weight = att.k_proj.weight
    with torch.no_grad():
        a, b = fused(input_ids=ids8).logits, plain(input_ids=ids8).logits
    assert _sqnr(b, a) > 40, _sqnr(b, a)  # (a stale stacked k_proj would be a different weight matrix altogether)
    assert _sqnr(d_f, a) < 30             # ... and the swap did change the output


def test_quantize_llm_then_pack_linear_runs_and_is_bit_identical(tiny_llama):
    """`quantize_llm_` followed by `pack_linear_`: the blocks' stacked q/k/v and gate/up weights are packed ONCE and the
    per-projection layers become views of that stream (nothing is left behind in the reference layout), prefill- and decode-sized
    forwards work and equal the unpacked model bit for bit; `unpack_linear_` restores the reference `state_dict`."""
    import copy
    import torchmx  # noqa: F401
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.layers.mx_linear import MXInferenceLinear
    from torchmx.layers.packed_linear import PackedMXLinear
    from torchmx.quant_api import pack_linear_, quantize_llm_, unpack_linear_
    model, cfg = tiny_llama
    lin = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    qm = copy.deepcopy(model)
    quantize_llm_(qm, QAttentionConfig(projection_config=lin), lin)
    sd_ref = {k: (v._data.clone(), v._scale_e8m0.clone()) for k, v in qm.state_dict().items() if hasattr(v, "_data")}
    ids = torch.randint(0, cfg.vocab_size, (2, 96), device=DEV)
    with torch.no_grad():
        want_prefill, want_decode = qm(input_ids=ids).logits, qm(input_ids=ids[:, :3]).logits
    n = pack_linear_(qm)
    assert n == 2 * 7 + 1
    att, mlp = qm.model.layers[0].self_attn, qm.model.layers[0].mlp
    assert all(type(getattr(att, k)) is PackedMXLinear for k in ("q_proj", "k_proj", "v_proj", "o_proj"))
    assert type(att.__dict__["_qkv"]) is PackedMXLinear and type(mlp.__dict__["_gate_up"]) is PackedMXLinear
    assert att.k_proj.weight_packed.data_ptr() == att.__dict__["_qkv"].weight_packed[att.q_proj.out_features:].data_ptr()
    assert not any(isinstance(m, MXInferenceLinear) for m in qm.modules())
    with torch.no_grad():
        assert torch.equal(qm(input_ids=ids).logits, want_prefill)
        assert torch.equal(qm(input_ids=ids[:, :3]).logits, want_decode)
    assert unpack_linear_(qm) == 2 * 7 + 1
    for k, v in qm.state_dict().items():
        if hasattr(v, "_data"):
            assert torch.equal(v._data, sd_ref[k][0]) and torch.equal(v._scale_e8m0, sd_ref[k][1]), k
    with torch.no_grad():
        assert torch.equal(qm(input_ids=ids).logits, want_prefill)


def test_left_padded_batch_with_sdpa_mask_has_no_nan(tiny_llama):
    """transformers' sdpa mask interface hands the attention block BOOLEAN masks in which the query rows of left-padding
    tokens may attend to nothing.  The MX attention block must not turn those rows into NaN (an all -inf softmax row would,
    and the NaN would reach every real token through the next layer's K / V): it fills with finfo.min like the reference's
    additive mask.  Real tokens of the padded sequence must match the same sequence run alone."""
    import copy
    import torchmx  # noqa: F401
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.quant_api import quantize_llm_
    model, cfg = tiny_llama
    qm = copy.deepcopy(model)
    qm.config._attn_implementation = "sdpa"
    lin = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    e = MXConfig("float8_e4m3", 32)
    quantize_llm_(qm, QAttentionConfig(projection_config=lin, query_config=e, key_config=e, value_config=e, attention_weights_config=e), lin)
    g = torch.Generator(device=DEV).manual_seed(4)
    ids = torch.randint(1, cfg.vocab_size, (2, 64), device=DEV, generator=g)
    mask = torch.ones(2, 64, dtype=torch.long, device=DEV)
    pad = 24
    mask[1, :pad] = 0  # sequence 1 is left-padded
    with torch.no_grad():
        out = qm(input_ids=ids, attention_mask=mask).logits
    assert not torch.isnan(out[0]).any() and not torch.isnan(out[1, pad:]).any(), "NaN leaked into real tokens"
    with torch.no_grad():
        alone = qm(input_ids=ids[1:, pad:], position_ids=torch.arange(64 - pad, device=DEV)[None]).logits
        both = qm(input_ids=ids, attention_mask=mask, position_ids=(mask.cumsum(-1) - 1).clamp(min=0)).logits
    assert not torch.isnan(both[1, pad:]).any()
    assert _sqnr(alone[0], both[1, pad:]) > 20, _sqnr(alone[0], both[1, pad:])


def test_decode_under_sdpa_groups_query_heads_instead_of_repeating_kv(tiny_llama, monkeypatch):
    """decode steps of a grouped-query model configured for sdpa: the block attends with the query heads of a key / value head
    on the query-length axis (no repeat_kv copies of the cache); same logits as transformers' own sdpa path up to the
    summation order of the attention kernel, and K1b feeds down_proj at decode sizes ([batch, 1, hidden] activations)"""
    import copy
    import torchmx  # noqa: F401
    from torchmx import mlp_ops
    from torchmx.config import MXConfig, QAttentionConfig, QLinearConfig
    from torchmx.layers import mx_llama_attention as mla
    from torchmx.quant_api import quantize_llm_
    from transformers.cache_utils import StaticCache
    model, cfg = tiny_llama
    qm = copy.deepcopy(model)
    qm.config._attn_implementation = "sdpa"
    lin = QLinearConfig(weights_config=MXConfig("float6_e3m2", 32), activations_config=MXConfig("float8_e4m3", 32))
    quantize_llm_(qm, QAttentionConfig(projection_config=lin), lin)
    for layer in qm.model.layers:
        layer.self_attn.config._attn_implementation = "sdpa"
    ids = torch.randint(0, cfg.vocab_size, (3, 40), device=DEV)
    outs = []
    for grouped in (True, False):
        monkeypatch.setattr(mla, "GROUPED_DECODE_SDPA", grouped)
        cache = StaticCache(config=qm.config, max_cache_len=64)
        n0 = mlp_ops.stats["fused_silu_mul"]
        with torch.no_grad():
            qm(input_ids=ids[:, :32], past_key_values=cache, use_cache=True)
            steps = [qm(input_ids=ids[:, 32 + i:33 + i], past_key_values=cache, use_cache=True).logits for i in range(8)]
        assert mlp_ops.stats["fused_silu_mul"] - n0 == 2 * 9  # prefill + 8 decode steps, two layers: gating + quantization in one launch
        outs.append(torch.cat(steps, 1))
    assert _sqnr(outs[1], outs[0]) > 35, _sqnr(outs[1], outs[0])


def test_additive_mask_cache_does_not_leak_graph_pool_tensors():
    """the boolean -> additive mask conversion is shared by the layers of one forward pass, also while a CUDA graph is captured,
    but a tensor made during a capture (it lives in the graph's pool) is never handed to a later eager call or capture"""
    import torchmx  # noqa: F401
    from torchmx.layers import mx_llama_attention as mla
    mask = torch.ones(2, 1, 4, 64, dtype=torch.bool, device=DEV).tril_(32)
    a0 = mla._additive_mask(mask, torch.bfloat16, 0)
    assert mla._additive_mask(mask, torch.bfloat16, 1) is a0          # eager: one conversion per mask object
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            c0 = mla._additive_mask(mask, torch.bfloat16, 0)
            c1 = mla._additive_mask(mask, torch.bfloat16, 1)
            c5 = mla._additive_mask(mask, torch.bfloat16, 5)
    assert c0 is not a0 and c1 is c0 and c5 is c0                    # the capture makes its own, once
    e = mla._additive_mask(mask, torch.bfloat16, 3)
    assert e is not c0                                               # ... and it does not escape the capture
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g2, stream=st):
            d0 = mla._additive_mask(mask, torch.bfloat16, 0)
    assert d0 is not c0 and d0 is not e
    assert torch.equal(e, torch.zeros_like(e).masked_fill_(~mask, torch.finfo(torch.bfloat16).min))
