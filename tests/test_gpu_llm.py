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
    before = dict(mx_gemm.stats)
    with torch.no_grad():
        out_tc = qm(input_ids=ids).logits
    n_mm = mx_gemm.stats["tensor_core"] - before["tensor_core"]
    # 2 layers x (4 projections + 3 MLP linears) + lm_head, plus 2 attention contractions per layer when Q/K/V are quantized
    assert n_mm == 15 + (4 if qkv else 0) and mx_gemm.stats["fallback"] == before["fallback"]
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
