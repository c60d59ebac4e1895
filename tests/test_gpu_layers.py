"""GPU tier: layer-level SQNR floors of the MX attention / MLP blocks on the reference's tiny configurations (reference:
tests/layers/test_mx_llama_attention.py, test_mx_qwen2_attention.py with the tables of tests/layers/conftest.py:22-54): hidden
128, 2 heads (head_dim 64 -- the attention contractions then take the dequantize path, as every contraction does in the
reference), intermediate 128, hidden_states = rand(2, 128, 128), the eight activation x weight element-type combinations.
The floors are the reference's minus 3 dB: its tests are `@flaky(max_runs=5)` over random module initialisations, here one
seeded initialisation has to pass."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MARGIN = 3.0
COMBOS = {"0": ("float8_e4m3", "float6_e3m2"), "1": ("float8_e4m3", "float4_e2m1"), "2": ("float6_e3m2", "float6_e3m2"),
          "3": ("float6_e3m2", "float4_e2m1"), "4": ("float6_e2m3", "float6_e3m2"), "5": ("float6_e2m3", "float4_e2m1"),
          "6": ("float4_e2m1", "float6_e3m2"), "7": ("float4_e2m1", "float4_e2m1")}
ATTEN_LINEAR = {"0": 18, "1": 13, "2": 17, "3": 12, "4": 18, "5": 13, "6": 12, "7": 10}
ATTEN_ALL_QUANT = {"0": 17, "1": 11, "2": 16, "3": 12, "4": 17, "5": 12, "6": 12, "7": 8}
MLP = {"0": 16, "1": 9, "2": 14, "3": 8, "4": 16, "5": 9, "6": 10, "7": 7}


def _sqnr(ref, x):
    return float(20 * torch.log10(ref.float().norm() / (ref.float() - x.float()).norm()))


def _family(name):
    if name == "llama":
        from transformers.models.llama import modeling_llama as m
        from transformers import LlamaConfig as C
        return m, C, m.LlamaAttention, m.LlamaMLP, m.LlamaRotaryEmbedding
    from transformers.models.qwen2 import modeling_qwen2 as m
    from transformers import Qwen2Config as C
    return m, C, m.Qwen2Attention, m.Qwen2MLP, m.Qwen2RotaryEmbedding


def _setup(name):
    m, C, Att, Mlp, Rot = _family(name)
    cfg = C(hidden_size=128, num_key_value_heads=2, num_attention_heads=2, num_hidden_layers=2, intermediate_size=128)
    cfg._attn_implementation = "sdpa"  # no mask tensor: both the bf16 block and the MX block apply the causal rule themselves
    torch.manual_seed(42)
    x = torch.rand(2, 128, 128, dtype=torch.bfloat16, device=DEV)
    pos = torch.arange(128, device=DEV)[None]
    pe = Rot(cfg).to(DEV)(x, pos)
    return cfg, Att, Mlp, x, pe


def _lin(a, w):
    from torchmx.config import MXConfig, QLinearConfig
    return QLinearConfig(weights_config=MXConfig(w, 32), activations_config=MXConfig(a, 32))


@pytest.mark.parametrize("family", ["llama", "qwen2"])
@pytest.mark.parametrize("mode", list(COMBOS))
@pytest.mark.parametrize("all_quant", [False, True])
def test_attention_block_sqnr(family, mode, all_quant):
    import torchmx  # noqa: F401
    from torchmx.config import MXConfig, QAttentionConfig
    from torchmx.layers.mx_llama_attention import ATTENTION_LAYERS
    a, w = COMBOS[mode]
    cfg, Att, _, x, pe = _setup(family)
    torch.manual_seed(42)
    layer = Att(cfg, layer_idx=0).to(DEV, torch.bfloat16).eval()
    with torch.no_grad():
        hp = layer(x, position_embeddings=pe, attention_mask=None)[0]
    e = MXConfig(a, 32)
    qc = QAttentionConfig(projection_config=_lin(a, w), query_config=e, key_config=e, value_config=e, attention_weights_config=e) if all_quant \
        else QAttentionConfig(projection_config=_lin(a, w))
    q = ATTENTION_LAYERS[Att].from_float(layer, qc).eval()
    with torch.no_grad():
        out = q(x, position_embeddings=pe, attention_mask=None)[0]
    floor = (ATTEN_ALL_QUANT if all_quant else ATTEN_LINEAR)[mode] - MARGIN
    assert out.shape == hp.shape and _sqnr(hp, out) >= floor, _sqnr(hp, out)


@pytest.mark.parametrize("family", ["llama", "qwen2"])
@pytest.mark.parametrize("mode", list(COMBOS))
def test_mlp_block_sqnr(family, mode):
    import torchmx  # noqa: F401
    from torchmx.layers.mx_llama_attention import MLP_LAYERS
    a, w = COMBOS[mode]
    cfg, _, Mlp, x, _ = _setup(family)
    torch.manual_seed(42)
    layer = Mlp(cfg).to(DEV, torch.bfloat16).eval()
    with torch.no_grad():
        hp = layer(x)
    q = MLP_LAYERS[Mlp].from_float(layer, _lin(a, w)).eval()
    with torch.no_grad():
        out = q(x)
    assert out.shape == hp.shape and _sqnr(hp, out) >= MLP[mode] - MARGIN, _sqnr(hp, out)
