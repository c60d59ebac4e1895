"""CPU tier: the C oracle against the committed reference outputs (tests/golden, produced by
oracle/gen_golden.py from the unmodified reference), plus hand-written known-answer vectors taken
from the reference's own test-suite."""
import numpy as np
import pytest

from oracle import grids
from tests.util import ELEMS, FP_ELEMS, MODES, assert_bits_equal, sha


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("elem", ELEMS)
def test_quantize_grid_small_matches_reference_digest(oracle, digests, elem, mode):
    grid = grids.quant_grid_small()
    assert sha(grid) == digests["quant_grid_small_input"]
    scales, codes = oracle.quantize(grid, elem, 32, hw_exact=(mode == "hw_exact"), threads=4)
    assert sha(scales) == digests[f"quant_grid_small/{elem}/{mode}/scales"]
    assert sha(codes) == digests[f"quant_grid_small/{elem}/{mode}/codes"]


@pytest.mark.parametrize("target", ["bf16", "f32"])
@pytest.mark.parametrize("elem", ELEMS)
def test_dequantize_grid_matches_reference_digest(oracle, digests, elem, target):
    from tests.util import canon_nan
    codes, scales = grids.dequant_grid(elem)
    out = oracle.dequantize(codes, scales, elem, 32, target, 1)
    out = out.view(np.uint32) if target == "f32" else out
    assert sha(canon_nan(out)) == digests[f"dequant_grid/{elem}/{target}"]


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("elem", ELEMS)
def test_structured_cases_match_reference_fixture(oracle, fixtures, elem, mode):
    for name, (bits, bs) in grids.structured_cases().items():
        key = f"struct/{name}/{elem}/{mode}/codes"
        if key not in fixtures:
            assert elem == "float4_e2m1" and bits.size % 2
            continue
        assert np.array_equal(fixtures[f"struct/{name}/x"], bits)
        scales, codes = oracle.quantize(bits, elem, bs, hw_exact=(mode == "hw_exact"))
        assert_bits_equal(scales, fixtures[f"struct/{name}/{elem}/{mode}/scales"], f"{name} scales")
        assert_bits_equal(codes, fixtures[key], f"{name} codes")


@pytest.mark.parametrize("elem", ELEMS)
def test_readme_example_matches_reference_fixture(oracle, fixtures, elem):
    """BASELINE config 1: 128x128 bf16 randn, block 32 -> to_mx -> to_dtype."""
    x = fixtures["readme/x"]
    for mode in MODES:
        scales, codes = oracle.quantize(x, elem, 32, hw_exact=(mode == "hw_exact"))
        assert_bits_equal(scales, fixtures[f"readme/{elem}/{mode}/scales"])
        assert_bits_equal(codes, fixtures[f"readme/{elem}/{mode}/codes"])
    assert_bits_equal(oracle.dequantize(codes, scales, elem, 32, "bf16"), fixtures[f"readme/{elem}/bf16"])
    assert_bits_equal(oracle.dequantize(codes, scales, elem, 32, "f32").view(np.uint32), fixtures[f"readme/{elem}/f32"])


def _bf16(sign, exp, man):
    return ((np.asarray(sign, np.uint16) << 15) | (np.asarray(exp, np.uint16) << 7) | np.asarray(man, np.uint16)).astype(np.uint16)


# ---- known-answer vectors written by hand in the reference's tests ----------------------------------
@pytest.mark.parametrize("mode", list(MODES))
def test_kat_e4m3_normal_to_normal(oracle, mode):
    """/root/reference/tests/test_mx_quantization.py:12-47"""
    man = [0b1111111, 0b0001010, 0b1000001, 0b1, 0b0101010, 0]
    sgn = np.array([1, 0, 0, 1, 0, 0])
    exp = np.array([[5, 5, 5, 5, 5, 19], [100, 100, 100, 100, 100, 111], [240, 240, 240, 240, 240, 249]])
    x = _bf16(sgn[None, :], exp, np.array(man)[None, :])
    gt_man = np.array([0b0, 0b001, 0b100, 0b0, 0b011, 0])
    gt_exp = np.array([[2, 1, 1, 1, 1, 15], [5, 4, 4, 4, 4, 15], [7, 6, 6, 6, 6, 15]])
    gt = ((sgn[None, :] << 7) | (gt_exp << 3) | gt_man[None, :]).astype(np.uint8)
    scales, codes = oracle.quantize(x, "float8_e4m3", 6, hw_exact=(mode == "hw_exact"))
    assert np.array_equal(codes, gt)
    assert np.array_equal(scales, np.array([[11], [103], [241]], dtype=np.uint8))


@pytest.mark.parametrize("mode", list(MODES))
def test_kat_e4m3_normal_to_subnormal(oracle, mode):
    """/root/reference/tests/test_mx_quantization.py:75-110"""
    man = np.array([0b1111111, 0b0001010, 0b1000001, 0b1, 0b0101010, 0])
    sgn = np.array([1, 0, 0, 1, 0, 1])
    exp = np.full((3, 6), 100)
    exp[:, -1] = [118, 116, 115]
    x = _bf16(sgn[None, :], exp, man[None, :])
    gt_man = np.array([[1, 1, 1, 1, 1, 0], [0b100, 0b010, 0b011, 0b010, 0b011, 0], [0b0, 0b100, 0b110, 0b100, 0b101, 0]])
    gt_exp = np.array([[0, 0, 0, 0, 0, 15], [0, 0, 0, 0, 0, 15], [1, 0, 0, 0, 0, 15]])
    gt = ((sgn[None, :] << 7) | (gt_exp << 3) | gt_man).astype(np.uint8)
    scales, codes = oracle.quantize(x, "float8_e4m3", 6, hw_exact=(mode == "hw_exact"))
    assert np.array_equal(codes, gt)
    assert np.array_equal(scales.reshape(-1), np.array([110, 108, 107], dtype=np.uint8))


@pytest.mark.parametrize("mode", list(MODES))
def test_kat_e4m3_bf16_subnormal_inputs(oracle, mode):
    """/root/reference/tests/test_mx_quantization.py:147-185"""
    man = np.array([0b1111111, 0b0001010, 0b1000001, 0b0110011, 0b0101010, 0])
    sgn = np.array([0, 1, 0, 1, 0, 1])
    exp = np.zeros((3, 6), dtype=np.int64)
    exp[:, -1] = [12, 13, 14]
    x = _bf16(sgn[None, :], exp, man[None, :])
    gt_man = np.array([[0b0, 0b101, 0b000, 0b101, 0b010, 0], [0b0, 0b10, 0b0, 0b101, 0b010, 0], [0b0, 0b1, 0b0, 0b110, 0b101, 0]])
    gt_exp = np.array([[4, 0, 3, 2, 2, 15], [3, 0, 2, 1, 1, 15], [2, 0, 1, 0, 0, 15]])
    gt = ((sgn[None, :] << 7) | (gt_exp << 3) | gt_man).astype(np.uint8)
    scales, codes = oracle.quantize(x, "float8_e4m3", 6, hw_exact=(mode == "hw_exact"))
    assert np.array_equal(codes, gt)
    assert np.array_equal(scales.reshape(-1), np.array([4, 5, 6], dtype=np.uint8))


@pytest.mark.parametrize("mode", list(MODES))
def test_kat_e3m2_normal_to_normal_and_subnormal(oracle, mode):
    """/root/reference/tests/test_mx_quantization.py:211-246 and :274-310 (code in bits [5:0])"""
    man = np.array([0b1111111, 0b0011010, 0b1000001, 0b1, 0b0111010, 0])
    sgn = np.array([1, 0, 0, 1, 0, 1])
    exp = np.array([[5, 5, 5, 5, 5, 11], [100, 100, 100, 100, 100, 103], [250, 250, 250, 250, 250, 251]])
    x = _bf16(sgn[None, :], exp, man[None, :])
    gt_man = np.array([0b0, 0b01, 0b10, 0b0, 0b10, 0])
    gt_exp = np.array([[2, 1, 1, 1, 1, 7], [5, 4, 4, 4, 4, 7], [7, 6, 6, 6, 6, 7]])
    gt = ((sgn[None, :] << 5) | (gt_exp << 2) | gt_man[None, :]).astype(np.uint8)
    scales, codes = oracle.quantize(x, "float6_e3m2", 6, hw_exact=(mode == "hw_exact"))
    assert np.array_equal(codes, gt)
    assert np.array_equal(scales.reshape(-1), np.array([7, 99, 247], dtype=np.uint8))

    exp = np.full((3, 6), 100)
    exp[:, -1] = [109, 108, 107]
    x = _bf16(sgn[None, :], exp, man[None, :])
    gt_man = np.array([[1, 1, 1, 1, 1, 0], [0b10, 0b1, 0b10, 0b1, 0b1, 0], [0b0, 0b10, 0b11, 0b10, 0b11, 0]])
    gt_exp = np.array([[0, 0, 0, 0, 0, 7], [0, 0, 0, 0, 0, 7], [1, 0, 0, 0, 0, 7]])
    gt = ((sgn[None, :] << 5) | (gt_exp << 2) | gt_man).astype(np.uint8)
    scales, codes = oracle.quantize(x, "float6_e3m2", 6, hw_exact=(mode == "hw_exact"))
    assert np.array_equal(codes, gt)
    assert np.array_equal(scales.reshape(-1), np.array([105, 104, 103], dtype=np.uint8))


@pytest.mark.parametrize("elem,maxv", [("float8_e4m3", 448.0), ("float6_e3m2", 28.0), ("float6_e2m3", 7.5), ("float4_e2m1", 6.0)])
def test_kat_saturation_roundtrip(oracle, elem, maxv):
    """values just below the next binade saturate to +-max * scale
    (/root/reference/tests/test_mx_quantization.py:49-73, 248-272)"""
    x = _bf16([1, 0, 1, 0], [100, 100, 100, 100], [0b1111110, 0b1111111, 0b1111111, 0b1111110])
    pow2 = {"float8_e4m3": 8, "float6_e3m2": 4, "float6_e2m3": 2, "float4_e2m1": 2}[elem]
    scales, codes = oracle.quantize(x[None, :], elem, 4)
    assert scales.reshape(-1)[0] == 100 - pow2
    y = oracle.bf16_bits_to_f32(oracle.dequantize(codes, scales, elem, 4, "bf16"))
    want = np.array([-1, 1, -1, 1], dtype=np.float32) * maxv * np.float32(2.0) ** (100 - pow2 - 127)
    assert np.array_equal(y.reshape(-1), want)


def test_nan_blocks(oracle, fixtures):
    """Inf/NaN anywhere in a block -> scale 255, codes +0, dequantized block all NaN
    (/root/reference/tests/test_mx_tensor.py:102-160)"""
    x = fixtures["special/x"]
    for elem in ELEMS:
        for mode in MODES:
            scales, codes = oracle.quantize(x, elem, 4, hw_exact=(mode == "hw_exact"))
            assert_bits_equal(scales, fixtures[f"special/{elem}/{mode}/scales"])
            assert_bits_equal(codes, fixtures[f"special/{elem}/{mode}/codes"])
            assert (scales == 255).all()
        out = oracle.bf16_bits_to_f32(oracle.dequantize(codes, scales, elem, 4, "bf16"))
        assert np.isnan(out).all()


def test_hw_exact_differs_from_simulated_only_in_nan_blocks(oracle, digests):
    grid = grids.quant_grid_small()
    for elem in FP_ELEMS:
        s0, c0 = oracle.quantize(grid, elem, 32, hw_exact=False, threads=4)
        s1, c1 = oracle.quantize(grid, elem, 32, hw_exact=True, threads=4)
        assert np.array_equal(s0, s1)
        per = 16 if elem == "float4_e2m1" else 32
        diff = (c0 != c1).reshape(-1, per).any(axis=1)
        assert diff.any(), "the quirk should be visible on the grid"
        assert (s0.reshape(-1)[diff] == 255).all()


def test_gemm_oracle_small(oracle):
    rng = np.random.default_rng(0)
    a = oracle.f32_to_bf16_bits(rng.standard_normal((5, 64)).astype(np.float32))
    b = oracle.f32_to_bf16_bits(rng.standard_normal((7, 64)).astype(np.float32))
    got = oracle.gemm_nt(a, b)
    want = oracle.bf16_bits_to_f32(a).astype(np.float64) @ oracle.bf16_bits_to_f32(b).astype(np.float64).T
    assert np.allclose(got, want, rtol=1e-6, atol=1e-6)
