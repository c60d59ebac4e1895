"""Host side of the fused attention kernels: K4a (csrc/mxq_softmax.cu, C entry `mxq_softmax_quantize`: the chain between the two MX
matmuls) and K4b (csrc/mxq_flash_attention.cu, C entry `mxq_flash_attention`: both matmuls and the chain as one kernel).

`softmax_to_mx(scores, scaling, mask, causal, elem_dtype)` is the chain between the two MX matmuls of the reference's MX
attention block (torchmx/layers/mx_llama_attention.py:214-239: `/ sqrt(head_dim)`, `+ causal_mask`, fp32 softmax, `.to(bf16)`,
`MXTensor.to_mx(..., attention_weights_config)`) as one pass over the bf16 scores.  There is no CPU / PyTorch fallback: when the
shape does not qualify the function returns None and the caller runs the unfused chain of K1 + aten ops.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _C, dtypes
from . import env_variables as env
from .mx_tensor import MXTensor, _require_cuda, _stream_ptr

stats = {"fused_softmax": 0, "unfused_softmax": 0, "flash_attention": 0}
_ENABLED = os.environ.get("MXQ_FUSED_SOFTMAX", "1") != "0"
_FLASH = os.environ.get("MXQ_FLASH_ATTENTION", "1") != "0"
MAX_KV = 32768


def set_fused_softmax(on: bool) -> bool:
    """Switch the fused kernel on / off (tests compare the two paths); returns the previous setting."""
    global _ENABLED
    prev, _ENABLED = _ENABLED, bool(on)
    return prev


def softmax_to_mx(scores: torch.Tensor, scaling: float, mask: Optional[torch.Tensor], causal: bool, elem_dtype: dtypes.DType,
                  block_size: int = 32) -> Optional[MXTensor]:
    """bf16 scores [batch, heads, q_len, kv_len] -> MXTensor of softmax(scores * scaling + mask) quantized along kv.

    mask: None or an additive bf16 tensor broadcastable to the scores ([b | 1, h | 1, q_len, >= kv_len], kv stride 1);
    causal: additionally hide kv index j > q + (kv_len - q_len).  Returns None when the kernel does not apply.
    """
    if not _ENABLED or block_size != 32 or scores.dim() != 4 or scores.dtype != torch.bfloat16 or not scores.is_contiguous():
        return None
    b, h, q_len, kv_len = scores.shape
    if kv_len % 32 or kv_len > MAX_KV or scores.numel() == 0 or scores.data_ptr() % 32:
        return None
    _require_cuda(scores, "softmax_to_mx")
    a = _C.SoftmaxArgs()
    if mask is not None:
        if mask.dtype != torch.bfloat16 or mask.dim() != 4 or mask.device != scores.device or mask.shape[-1] < kv_len \
                or mask.shape[2] != q_len or mask.shape[0] not in (1, b) or mask.shape[1] not in (1, h) or (kv_len > 1 and mask.stride(3) != 1):
            return None
        a.mask = mask.data_ptr()
        a.mask_stride_b = 0 if mask.shape[0] == 1 else mask.stride(0)
        a.mask_stride_h = 0 if mask.shape[1] == 1 else mask.stride(1)
        a.mask_stride_q = mask.stride(2)
    is_fp4 = elem_dtype == dtypes.float4_e2m1
    codes = torch.empty((b, h, q_len, kv_len // 2 if is_fp4 else kv_len), dtype=torch.int8 if elem_dtype == dtypes.int8 else torch.uint8,
                        device=scores.device)
    scales = torch.empty((b, h, q_len, kv_len // 32), dtype=torch.uint8, device=scores.device)
    a.scores = scores.data_ptr()
    a.batch, a.heads, a.q_len, a.kv_len = b, h, q_len, kv_len
    a.scaling = float(scaling)
    a.causal = 1 if causal else 0
    a.elem = dtypes.ELEM_ID[elem_dtype.name]
    a.flags = _C.FLAG_HW_EXACT if (elem_dtype in dtypes.SUPPORTED_FP_ELEM_DTYPES and env.MX_EXACT_QUANTIZATION == "True") else 0
    a.codes, a.scales = codes.data_ptr(), scales.data_ptr()
    rc = _C.lib().mxq_softmax_quantize(a, scores.device.index, _stream_ptr(scores))
    if rc == _C.ERR_UNSUPPORTED_SHAPE:
        return None
    _C.check(rc, "mxq_softmax_quantize")
    stats["fused_softmax"] += 1
    return MXTensor(scales, codes, elem_dtype, 32, scores.dtype)


def set_flash_attention(on: bool) -> bool:
    """Switch K4b on / off (tests compare it with the K3 -> K4a -> K3 chain); returns the previous setting."""
    global _FLASH
    prev, _FLASH = _FLASH, bool(on)
    return prev


def _byte_operand(t: MXTensor):
    """contiguous MXTensor blocked along its last dim -> (one-byte-per-element codes, MXQ_OPERAND_* format, scales) or None"""
    from . import mx_gemm
    if mx_gemm._DISABLED or torch.compiler.is_compiling() or not mx_gemm._qualifies(t) or t._block_dim != t._data.dim() - 1 or not t._data.is_contiguous() or not t._scale_e8m0.is_contiguous():
        return None
    codes, fmt = mx_gemm._operand_rows(t._data, t._elem_dtype, None, packed=False)  # e4m3 / e5m2 as they are, fp6 / fp4 as E4M3 bytes
    return codes, fmt, t._scale_e8m0


def flash_attention(q_mx: MXTensor, k_mx: MXTensor, vt_mx: MXTensor, scaling: float, mask: Optional[torch.Tensor], causal: bool,
                    p_elem_dtype: dtypes.DType, p_block_size: int = 32, return_probs: bool = False):
    """The reference's MX attention (torchmx/layers/mx_llama_attention.py:195-243) in one launch.

    q_mx [b, h, q, 128] and k_mx [b, h_kv, kv, 128] quantized along head_dim, vt_mx [b, h_kv, 128, kv] = V quantized along the
    key axis (the reference's quantize-the-transpose), all contiguous with block size 32; mask: None or additive bf16
    broadcastable to [b, h, q, kv].  Returns the attention output already in the [b, q, h, 128] layout the reference transposes
    to next (bf16) -- plus the MXTensor of P when `return_probs` -- or None when the kernel does not apply (the caller then
    runs the K3 -> K4a -> K3 chain)."""
    if not _FLASH or p_block_size != 32 or p_elem_dtype == dtypes.int8 or q_mx.dim() != 4 or k_mx.dim() != 4 or vt_mx.dim() != 4:
        return None
    b, h, q_len, d = q_mx.shape
    hk, kv_len = k_mx.shape[1], k_mx.shape[2]
    if d != 128 or k_mx.shape != (b, hk, kv_len, d) or vt_mx.shape != (b, hk, d, kv_len) or h % hk or kv_len % 32 or kv_len < q_len or q_len == 0:
        return None
    masked = causal or mask is not None
    if kv_len > (8192 if masked else 1024):  # (rows whose block sums K4a adds in an order K4b does not reproduce)
        return None
    ops = [_byte_operand(t) for t in (q_mx, k_mx, vt_mx)]
    if any(o is None for o in ops):
        return None
    dev = q_mx._data.device
    a = _C.AttentionArgs()
    if mask is not None:
        if mask.dtype != torch.bfloat16 or mask.dim() != 4 or mask.device != dev or mask.shape[-1] < kv_len or mask.shape[2] != q_len \
                or mask.shape[0] not in (1, b) or mask.shape[1] not in (1, h) or (kv_len > 1 and mask.stride(3) != 1):
            return None
        a.mask = mask.data_ptr()
        a.mask_stride_b = 0 if mask.shape[0] == 1 else mask.stride(0)
        a.mask_stride_h = 0 if mask.shape[1] == 1 else mask.stride(1)
        a.mask_stride_q = mask.stride(2)
    (a.q_codes, a.q_format, a.q_scales), (a.k_codes, a.k_format, a.k_scales), (a.vt_codes, a.v_format, a.vt_scales) = \
        [(c.data_ptr(), f, s.data_ptr()) for c, f, s in ops]
    a.batch, a.heads, a.kv_heads, a.q_len, a.kv_len, a.head_dim = b, h, hk, q_len, kv_len, d
    a.scaling, a.causal = float(scaling), 1 if causal else 0
    # One query position (decode) with grouped-query heads: the query heads that share a key / value head become the ROWS of one
    # query tile ([b, h, 1, d] is [b, h_kv, groups, d] in memory), so a CTA works on `groups` live rows instead of one and the grid
    # shrinks by that factor.  Same rows, same keys, same arithmetic per row; the mask row is shared by the group.
    grouped = q_len == 1 and h > hk and not causal and (mask is None or mask.shape[1] == 1)
    if grouped:
        a.heads, a.q_len, a.mask_stride_q = hk, h // hk, 0
    a.p_elem = dtypes.ELEM_ID[p_elem_dtype.name]
    a.flags = _C.FLAG_HW_EXACT if (p_elem_dtype in dtypes.SUPPORTED_FP_ELEM_DTYPES and env.MX_EXACT_QUANTIZATION == "True") else 0
    out = torch.empty((b, q_len, h, d), dtype=torch.bfloat16, device=dev)
    a.out, a.out_batch_stride, a.out_row_stride, a.out_head_stride = out.data_ptr(), out.stride(0), out.stride(1), out.stride(2)
    if grouped:  # element (batch, key head, group row g, channel) of the kernel's view is head (key head * groups + g) of the one token
        a.out_row_stride, a.out_head_stride = out.stride(2), (h // hk) * out.stride(2)
    probs = None
    if return_probs:
        is_fp4 = p_elem_dtype == dtypes.float4_e2m1
        p_codes = torch.empty((b, h, q_len, kv_len // 2 if is_fp4 else kv_len), dtype=torch.uint8, device=dev)
        p_scales = torch.empty((b, h, q_len, kv_len // 32), dtype=torch.uint8, device=dev)
        a.p_codes, a.p_scales = p_codes.data_ptr(), p_scales.data_ptr()
        probs = MXTensor(p_scales, p_codes, p_elem_dtype, 32, torch.bfloat16)
    rc = _C.lib().mxq_flash_attention(a, dev.index, _stream_ptr(q_mx._data))
    if rc == _C.ERR_UNSUPPORTED_SHAPE:
        return None
    _C.check(rc, "mxq_flash_attention")
    stats["flash_attention"] += 1
    return (out, probs) if return_probs else out
