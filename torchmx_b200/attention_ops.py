"""Host side of the fused attention-probability kernel (K4a, csrc/mxq_softmax.cu, C entry `mxq_softmax_quantize`).

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

stats = {"fused_softmax": 0, "unfused_softmax": 0}
_ENABLED = os.environ.get("MXQ_FUSED_SOFTMAX", "1") != "0"
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
