"""MXInferenceLinear: nn.Linear whose weight is pre-quantized to MX and whose activation is
quantized on every forward (reference: /root/reference/torchmx/layers/mx_linear.py:8-95).

forward = K1 (activation quantize) -> MX matmul (aten.linear / addmm override -> K3 tensor-core
kernel, or the dequantize path when the operands do not qualify).
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from .. import env_variables as env
from .. import mx_gemm
from ..config import QLinearConfig
from ..mx_tensor import MXTensor


# q/k/v (and gate/up) projections of a decoder layer quantize the SAME activation tensor with the same config; the
# reference quantizes it once per layer (mx_linear.py:63-66), i.e. three (two) times.  One entry is remembered: if the next
# layer is handed the very same tensor (object identity, storage pointer, geometry and version counter all equal) the MX
# tensor is reused -- bit-identical by construction, one K1 launch instead of three.  The entry keeps `x` alive, so its
# storage cannot be recycled under the key; MXQ_ACT_REUSE=0 turns it off.
_ACT_REUSE = os.environ.get("MXQ_ACT_REUSE", "1") != "0"
_last_act = None


def _quantize_activation(x: torch.Tensor, elem_dtype, block_size: int) -> MXTensor:
    global _last_act
    if not _ACT_REUSE or type(x) is not torch.Tensor:
        return MXTensor.to_mx(x, elem_dtype, block_size)
    key = (x.data_ptr(), x._version, tuple(x.shape), tuple(x.stride()), x.dtype, elem_dtype.name, block_size)
    hit = _last_act
    if hit is not None and hit[0] is x and hit[1] == key:
        return hit[2]
    x_mx = MXTensor.to_mx(x, elem_dtype, block_size)
    _last_act = (x, key, x_mx)
    return x_mx


def clear_activation_cache() -> None:
    global _last_act
    _last_act = None


class MXInferenceLinear(torch.nn.Linear):
    def extra_repr(self) -> str:
        return f"{super().extra_repr()}, qconfig={self.qconfig}"

    @classmethod
    @torch.no_grad()
    def from_float(cls, mod: torch.nn.Linear, qconfig: QLinearConfig) -> "MXInferenceLinear":
        """Swap-in constructor (reference: mx_linear.py:21-59): the module skeleton is built on the
        meta device, the weight is quantized once unless it lives on `meta` itself (accelerate
        offload), the bias object is shared with the source module."""
        # Same result as building `cls(in, out, bias=False)` on the meta device and swapping the weight in (the
        # reference's recipe), without running nn.Linear.__init__ / reset_parameters per layer: whole-model
        # quantization is host-bound otherwise (the quantize kernel of a 4096 x 14336 weight takes ~30 us).
        new = cls.__new__(cls)
        torch.nn.Module.__init__(new)
        new.in_features, new.out_features = mod.in_features, mod.out_features
        new.qconfig = qconfig
        w = mod.weight.data
        if w.device.type != "meta":
            wc = qconfig.weights_config
            new.weight = torch.nn.Parameter(MXTensor.to_mx(w, wc.elem_dtype, wc.block_size), requires_grad=False)
        else:
            new.weight = torch.nn.Parameter(torch.empty(mod.out_features, mod.in_features, device="meta", dtype=w.dtype), requires_grad=False)
        if mod.bias is None:
            new.register_parameter("bias", None)
        else:
            new.bias = mod.bias
        return new

    def _weight_mx(self) -> MXTensor:
        if isinstance(self.weight, MXTensor):  # (a Parameter made from an MXTensor IS that subclass; `.data` would dispatch a detach)
            return self.weight
        w = self.weight.data
        # weights that were on `meta` at conversion time arrive high-precision (often fp32) at call time
        # (reference: mx_linear.py:68-92): quantize on the fly, leave self.weight untouched
        wc = self.qconfig.weights_config
        return MXTensor.to_mx(w.to(torch.bfloat16), wc.elem_dtype, wc.block_size)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ac = self.qconfig.activations_config
        bias = self.bias
        if not isinstance(self.weight, MXTensor) and bias is not None:
            bias = bias.to(torch.bfloat16)
        w_mx = self._weight_mx()
        if ac.elem_dtype_name == "float8_e4m3" and ac.block_size == 32:
            # decode-sized activations: quantization fused into the weight-streaming GEMM (one launch, no code round trip)
            out = mx_gemm.linear_fused_act_quant(x, w_mx, bias, env.MX_EXACT_QUANTIZATION == "True")
            if out is not None:
                return out
        x_mx = _quantize_activation(x, ac.elem_dtype, ac.block_size)
        # F.linear(x_mx, w_mx, bias) reaches the same kernel through the dispatcher (aten.t + aten.mm / addmm on MXTensor
        # views, ~50 us of host time per layer); hand the operands to the tensor-core path directly when they qualify
        out = mx_gemm.try_tensor_core(torch.ops.aten.linear.default, x_mx, w_mx, (), (bias,), count_fallback=False)
        if out is not None:
            return out
        return F.linear(x_mx, w_mx, bias)
