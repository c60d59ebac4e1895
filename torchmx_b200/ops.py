"""aten overrides for MXTensor (dispatch table), API-compatible with /root/reference/torchmx/ops.py.

Compute ops (linear :29-41, mm/matmul :60-68, bmm :99-107, addmm :110-119): the reference
dequantizes both operands and calls the plain bf16 aten op.  Here the contraction runs on the
tcgen05 block-scaled tensor-core kernel (mxq_gemm) whenever the operands qualify -- FP element
types, block size 32 along the contraction dim for both operands, no padding -- and otherwise on
the reference's own recipe with the CUDA dequantize kernels feeding the aten op.

Layout ops (detach :44-57, expand :71-96, t :122-136, transpose :139-158, view :181-248,
_to_copy :251-276, sum :161-178) only rewrap `_data` / `_scale_e8m0` and keep `_block_dim` up to
date; they define the strided operands the compute ops receive.
"""
from __future__ import annotations

import torch
from torch.utils._pytree import tree_map

from . import dtypes
from .mx_tensor import MXTensor
from .utils import tensor_size_hp_to_fp4x2

# same global side effect as the reference (ops.py:16-19): bf16 GEMMs accumulate in fp32 end to end
torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = False
torch.backends.cuda.matmul.allow_fp16_reduced_precision_reduction = False

aten = torch.ops.aten
implements = MXTensor.implements


def _rewrap(old: MXTensor, scale, data, *, block_dim=None, orig_dtype=None) -> MXTensor:
    new = MXTensor(scale, data, old._elem_dtype, old._block_size, orig_dtype or old._orig_dtype, old._padding,
                   old._block_dim if block_dim is None else block_dim)
    # Layout ops produce short-lived views (F.linear decomposes into aten.t + aten.mm).  Remember the long-lived tensor
    # they came from: the tensor-core operand shadow of a weight is cached on THAT Python object (mx_gemm._operand_rows).
    new._mxq_origin = getattr(old, "_mxq_origin", old)
    return new


# ---- compute ops -----------------------------------------------------------------------------------
def _hp(x: MXTensor) -> torch.Tensor:
    return x.to_dtype(x._orig_dtype)


def _contract(aten_op, a: MXTensor, b: MXTensor, *extra_front, extra_back=()):
    """Shared body of the four compute overrides."""
    from . import mx_gemm  # late import: the GEMM host module needs MXTensor defined

    out = mx_gemm.contract(aten_op, a, b, extra_front, extra_back)
    if out is not None:
        return out
    # Not reached for plain CUDA MXTensors that advertise bf16.  What is left: torch.compile tracing (the inner tensors are fake;
    # the traced graph spells the reference's recipe with the two custom ops) and advertised dtypes other than bf16.
    return aten_op(*extra_front, _hp(a), _hp(b), *extra_back)


@implements([aten.linear.default])
def mx_linear(aten_op, types, args, kwargs=None):
    a, b = args[0], args[1]
    bias = args[2] if len(args) > 2 else None
    assert isinstance(a, MXTensor) and isinstance(b, MXTensor)
    return _contract(aten_op, a, b, extra_back=(bias,))


@implements([aten.mm.default, aten.matmul.default])
def mx_mm(aten_op, types, args, kwargs=None):
    a, b = args[0], args[1]
    assert isinstance(a, MXTensor) and isinstance(b, MXTensor)
    return _contract(aten_op, a, b)


@implements([aten.bmm.default])
def mx_bmm(aten_op, types, args, kwargs=None):
    a, b = args[0], args[1]
    assert isinstance(a, MXTensor) and isinstance(b, MXTensor)
    return _contract(aten_op, a, b)


@implements([aten.addmm.default])
def mx_addmm(aten_op, types, args, kwargs=None):
    bias, a, b = args[0], args[1], args[2]  # aten.addmm(bias, mat1, mat2)
    assert isinstance(a, MXTensor) and isinstance(b, MXTensor)
    return _contract(aten_op, a, b, bias)


@implements([aten.sum.dim_IntList])
def mx_cast_up_op(aten_op, types, args, kwargs=None):
    """Fallback: dequantize every MXTensor argument, then run the op (needed by addmm's bias
    gradient; reference: ops.py:161-178)."""
    unwrap = lambda x: _hp(x) if isinstance(x, MXTensor) else x  # noqa: E731
    return aten_op(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))


# ---- layout ops ------------------------------------------------------------------------------------
@implements([aten.detach.default])
def mx_desugar_op(aten_op, types, args, kwargs=None):
    old = args[0]
    return _rewrap(old, old._scale_e8m0, aten_op(old._data, *args[1:], **(kwargs or {})))


def _inner_sizes(old: MXTensor, logical_size, block_dim):
    """logical (outer) size -> (scale size, data size) along `block_dim` (reference: ops.py:79-86, 224-232)."""
    scale_size, data_size = list(logical_size), list(logical_size)
    scale_size[block_dim] = (scale_size[block_dim] + old._padding) // old._block_size
    if old._elem_dtype == dtypes.float4_e2m1:
        data_size = tensor_size_hp_to_fp4x2(data_size, block_dim)
    return scale_size, data_size


@implements([aten.expand.default])
def mx_expand(aten_op, types, args, kwargs=None):
    """Minimal expand needed by 4-D matmul before bmm (reference: ops.py:71-96)."""
    old = args[0]
    scale_size, data_size = _inner_sizes(old, args[1], old._block_dim)
    kw = kwargs or {}
    return _rewrap(old, aten_op(old._scale_e8m0, scale_size, *args[2:], **kw), aten_op(old._data, data_size, *args[2:], **kw))


@implements([aten.t.default])
def mx_t(aten_op, types, args, kwargs=None):
    old = args[0]
    assert old._block_dim in (0, 1)
    return _rewrap(old, old._scale_e8m0.t(), old._data.t(), block_dim=1 - old._block_dim)


@implements([aten.transpose.int])
def mx_transpose(aten_op, types, args, kwargs=None):
    old, d0, d1 = args[0], args[1], args[2]
    nd = old._data.dim()
    p0, p1 = d0 % nd, d1 % nd
    bd = old._block_dim
    new_bd = p1 if bd == p0 else (p0 if bd == p1 else bd)
    kw = kwargs or {}
    return _rewrap(old, aten_op(old._scale_e8m0, d0, d1, **kw), aten_op(old._data, d0, d1, **kw), block_dim=new_bd)


@implements([aten.view.default, aten._unsafe_view.default])
def mx_view_op(aten_op, types, args, kwargs=None):
    """view is only meaningful while blocks stay intact: blocked dim last (needed by aten.linear on
    >2-D inputs) or second-to-last of a 4-D tensor (attention matmuls); reference: ops.py:181-248."""
    old, new_size = args[0], list(args[1])
    if len(new_size) == 1 and old._padding > 0:
        raise AssertionError("View op is not supported when the tensor is padded and the new size is 1D")
    neg_bd = old._block_dim - old._data.dim()
    assert neg_bd >= -2, "View Op is supported only when block_dim is last/second_last dim"
    if neg_bd == -2:
        assert old._data.dim() == 4, "For the view op when block_dim is second last dim, the tensor must be 4D"
    new_bd = len(new_size) + neg_bd
    scale_size, data_size = _inner_sizes(old, new_size, new_bd)
    kw = kwargs or {}
    return _rewrap(old, aten_op(old._scale_e8m0, scale_size, *args[2:], **kw), aten_op(old._data, data_size, *args[2:], **kw),
                   block_dim=new_bd)


@implements([aten._to_copy.default])
def autocast_to_copy(aten_op, types, args, kwargs=None):
    """Under autocast an MXTensor is asked to change dtype: only the advertised dtype changes
    (reference: ops.py:251-276)."""
    old = args[0]
    assert isinstance(old, MXTensor)
    kwargs = kwargs or {}
    assert len(kwargs) == 1 and "dtype" in kwargs, "Only support dtype kwarg for autocast"
    assert kwargs["dtype"] in {torch.float16, torch.bfloat16}, "Only support floating point conversion for autocast w/ MXTensor"
    return _rewrap(old, old._scale_e8m0, old._data, orig_dtype=kwargs["dtype"])
