"""Tensor-parallel MX linears (BASELINE config 5): Megatron-style column / row sharding of
`MXInferenceLinear` (reference layer: /root/reference/torchmx/layers/mx_linear.py:8-95).

* column-parallel (q/k/v/gate/up): out_features is split, the input is replicated, no communication;
* row-parallel (o/down): in_features is split on a multiple of the MX block (32; 128 keeps the shard on
  the tensor-core path), every rank quantizes ITS activation shard and ITS weight shard -- bit-identical
  to slicing the unsharded MX tensors, because blocks never straddle a shard boundary -- and the bf16
  partial products are summed with one all-reduce over NVLink (NCCL).  Only the summation order differs
  from the single-GPU layer (G bf16 partials instead of one fp32 accumulation).

One process per GPU; `group=None` means the default process group.  With world_size 1 both classes
degenerate to `MXInferenceLinear`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..config import QLinearConfig
from .mx_linear import MXInferenceLinear


def shard_bounds(n: int, world: int, rank: int, multiple: int = 1) -> Tuple[int, int]:
    """[lo, hi) of `n` items owned by `rank`, cut points on multiples of `multiple`; `n` must divide evenly into
    world * multiple units so every rank gets the same shape (what NCCL collectives and CUDA graphs want)."""
    if n % (world * multiple) != 0:
        raise ValueError(f"cannot split {n} into {world} equal shards of a multiple of {multiple}")
    per = n // world
    return rank * per, (rank + 1) * per


def _world_rank(group) -> Tuple[int, int]:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


class ColumnParallelMXLinear(MXInferenceLinear):
    """Output features [lo, hi) of the full layer; forward needs no communication."""

    @classmethod
    @torch.no_grad()
    def from_float(cls, mod: torch.nn.Linear, qconfig: QLinearConfig, group=None, world_rank: Optional[Tuple[int, int]] = None):
        world, rank = world_rank or _world_rank(group)
        lo, hi = shard_bounds(mod.out_features, world, rank)
        shard = torch.nn.Linear(mod.in_features, hi - lo, bias=False, device="meta")
        shard.weight = torch.nn.Parameter(mod.weight.data[lo:hi].contiguous(), requires_grad=False)
        shard.bias = None if mod.bias is None else torch.nn.Parameter(mod.bias.data[lo:hi].contiguous(), requires_grad=False)
        new = super().from_float(shard, qconfig)
        new.tp_world, new.tp_rank, new.tp_group = world, rank, group
        new.full_out_features = mod.out_features
        return new


class RowParallelMXLinear(MXInferenceLinear):
    """Input features [lo, hi) of the full layer; forward = local MX matmul + all-reduce (sum) of the bf16 partials.
    The bias (if any) is added by rank 0 only, before the reduction, so it is counted once."""

    @classmethod
    @torch.no_grad()
    def from_float(cls, mod: torch.nn.Linear, qconfig: QLinearConfig, group=None, world_rank: Optional[Tuple[int, int]] = None):
        world, rank = world_rank or _world_rank(group)
        block = max(qconfig.weights_config.block_size, qconfig.activations_config.block_size)
        lo, hi = shard_bounds(mod.in_features, world, rank, multiple=block)
        shard = torch.nn.Linear(hi - lo, mod.out_features, bias=False, device="meta")
        shard.weight = torch.nn.Parameter(mod.weight.data[:, lo:hi].contiguous(), requires_grad=False)
        shard.bias = mod.bias if rank == 0 else None
        new = super().from_float(shard, qconfig)
        new.tp_world, new.tp_rank, new.tp_group = world, rank, group
        new.full_in_features = mod.in_features
        return new

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = super().forward(x)
        if self.tp_world > 1:
            import torch.distributed as dist
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.tp_group)
        return y
