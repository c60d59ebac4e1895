"""Tensor-parallel MX linears (BASELINE config 5): Megatron-style column / row sharding of
`MXInferenceLinear` (reference layer: /root/reference/torchmx/layers/mx_linear.py:8-95).

* column-parallel (q/k/v/gate/up): out_features is split, the input is replicated, no communication;
* row-parallel (o/down): in_features is split on a multiple of the MX block (32; 128 keeps the shard on
  the tensor-core path), every rank quantizes ITS activation shard and ITS weight shard -- bit-identical
  to slicing the unsharded MX tensors, because blocks never straddle a shard boundary -- and the bf16
  partial products are summed with one all-reduce over NVLink (NCCL).  Only the summation order differs
  from the single-GPU layer (G bf16 partials instead of one fp32 accumulation).

Fused mode (`RowParallelMXLinear.enable_fused_allreduce(pool)`): the GEMM epilogue itself performs the reduction -- every
rank adds its bf16 partial tile into a symmetric output buffer of ALL ranks through the NVLink multicast address
(`multimem.red`, the sum is formed in the NVSwitch), tile by tile while the remaining tiles are still being computed; one
cross-rank barrier replaces the NCCL all-reduce kernel and the partials never take a round trip through HBM.

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


class FusedAllReducePool:
    """Three rotating symmetric [max_rows, features] bf16 output buffers (torch symmetric memory: mapped on every rank of the
    group, with a multicast address) shared by all row-parallel layers of one width.

    Step k uses buffer k % 3.  Before the launch of step k this rank zeroes the rows buffer (k+1) % 3 had dirtied (last used
    at step k-2, long consumed on this stream); the other ranks can only add into that buffer at step k+1, i.e. after the
    barrier of step k, which this rank enters after the zeroing in stream order -- so one barrier per layer is enough.  The
    tensor a layer returns is a view of the buffer: it is valid until two more fused layers of the same pool have run."""

    def __init__(self, features: int, max_rows: int, group=None, fused_max_rows: int = 128):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self.features, self.max_rows, self.group = features, max_rows, group
        # Every rank PUSHES its whole partial to every rank (the multicast add is applied at each destination), so a GPU
        # receives world x rows x features x 2 bytes per layer: latency-optimal for decode-sized activations (one launch, one
        # barrier; 30 % faster than NCCL at 8 GPUs, batch 32), bandwidth-wasteful for prefill (8 x the ring / NVLS traffic:
        # measured 95 ms vs 62 ms at 8 GPUs, 2048 tokens).  Above `fused_max_rows` the layer uses the NCCL all-reduce.
        self.fused_max_rows = min(fused_max_rows, max_rows)
        self.bufs, self.hdls, self.dirty = [], [], [0, 0, 0]
        for _ in range(3):
            t = symm_mem.empty((max_rows, features), dtype=torch.bfloat16, device=dev)
            h = symm_mem.rendezvous(t, group.group_name)
            if not h.multicast_ptr:
                raise RuntimeError("FusedAllReducePool: this system exposes no NVLink multicast (NVLS) address")
            t.zero_()
            self.bufs.append(t)
            self.hdls.append(h)
        self.step = 0
        torch.cuda.synchronize()
        dist.barrier(group)

    def reset(self) -> None:
        """Start a new sequence of fused layers (call once per forward pass, e.g. at the top of a captured CUDA graph): zero
        whatever earlier steps left behind, restart the rotation at buffer 0 and meet the other ranks, so a replayed graph
        always begins from the same, clean state whatever its number of layers."""
        for i in range(3):
            if self.dirty[i]:
                self.bufs[i][: self.dirty[i]].zero_()
                self.dirty[i] = 0
        self.step = 0
        self.hdls[0].barrier(channel=1)

    def next(self, rows: int):
        assert rows <= self.max_rows
        k = self.step
        self.step += 1
        cur, nxt = k % 3, (k + 1) % 3
        if self.dirty[nxt]:
            self.bufs[nxt][: self.dirty[nxt]].zero_()
            self.dirty[nxt] = 0
        self.dirty[cur] = max(self.dirty[cur], rows)
        return self.bufs[cur][:rows], self.hdls[cur]


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

    def enable_fused_allreduce(self, pool: "FusedAllReducePool") -> None:
        assert pool.features == self.out_features
        self._fused_pool = pool

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        pool = getattr(self, "_fused_pool", None)
        if self.tp_world > 1 and pool is not None and x.numel() // x.shape[-1] <= pool.fused_max_rows:
            from .. import mx_gemm
            rows = x.numel() // x.shape[-1]
            view, hdl = pool.next(rows)
            target = mx_gemm.FusedOutput(view, hdl.multicast_ptr)
            y = super().forward(x, _fused=target)
            if target.taken:             # the GEMM took the buffer: its epilogue already reduced across ranks
                hdl.barrier(channel=0)   # every rank's adds have landed everywhere
                return y
            # the operands did not qualify for the fused epilogue: y is a plain local partial
        else:
            y = super().forward(x)
        if self.tp_world > 1:
            import torch.distributed as dist
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.tp_group)
        return y
