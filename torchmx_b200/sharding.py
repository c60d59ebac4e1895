"""Layer-sharded whole-model weight quantization across GPUs (BASELINE config 5).

`quantize_linear_` is a strictly sequential walk over independent layers (reference:
torchmx/quant_api.py:161-215), so the natural multi-GPU form is: one process per GPU, every rank
quantizes a contiguous range of the Linear layers, nothing is exchanged (no collective on the data
path).  This module only decides who owns what; torch.distributed is used by callers for the
barrier / timing reduction.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch


def linear_layer_names(model: torch.nn.Module) -> List[str]:
    """qualified names of the modules `quantize_linear_` would replace (exact nn.Linear type)."""
    return [name for name, mod in model.named_modules() if type(mod) is torch.nn.Linear]


def layer_shard(names: Sequence[str], rank: int, world_size: int, weights: Sequence[int] = None) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of `names` owned by `rank`.

    Without `weights` the split is by count (ranks differ by at most one layer).  With `weights`
    (e.g. parameter counts) the cut points are placed on the prefix sums so every rank gets about
    the same number of weight elements -- lm_head (128256 x hidden) would otherwise unbalance the
    last rank of a Llama.
    """
    n = len(names)
    assert 0 <= rank < world_size
    if weights is None:
        lo = n * rank // world_size
        hi = n * (rank + 1) // world_size
        return lo, hi
    assert len(weights) == n
    total = sum(weights)
    prefix = [0]
    for w in weights:
        prefix.append(prefix[-1] + w)

    def cut(r: int) -> int:
        if r <= 0:
            return 0
        if r >= world_size:
            return n
        target = total * r / world_size
        # first index whose prefix sum reaches the target
        i = min(range(n + 1), key=lambda j: abs(prefix[j] - target))
        return i

    return cut(rank), cut(rank + 1)


def shard_filter(model: torch.nn.Module, rank: int, world_size: int, balance_by_numel: bool = True) -> Callable[[str], bool]:
    """`layer_filter` for quantize_linear_: True for the Linear layers this rank owns."""
    names = linear_layer_names(model)
    weights = None
    if balance_by_numel:
        mods = dict(model.named_modules())
        weights = [mods[n].weight.numel() for n in names]
    lo, hi = layer_shard(names, rank, world_size, weights)
    mine = set(names[lo:hi])
    return lambda fqn: fqn in mine
