"""Restatement of the two names the reference takes from torchao.utils
(torchmx/mx_tensor.py:20).  Test infrastructure only."""
import torch

TORCH_VERSION_AT_LEAST_2_4 = True
TORCH_VERSION_AT_LEAST_2_5 = True


class TorchAOBaseTensor(torch.Tensor):
    """Tensor base class with a per-class aten-override table.

    `implements(ops)` returns a decorator that files `fn` under each aten
    overload; `__torch_dispatch__` looks the overload up and calls
    `fn(func, types, args, kwargs)`.  That is the whole surface
    torchmx/ops.py:26 and torchmx/mx_tensor.py:357 rely on.
    """

    @classmethod
    def implements(cls, aten_ops):
        if "_ATEN_TABLE" not in cls.__dict__:
            cls._ATEN_TABLE = {}
        if not isinstance(aten_ops, (list, tuple)):
            aten_ops = [aten_ops]

        def deco(fn):
            for op in aten_ops:
                cls._ATEN_TABLE[op] = fn
            return fn

        return deco

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        kwargs = {} if kwargs is None else kwargs
        table = getattr(cls, "_ATEN_TABLE", {})
        if func in table:
            return table[func](func, types, args, kwargs)
        raise NotImplementedError(f"{cls.__name__} dispatch: no override for {func}")
