"""Import stubs for torchmx/quant_api.py:9-12 (not on the hot path)."""


class LinearActivationQuantizedTensor:  # pragma: no cover
    pass


def to_linear_activation_quantized(weight, quant_fn):  # pragma: no cover
    raise NotImplementedError("torchao shim: not part of the oracle")
