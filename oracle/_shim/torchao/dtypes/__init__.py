"""Import stub for torchmx/quant_api.py:8 (not on the hot path)."""


class AffineQuantizedTensor:  # pragma: no cover
    pass
