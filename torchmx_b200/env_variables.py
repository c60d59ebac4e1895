"""Environment toggles, same names and defaults as /root/reference/torchmx/env_variables.py:1-16.

`MX_EXACT_QUANTIZATION` is read once at import into a module attribute and compared to the
*string* "True" at every quantize call (reference: torchmx/mx_tensor.py:80-83; its tests flip
the attribute at run time, tests/conftest.py:66-69), so callers may mutate it after import.
"""
import os

TORCHMX_LOG_LEVEL = os.getenv("LOG_LEVEL", "INFO")
TORCHMX_LOG_FILE = os.getenv("LOG_FILE", None)
MX_EXACT_QUANTIZATION = os.getenv("MX_HARDWARE_EXACT_QUANTIZATION", "False")
