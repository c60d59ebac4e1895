"""Minimal stand-in for torchao 0.6.1 (pinned by the reference: pyproject.toml:23).

TEST INFRASTRUCTURE ONLY.  torchao is not installed in this image and cannot be
downloaded; the reference imports a base tensor class and three fp4/fp6 cast
helpers from it (torchmx/mx_tensor.py:20, torchmx/mx_quantization_utils.py:4-8).
This shim restates just those, so that the *unmodified* reference under
/root/reference can be imported on CPU by oracle/gen_golden.py.  Nothing in the
product package (torchmx_b200/) imports it.
"""
__version__ = "0.6.1"
