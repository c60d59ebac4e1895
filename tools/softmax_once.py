"""One K4a launch per mask mode on [1, 32, 2048, 2048] bf16 scores (for ncu): causal rule, explicit additive mask, no mask."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchmx_b200  # noqa
from torchmx_b200 import attention_ops, dtypes
H, S = 32, 2048
scores = (torch.randn(1, H, S, S, device="cuda") * 11).to(torch.bfloat16)
hidden = torch.ones(S, S, dtype=torch.bool, device="cuda").triu_(1)
addm = torch.zeros(1, 1, S, S, device="cuda", dtype=torch.bfloat16).masked_fill_(hidden, float("-inf"))
for _ in range(2):
    attention_ops.softmax_to_mx(scores, 128 ** -0.5, None, True, dtypes.float8_e4m3, 32)
    attention_ops.softmax_to_mx(scores, 128 ** -0.5, addm, False, dtypes.float8_e4m3, 32)
    attention_ops.softmax_to_mx(scores, 128 ** -0.5, None, False, dtypes.float8_e4m3, 32)
torch.cuda.synchronize()
print("ok")
