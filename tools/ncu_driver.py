"""Launches each hot kernel a few times at one size -- the command ncu profiles (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import ops  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1 << lg
torch.manual_seed(0)
x = torch.randn(n, device="cuda")
g = torch.randn(n, device="cuda")
s_t = torch.tensor(3.0 / 127, dtype=torch.float64, device="cuda")
spec = ops.QSpec(-128, 127)
gs = ops.lsq_grad_scale(127, n)
for _ in range(reps):
    ops.fake_quant_forward(x, 3.0 / 127, 0, spec)
    ops.fake_quant_backward_ste(x, g, 3.0 / 127, 0, spec)
    ops.lsq_backward(x, g, s_t, 0, spec, gs, ds_dtype=torch.float64)
    ops.fake_quant_forward_backward(x, g, 3.0 / 127, 0, spec)
    ops.observe(x)
torch.cuda.synchronize()
print("ok")
