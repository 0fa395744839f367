"""NHWC kernels on YOLOv8s-sized feature maps, a few launches each -- the command ncu profiles for the small-tensor regime.
    python tools/ci_ncu_driver.py [C H [reps]]        default: 256 40 (26 M elements at batch 64)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import ops  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 40
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.cuda.set_stream(torch.cuda.Stream())
x = torch.randn(64, C, H, H, device="cuda").contiguous(memory_format=torch.channels_last)
g = torch.randn(64, C, H, H, device="cuda").contiguous(memory_format=torch.channels_last)
wide = torch.randn(64, 2 * C, H, H, device="cuda").contiguous(memory_format=torch.channels_last)
gp = wide[:, C:]  # pitched grad_output (what torch.cat's backward hands out)
y = torch.empty_like(x)
b = torch.randn(C, device="cuda")
sc = torch.full((1, C, 1, 1), 0.02, device="cuda")
zc = torch.full((1, C, 1, 1), 3.3, device="cuda")
pc = ops.QSpec(0, 255, ch_axis=1, zp_learned=True, pre_relu=True)
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
st = ops.new_observer_state(1, x.device)
for _ in range(reps):
    flush.fill_(1.0)
    y.copy_(x)
    flush.fill_(1.0)
    ops.ci_forward(x, b, sc, zc, pc, out=y)
    flush.fill_(1.0)
    ops.ci_backward(x, b, g, sc, zc, pc, 1e-3, None, True, True, True)
    flush.fill_(1.0)
    ops.ci_backward(x, b, gp, sc, zc, pc, 1e-3, None, True, True, True)
    flush.fill_(1.0)
    ops.observe(x, ch_axis=1)
    flush.fill_(1.0)
    ops.ci_forward(x, b, sc, zc, pc, out=y, second=(0.02, 0, ops.QSpec(-8, 7)))   # y + the next layer's quantize_inp
    flush.fill_(1.0)
    ops.ci_epilogue_observe(x, st, 8, True, 1e-8, "relu", bias=b)                  # calibration epilogue
torch.cuda.synchronize()
print("ok")
