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
# channels_last activation [64, 32, 320, 320] (209.7 M elements) for the channel-innermost kernels
B = max(1, n // (32 * 320 * 320))
xa = torch.randn(B, 32, 320, 320, device="cuda").contiguous(memory_format=torch.channels_last)
ga = torch.randn(B, 32, 320, 320, device="cuda").contiguous(memory_format=torch.channels_last)
sc = torch.full((1, 32, 1, 1), 0.02, device="cuda")
zc = torch.full((1, 32, 1, 1), 3.3, device="cuda")
bias = torch.randn(32, device="cuda")
pc = ops.QSpec(0, 255, ch_axis=1, zp_learned=True, pre_relu=True)
xn = xa.contiguous()
gn = ga.contiguous()
# the weight bank over YOLOv8s-like weight shapes (57 tensors, ~11 M elements), per-channel W4 asymmetric LSQ
from vsiquantization_b200.bank import _Plan  # noqa: E402
import ctypes  # noqa: E402
from vsiquantization_b200 import _lib  # noqa: E402
shapes = [(32, 3, 3, 3), (64, 32, 3, 3)] + [(128, 64, 3, 3)] * 6 + [(256, 128, 3, 3)] * 12 + [(512, 256, 3, 3)] * 3 + \
    [(256, 256, 1, 1)] * 20 + [(128, 128, 3, 3)] * 14
items, grads = [], []
for shp in shapes:
    w = torch.randn(shp, device="cuda") * 0.1
    C = shp[0]
    items.append((None, w, torch.full((C,), 0.03, device="cuda", dtype=torch.float64), torch.full((C,), 7.3, device="cuda"),
                  ops.QSpec(0, 15, ch_axis=0, zp_learned=True), 2, ops.lsq_grad_scale(15, w.numel(), C), None))
    grads.append(torch.randn(shp, device="cuda"))
plan = _Plan(items, torch.device("cuda"))
y_flat = torch.empty(plan.total_out, device="cuda")
dx_flat = torch.empty(plan.total_out, device="cuda")
ds_flat = torch.empty(plan.total_q, dtype=torch.float64, device="cuda")
dz_flat = torch.empty(plan.total_q, device="cuda")
mt_ws = torch.empty(plan.ws_bytes, dtype=torch.uint8, device="cuda")
gptrs = (ctypes.c_void_p * len(items))(*[t.data_ptr() for t in grads])
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    ops.fake_quant_forward(x, 3.0 / 127, 0, spec)
    ops.fake_quant_backward_ste(x, g, 3.0 / 127, 0, spec)
    ops.lsq_backward(x, g, s_t, 0, spec, gs, ds_dtype=torch.float64)
    ops.fake_quant_forward_backward(x, g, 3.0 / 127, 0, spec)
    ops.observe(x)
    ops.ci_forward(xa, bias, sc, zc, pc)
    ops.ci_backward(xa, bias, ga, sc, zc, pc, 1e-3, None, True, True, True)
    ops.lsq_backward(xn, gn, sc, zc, pc, 1e-3, want_dz=True)
    ops.observe(xn, ch_axis=1)
    ops.observe(xa, ch_axis=1)
    _lib.check(_lib.lib.vsiq_mt_fake_quant_fwd(plan.host, plan.dev.data_ptr(), len(items), y_flat.data_ptr(), st))
    _lib.check(_lib.lib.vsiq_mt_lsq_bwd(plan.host, plan.dev.data_ptr(), len(items), gptrs, dx_flat.data_ptr(),
                                        ds_flat.data_ptr(), dz_flat.data_ptr(), mt_ws.data_ptr(), mt_ws.numel(), st))
torch.cuda.synchronize()
print("ok")
