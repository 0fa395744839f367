"""Fake-quant kernels on the activation shapes of YOLOv8s at batch 64 @640 (per tensor and per channel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import ops
from tools.microbench import timed

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shapes = [(B, 32, 320, 320), (B, 64, 160, 160), (B, 128, 80, 80), (B, 256, 40, 40), (B, 512, 20, 20), (B, 64, 80, 80), (B, 128, 20, 20)]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
tot = {}
for shp in shapes:
    x = torch.randn(shp, device="cuda"); g = torch.randn(shp, device="cuda")
    n = x.numel(); C = shp[1]
    fl = flush if n * 4 < (256 << 20) else None
    s_t = torch.tensor(0.02, dtype=torch.float64, device="cuda")
    sc = torch.full((1, C, 1, 1), 0.02, device="cuda"); zc = torch.full((1, C, 1, 1), 3.3, device="cuda")
    y = torch.empty_like(x); dx = torch.empty_like(x)
    cases = [
        ("copy", 8, lambda: y.copy_(x)),
        ("fwd pt relu", 8, lambda: ops.fake_quant_forward(x, s_t, 0, ops.QSpec(-128, 127, pre_relu=True), out=y)),
        ("lsq pt relu", 12, lambda: ops.lsq_backward(x, g, s_t, 0, ops.QSpec(-128, 127, pre_relu=True), 1e-3, ds_dtype=torch.float64, dx_out=dx)),
        ("fwd pc relu", 8, lambda: ops.fake_quant_forward(x, sc, zc, ops.QSpec(0, 255, ch_axis=1, zp_learned=True, pre_relu=True), out=y)),
        ("lsq pc relu", 12, lambda: ops.lsq_backward(x, g, sc, zc, ops.QSpec(0, 255, ch_axis=1, zp_learned=True, pre_relu=True), 1e-3, want_dz=True, dx_out=dx)),
        ("observe pt", 4, lambda: ops.observe(x)),
        ("observe pc", 4, lambda: ops.observe(x, ch_axis=1)),
    ]
    line = f"{str(shp):22s} {n/1e6:7.1f}M "
    for name, bpe, fn in cases:
        med, best = timed(fn, 10, fl)
        line += f"| {name} {med*1e3:7.1f}us {bpe*n/med/1e6:6.0f} "
        tot[name] = tot.get(name, 0) + med
    print(line)
print("sum ms:", {k: round(v, 3) for k, v in tot.items()})
