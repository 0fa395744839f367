"""Channel-innermost kernels on YOLOv8s activation shapes (channels_last), vs the flat per-tensor kernels and a copy."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import ops
from tools.microbench import timed
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
if os.environ.get("BENCH_SIDE_STREAM", "1") != "0":  # programmatic dependent launch needs a non-default stream
    torch.cuda.set_stream(torch.cuda.Stream())
shapes = [(B, 32, 320, 320), (B, 64, 160, 160), (B, 128, 80, 80), (B, 256, 40, 40), (B, 512, 20, 20), (B, 64, 80, 80), (B, 128, 20, 20)]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for shp in shapes:
    C = shp[1]
    x = torch.randn(shp, device="cuda").contiguous(memory_format=torch.channels_last)
    g = torch.randn(shp, device="cuda").contiguous(memory_format=torch.channels_last)
    y = torch.empty_like(x); n = x.numel()
    gw = torch.randn(shp[0], 2 * C, shp[2], shp[3], device="cuda").contiguous(memory_format=torch.channels_last)[:, C:]  # pitched g
    fl = flush if n * 4 < (256 << 20) else None
    s_t = torch.tensor(0.02, device="cuda"); b = torch.randn(C, device="cuda")
    sc = torch.full((1, C, 1, 1), 0.02, device="cuda"); zc = torch.full((1, C, 1, 1), 3.3, device="cuda")
    pt = ops.QSpec(-128, 127, pre_relu=True); pc = ops.QSpec(0, 255, ch_axis=1, zp_learned=True, pre_relu=True)
    cases = [
        ("copy", 8, lambda: y.copy_(x)),
        ("flat fwd pt", 8, lambda: ops.fake_quant_forward(x, s_t, 0, pt, out=y)),
        ("ci fwd pt+b", 8, lambda: ops.ci_forward(x, b, s_t, 0, pt, out=y)),
        ("ci fwd pc+b", 8, lambda: ops.ci_forward(x, b, sc, zc, pc, out=y)),
        ("ci fwd2 (y + next layer's quantize_inp)", 12, lambda: ops.ci_forward(x, b, sc, zc, pc, out=y, second=(s_t, 0, ops.QSpec(-8, 7)))),
        ("flat lsq pt", 12, lambda: ops.lsq_backward(x, g, s_t, 0, pt, 1e-3)),
        ("ci lsq pt+b", 12, lambda: ops.ci_backward(x, b, g, s_t, 0, pt, 1e-3)),
        ("ci lsq pc+b", 12, lambda: ops.ci_backward(x, b, g, sc, zc, pc, 1e-3, None, True, True, True)),
        ("ci lsq pitched", 12, lambda: ops.ci_backward(x, b, gw, sc, zc, pc, 1e-3, None, True, True, True)),
        ("ci observe", 4, lambda: ops.observe(x, ch_axis=1)),
        ("sum(0,2,3)", 4, lambda: g.sum((0, 2, 3))),
    ]
    line = f"{str(shp):22s} {n/1e6:6.1f}M "
    for name, bpe, fn in cases:
        med, best = timed(fn, 10, fl)
        line += f"| {name} {med*1e3:6.1f}us {bpe*n/med/1e6:5.0f} "
    print(line)
