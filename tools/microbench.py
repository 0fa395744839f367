"""Kernel-level microbenchmark (CUDA events on the launching stream, inputs larger than L2 or L2 flushed).

    python tools/microbench.py [--log2n 26 28] [--iters 20]

Prints one line per kernel and size: time, algorithmic GB/s, fraction of the measured copy bandwidth
taken in the same run (and of MEASURED_PEAKS.json when present).  Not the contract bench (bench.py is).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vsiquantization_b200 import ops  # noqa: E402


def timed(fn, iters, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, nargs="+", default=[24, 26, 28])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--per-channel", action="store_true")
    args = ap.parse_args()
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    peak = peaks.get("hbm_gbs", 6650.0)
    dev = torch.device("cuda")
    print("device", torch.cuda.get_device_name(0), "peak(hbm_gbs)", peak)
    flush = torch.empty(256 << 20, dtype=torch.float32, device=dev)  # 1 GiB > 126 MB L2
    for lg in args.log2n:
        n = 1 << lg
        torch.manual_seed(0)
        x = torch.randn(n, device=dev)
        g = torch.randn(n, device=dev)
        y = torch.empty_like(x)
        fl = flush if n * 4 <= (256 << 20) else None
        med, best = timed(lambda: y.copy_(x), args.iters, fl)
        copy_gbs = 8 * n / med / 1e6
        print(f"n=2^{lg} torch copy_           {med:8.3f} ms  {copy_gbs:8.1f} GB/s (best {8 * n / best / 1e6:.1f})")
        s_t = torch.tensor(3.0 / 127, dtype=torch.float64, device=dev)
        spec = ops.QSpec(-128, 127)
        gs = ops.lsq_grad_scale(127, n)
        cases = [
            ("fq_fwd   W8 host-scale", 8, lambda: ops.fake_quant_forward(x, 3.0 / 127, 0, spec)),
            ("fq_fwd   W8 dev-scale ", 8, lambda: ops.fake_quant_forward(x, s_t, 0, spec)),
            ("fq_fwd   W4 asym      ", 8, lambda: ops.fake_quant_forward(x, 3.0 / 7, 8, ops.QSpec(0, 15))),
            ("ste_bwd  W8           ", 12, lambda: ops.fake_quant_backward_ste(x, g, 3.0 / 127, 0, spec)),
            ("lsq_bwd  W8           ", 12, lambda: ops.lsq_backward(x, g, s_t, 0, spec, gs, ds_dtype=torch.float64)),
            ("fwd+bwd fused W8      ", 16, lambda: ops.fake_quant_forward_backward(x, g, 3.0 / 127, 0, spec)),
            ("observe (minmax+stats)", 4, lambda: ops.observe(x)),
        ]
        if args.per_channel:
            for C in (64, 512, 4096):
                xc = x.view(C, n // C)
                gc = g.view(C, n // C)
                sc = (torch.rand(C, device=dev) * 1.5 + 0.5) * (3.0 / 127)
                zc = torch.zeros(C, device=dev)
                spc = ops.QSpec(-128, 127, ch_axis=0)
                gsc = ops.lsq_grad_scale(127, n, C)
                cases += [
                    (f"fq_fwd  per-ch C={C:<5d} ", 8, lambda xc=xc, sc=sc, zc=zc, spc=spc: ops.fake_quant_forward(xc, sc, zc, spc)),
                    (f"lsq_bwd per-ch C={C:<5d} ", 12, lambda xc=xc, gc=gc, sc=sc, zc=zc, spc=spc, gsc=gsc:
                     ops.lsq_backward(xc, gc, sc, zc, spc, gsc, want_dz=True)),
                ]
            if n >= 64 * 320 * 320:
                B = n // (64 * 320 * 320)
                if B >= 1:
                    m = B * 64 * 320 * 320
                    xa = x[:m].view(B, 64, 320, 320)
                    ga = g[:m].view(B, 64, 320, 320)
                    sa = (torch.rand(64, device=dev) * 1.5 + 0.5) * (3.0 / 127)
                    za = torch.full((64,), 3.3, device=dev)
                    spa = ops.QSpec(0, 255, ch_axis=1, zp_learned=True)
                    gsa = ops.lsq_grad_scale(255, m, 64)
                    sf = m / n
                    cases += [
                        (f"fq_fwd  NCHW B={B} ch1   ", 8 * sf, lambda: ops.fake_quant_forward(xa, sa, za, spa)),
                        (f"lsq_bwd NCHW B={B} ch1   ", 12 * sf, lambda: ops.lsq_backward(xa, ga, sa, za, spa, gsa, want_dz=True)),
                        (f"observe NCHW B={B} ch1   ", 4 * sf, lambda: ops.observe(xa, ch_axis=1)),
                    ]
        for name, bpe, fn in cases:
            med, best = timed(fn, args.iters, fl)
            gbs = bpe * n / med / 1e6
            print(f"n=2^{lg} {name} {med:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.3f} of measured peak, "
                  f"{gbs / copy_gbs:5.3f} of in-run copy (best {bpe * n / best / 1e6:.1f})")
        del x, g, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
