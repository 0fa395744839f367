"""Randomised oracle-vs-LIVE-reference check (build container only).

TEST INFRASTRUCTURE.  The golden vectors under tests/golden pin the oracle on a fixed set of cases; this script pins it
on hundreds of fresh ones by running the UNMODIFIED reference (imported through oracle/ref_shim.py, torch on CPU) beside
the C oracle on the same seeded inputs:

  * UniformQuantizer.quantize, fixed qparams (quantizers/uniform.py:34-56,81-96): y bit for bit, any bit width 2..8,
    symmetric / asymmetric, arbitrary integer zero-point, inputs salted with ties (k + 1/2) * s, +-0, denormals, +-inf, NaN;
  * the learnable-scale path (uniform.py:47-52,242-271; 0-dim float64 Parameter, quantization_manager.py:99):
    y and dx bit for bit, dscale within 1e-5 of the absolute mass of its terms (the reference's own fp32 sums cancel);
  * the same forward with degenerate scales (0, negative, 1e-30, 1e30, inf, NaN, 2^-41, 2^41);
  * the construction-time BN fold (modules/fused.py:92-108): bit for bit except where torch-on-CPU's vectorised sqrt
    (Sleef) is itself one ulp off the correctly rounded value -- there W' may differ by up to three ulps;
  * MinMaxObserver traces (observers/minmax.py:32-74): running min / max, scale and zero-point exactly (Python doubles);
  * LSQFakeQuantize per-channel learn phase (quantizers/lsq_module.py:147-173,254-274,317-358): y, dx bit for bit, per-channel
    dscale / dzero_point within 1e-5 of their mass.

Usage: python oracle/check_live_reference.py [--cases N] [--seed S]     (exit code 0 = all equal)
/root/reference does not exist on the GPU box: nothing that runs there may import this file."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import ref_shim  # noqa: E402


def bits_equal(a, b) -> bool:
    a, b = np.ascontiguousarray(a, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and bool(np.array_equal(na, nb)) and bool(
        np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32)))


def salted(rng, shape, s):
    x = (rng.standard_normal(shape) * rng.choice([0.3, 1.0, 4.0, 40.0])).astype(np.float32)
    flat = x.reshape(-1)
    k = rng.integers(-140, 140, size=max(1, flat.size // 8))
    pos = rng.choice(flat.size, size=k.size, replace=False)
    flat[pos] = ((k + 0.5) * np.float64(s)).astype(np.float32)          # rounding ties
    specials = np.array([0.0, -0.0, 1e-42, -1e-42, np.inf, -np.inf, np.nan, 3.0e38, -3.0e38], dtype=np.float32)
    pos = rng.choice(flat.size, size=min(flat.size, specials.size), replace=False)
    flat[pos] = specials[:pos.size]
    return x


def mass(x, g, s, z, qmin, qmax):
    with np.errstate(all="ignore"):
        v = x.astype(np.float32) / np.float32(s)
        q = np.clip(np.rint(v + np.float32(z)), qmin, qmax)
        t = np.abs(g.astype(np.float64) * (q - z)) + np.abs(g.astype(np.float64) * v)
    return float(np.nansum(t[np.isfinite(t)]))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    if not ref_shim.available():
        print("reference tree not present")
        return 2
    ref_shim.install()
    import torch
    from quantizers.uniform import UniformQuantizer as RefQ
    from observers.minmax import MinMaxObserver as RefObs
    from quantizers.lsq_module import LSQFakeQuantize
    from torch.quantization import MovingAveragePerChannelMinMaxObserver
    rng = np.random.default_rng(args.seed)
    bad = []

    for case in range(args.cases):
        bits = int(rng.integers(2, 9))
        sym = bool(rng.integers(0, 2))
        shape = tuple(int(d) for d in rng.integers(1, 9, size=int(rng.integers(1, 5))))
        s = float(rng.choice([0.004, 0.0236, 0.11, 0.43, 1.0, 7.5]) * rng.uniform(0.5, 2.0))
        q = RefQ(bits, sym)
        zp = 0 if sym else int(rng.integers(q.qmin, q.qmax + 1))
        x = salted(rng, shape, s)
        # ---- fixed qparams: forward
        with torch.no_grad():
            y_ref = q.quantize(torch.from_numpy(x.copy()), s, zp, False).numpy()
        y = oracle.fake_quant_fwd(x, s, zp, q.qmin, q.qmax)
        if not bits_equal(y, y_ref):
            bad.append(("fixed fwd", case, bits, sym, shape, s, zp))
        # ---- learnable scale (symmetric: the only learning flow the reference can execute)
        if sym:
            xf = np.where(np.isfinite(x), x, np.float32(0.25)).astype(np.float32)  # autograd of inf / NaN is NaN soup
            g = rng.standard_normal(shape).astype(np.float32)
            xt = torch.from_numpy(xf.copy()).requires_grad_(True)
            sp = torch.nn.Parameter(torch.tensor(np.float64(s)))
            yl = q.quantize(xt, sp, 0, True)
            yl.backward(torch.from_numpy(g))
            gs = oracle.grad_scale(q.qmax, xf.size)
            yo = oracle.fake_quant_fwd(xf, s, 0, q.qmin, q.qmax)
            dx, ds, _ = oracle.fake_quant_bwd(xf, g, s, 0, q.qmin, q.qmax, grad_scale=gs)
            if not bits_equal(yo, yl.detach().numpy()) or not bits_equal(dx, xt.grad.numpy()):
                bad.append(("learned y/dx", case, bits, shape, s))
            if abs(ds[0] - float(sp.grad)) > 1e-5 * gs * mass(xf, g, s, 0, q.qmin, q.qmax) + 1e-30:
                bad.append(("learned ds", case, bits, shape, s, ds[0], float(sp.grad)))
        # ---- observer trace
        obs = RefObs(sym)
        run_min = run_max = 0.0
        for _ in range(3):
            xb = (rng.standard_normal(shape) * rng.uniform(0.1, 5.0) + rng.uniform(-1, 1)).astype(np.float32)
            sc_ref, zp_ref = obs.forward(torch.from_numpy(xb))
            st = oracle.minmax_stats(xb)
            run_min, run_max = oracle.minmax_update(run_min, run_max, st[0, 0], st[0, 1])
            sc, zz = oracle.qparams(run_min, run_max, 8, sym)
            if (run_min, run_max, sc, zz) != (obs.min_val, obs.max_val, sc_ref, zp_ref):
                bad.append(("observer", case, sym, (run_min, run_max, sc, zz), (obs.min_val, obs.max_val, sc_ref, zp_ref)))
                break

    # ---- degenerate scales (zero, negative, tiny, huge, inf, NaN): the reference divides by whatever it is given
    for sdeg in (0.0, -0.05, 1e-30, 1e30, float("inf"), float("nan"), 2.0 ** -41, 2.0 ** 41):
        for sym in (True, False):
            q = RefQ(8, sym)
            zp = 0 if sym else 77
            x = salted(rng, (5, 7), 0.1)
            with torch.no_grad(), np.errstate(all="ignore"):
                y_ref = q.quantize(torch.from_numpy(x.copy()), sdeg, zp, False).numpy()
            if not bits_equal(oracle.fake_quant_fwd(x, sdeg, zp, q.qmin, q.qmax), y_ref):
                bad.append(("degenerate scale", sdeg, sym))

    # ---- LSQFakeQuantize, per channel, learn phase (asymmetric quint8 activations and symmetric qint8 weights)
    for case in range(max(4, args.cases // 6)):
        affine = bool(case % 2)
        C = int(rng.integers(2, 7))
        shape = (int(rng.integers(1, 4)), C, int(rng.integers(1, 6)), int(rng.integers(1, 6)))
        kw = dict(observer=MovingAveragePerChannelMinMaxObserver, ch_axis=1,
                  quant_min=0 if affine else -128, quant_max=255 if affine else 127,
                  dtype=torch.quint8 if affine else torch.qint8,
                  qscheme=torch.per_channel_affine if affine else torch.per_channel_symmetric)
        fq = LSQFakeQuantize(learn_scale=True, config_act=affine, **kw)
        x0 = (rng.standard_normal(shape) * 2 + (1.0 if affine else 0.0)).astype(np.float32)
        fq(torch.from_numpy(x0))            # observer phase: initialises scale_param / zero_point_param_float
        fq.disable_observer()
        x = (rng.standard_normal(shape) * 2.5 + (1.0 if affine else 0.0)).astype(np.float32)
        g = rng.standard_normal(shape).astype(np.float32)
        xt = torch.from_numpy(x.copy()).requires_grad_(True)
        y = fq(xt)
        y.backward(torch.from_numpy(g))
        s = fq.scale_param.detach().numpy().reshape(-1).astype(np.float64)
        zf = fq.zero_point_param_float.detach().numpy().reshape(-1).astype(np.float64)
        qmin, qmax = kw["quant_min"], kw["quant_max"]
        gs = oracle.grad_scale(qmax, x.size, C) * (5000.0 if affine else 1.0)
        yo = oracle.fake_quant_fwd(x, s, zf, qmin, qmax, ch_axis=1, zp_learned=True)
        dx, ds, dz = oracle.fake_quant_bwd(x, g, s, zf, qmin, qmax, ch_axis=1, zp_learned=True, grad_scale=gs,
                                           want_ds=True, want_dz=True)
        if not bits_equal(yo, y.detach().numpy()) or not bits_equal(dx, xt.grad.numpy()):
            bad.append(("lsq per-channel y/dx", case, shape, affine))
        ds_ref = fq.scale_param.grad.numpy().reshape(-1)
        dz_ref = fq.zero_point_param_float.grad.numpy().reshape(-1)
        for c in range(C):
            zc = float(np.clip(np.rint(np.float32(zf[c])), qmin, qmax))
            m = gs * mass(x[:, c], g[:, c], s[c], zc, qmin, qmax)
            zm = gs * float(np.sum(np.abs(g[:, c].astype(np.float64) * np.float32(s[c]))))
            if abs(ds[c] - ds_ref[c]) > 1e-5 * m + 1e-30 or abs(dz[c] - dz_ref[c]) > 1e-5 * zm + 1e-30:
                bad.append(("lsq per-channel ds/dz", case, c, ds[c], ds_ref[c], dz[c], dz_ref[c]))
    # ---- BN fold (modules/fused.py:92-108): the oracle is the IEEE op sequence; torch-on-CPU's vectorised sqrt (Sleef) is
    # one ulp off for a fraction of a percent of inputs, so the comparison is "within one ulp", and mostly bit for bit
    from modules.fused import ConvBnReLU
    ulp_off = total = 0
    for case in range(max(3, args.cases // 10)):
        cin, cout, k = int(rng.integers(1, 9)), int(rng.integers(1, 40)), int(rng.choice([1, 3]))
        cv = torch.nn.Conv2d(cin, cout, k, bias=bool(case % 2))
        bn = torch.nn.BatchNorm2d(cout, eps=0.001)
        with torch.no_grad():
            bn.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32)))
            bn.bias.copy_(torch.from_numpy(rng.standard_normal(cout).astype(np.float32) * 0.1))
            bn.running_mean.copy_(torch.from_numpy(rng.standard_normal(cout).astype(np.float32) * 0.1))
            bn.running_var.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32)))
        layer = ConvBnReLU(cv, bn, torch.nn.ReLU(), "MinMaxObserver", "UniformQuantizer", "MinMaxObserver",
                           "UniformQuantizer", True, True, True, 8, 8)
        Wf, bf = oracle.bn_fold(cv.weight.detach().numpy(), cv.bias.detach().numpy() if cv.bias is not None else None,
                                bn.weight.detach().numpy(), bn.bias.detach().numpy(), bn.running_mean.numpy(),
                                bn.running_var.numpy(), bn.eps)
        # a one-ulp sqrt moves t = gamma / std by at most one ulp, W * t by at most two more after rounding
        Wr, br = layer.conv_fuse.weight.detach().numpy(), layer.conv_fuse.bias.detach().numpy()
        d = np.abs(Wf.view(np.int32).astype(np.int64) - Wr.view(np.int32).astype(np.int64))
        total += d.size + br.size
        ulp_off += int((d > 0).sum()) + int((bf != br).sum())
        if d.max() > 3:
            bad.append(("bn fold: W' more than three ulps from the reference", case, int(d.max())))
        t_mag = np.abs(bn.weight.detach().numpy()) / np.sqrt(bn.running_var.numpy() + np.float32(bn.eps))
        b0 = np.abs(cv.bias.detach().numpy()) if cv.bias is not None else 0.0
        mass_b = np.abs(bn.bias.detach().numpy()) + (b0 + np.abs(bn.running_mean.numpy())) * t_mag
        if np.any(np.abs(bf.astype(np.float64) - br) > 4e-7 * mass_b):
            bad.append(("bn fold: b' off", case, float(np.abs(bf - br).max())))
    if total and ulp_off > 0.05 * total:
        bad.append(("bn fold: too many one-ulp differences", ulp_off, total))
    for b in bad[:20]:
        print("MISMATCH", b)
    print(f"bn fold: {ulp_off} of {total} values differ from torch-on-CPU by its Sleef sqrt (<= 3 ulps), the rest bit for bit")
    print(f"{args.cases} uniform/observer cases, {max(4, args.cases // 6)} LSQFakeQuantize cases, {len(bad)} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
