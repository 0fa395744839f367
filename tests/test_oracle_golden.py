"""The CPU oracle (oracle/vsiq_oracle.c) against the golden vectors produced by the reference itself.

This is what pins the oracle: every fixture under tests/golden/ was written by oracle/gen_golden.py
running the reference's own Python classes on CPU.  Bit-exact for fake-quantised values, codes, dx,
observer min/max and scales; stated tolerances for the order-dependent sums (ds, dz, BN moments).
"""
import os

import numpy as np
import pytest

import oracle
from conftest import bits_equal, first_mismatch, load_golden


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


# The reference's fp32 sums carry their own rounding error (cancellation between sum g*q and
# sum g*x/s, SURVEY.md 7 "ds tolerance"), so ds/dz are compared normalised by the absolute mass
# of the summed terms, which is what a 1e-5 "relative (fp32)" bound can mean for a cancelling sum.
def sum_mass(x, g, s, z, qmin, qmax):
    s32 = np.float32(s)
    v = x.astype(np.float32) / s32
    q = np.clip(np.rint(v + np.float32(z)), qmin, qmax)
    return float(np.sum(np.abs(g.astype(np.float64) * (q - z))) + np.sum(np.abs(g.astype(np.float64) * v)))


def test_uniform_fixed_bit_exact():
    G = load_golden("uniform_fixed")
    for tag in G["cases"]:
        scale, zp, qmin, qmax, bits, sym = G[f"{tag}_qp"]
        x, g = G[f"{tag}_x"], G[f"{tag}_g"]
        y, codes = oracle.fake_quant_fwd(x, scale, zp, int(qmin), int(qmax), want_codes=True)
        assert bits_equal(y, G[f"{tag}_y"]), (tag, first_mismatch(y, G[f"{tag}_y"]))
        assert bits_equal(codes, G[f"{tag}_codes"]), (tag, first_mismatch(codes, G[f"{tag}_codes"]))
        dx, _, _ = fake_bwd = oracle.fake_quant_bwd(x, g, scale, zp, int(qmin), int(qmax), want_ds=False)
        assert bits_equal(dx, G[f"{tag}_dx"]), (tag, first_mismatch(dx, G[f"{tag}_dx"]))
        # codes are integers inside [qmin, qmax] wherever x is not NaN
        ok = ~np.isnan(codes)
        assert np.all(codes[ok] == np.rint(codes[ok])) and codes[ok].min() >= qmin and codes[ok].max() <= qmax


def test_uniform_learned():
    G = load_golden("uniform_learned")
    for tag in G["cases"]:
        scale, zf, qmin, qmax, bits, sym, gs = G[f"{tag}_qp"]
        sym = bool(sym)
        x, g = G[f"{tag}_x"], G[f"{tag}_g"]
        y = oracle.fake_quant_fwd(x, scale, zf, int(qmin), int(qmax), zp_learned=not sym)
        assert bits_equal(y, G[f"{tag}_y"]), (tag, first_mismatch(y, G[f"{tag}_y"]))
        dx, ds, dz = oracle.fake_quant_bwd(x, g, scale, zf, int(qmin), int(qmax), zp_learned=not sym,
                                           grad_scale=gs, want_ds=True, want_dz=not sym)
        assert bits_equal(dx, G[f"{tag}_dx"]), (tag, first_mismatch(dx, G[f"{tag}_dx"]))
        assert gs == pytest.approx(oracle.grad_scale(int(qmax), x.size), rel=1e-15)
        zeff = float(np.clip(np.rint(np.float32(zf)), qmin, qmax)) if not sym else 0.0
        mass = gs * sum_mass(x, g, scale, zeff, qmin, qmax)
        assert abs(ds[0] - G[f"{tag}_ds"][0]) <= 1e-5 * mass, (tag, ds, G[f"{tag}_ds"], mass)
        if not sym:
            zmass = gs * float(np.sum(np.abs(g.astype(np.float64) * np.float32(scale))))
            assert abs(dz[0] - G[f"{tag}_dz"][0]) <= 1e-5 * zmass, (tag, dz, G[f"{tag}_dz"])


def test_calib_grad_scale_vector():
    """calib_grad_scale as a [C] tensor collapses to its sum on a 0-dim scale."""
    G = load_golden("uniform_learned")
    x, g = G["cgs_x"], G["cgs_g"]
    gs = oracle.grad_scale(127, x.size) * float(G["cgs_vec"].astype(np.float64).sum())
    y = oracle.fake_quant_fwd(x, 0.02, 0, -128, 127)
    assert bits_equal(y, G["cgs_y"])
    dx, ds, _ = oracle.fake_quant_bwd(x, g, 0.02, 0, -128, 127, grad_scale=gs)
    assert bits_equal(dx, G["cgs_dx"])
    mass = gs * sum_mass(x, g, 0.02, 0, -128, 127)
    assert abs(ds[0] - G["cgs_ds"][0]) <= 1e-5 * mass


def test_funlsq_mask_mode():
    G = load_golden("funlsq")
    s, z, qmin, qmax, gsc = G["qp"]
    x, g = G["x"], G["g"]
    # FunLSQ forward: round(w/s).clamp * s == the a5 forward with z = 0 except for the sign of zero
    y = oracle.fake_quant_fwd(x, s, 0, int(qmin), int(qmax))
    assert np.array_equal(y, G["y"])
    dx, ds, _ = oracle.fake_quant_bwd(x, g, s, 0, int(qmin), int(qmax), grad_scale=gsc, mask_mode=1)
    assert bits_equal(dx, G["dx"]), first_mismatch(dx, G["dx"])
    mass = gsc * float(np.sum(np.abs(g.astype(np.float64))) * 0.5 + 1.0)
    assert abs(ds[0] - G["ds"][0]) <= 1e-5 * mass


def test_observer_traces_bit_exact():
    G = load_golden("observer")
    for tag in G["cases"]:
        name, symtag, btag = tag.rsplit("_", 2)
        sym, bits = symtag == "sym", int(btag[1:])
        n = int(G[f"{name}_n"])
        run_min, run_max = 0.0, 0.0  # observers/minmax.py:28-29
        trace = G[f"{tag}_trace"]
        for i in range(n):
            st = oracle.minmax_stats(G[f"{name}_in{i}"])
            run_min, run_max = oracle.minmax_update(run_min, run_max, st[0, 0], st[0, 1])
            s, z = oracle.qparams(run_min, run_max, bits, sym)
            exp = trace[i]
            assert run_min == exp[0] and run_max == exp[1], (tag, i, run_min, run_max, exp)
            assert s == exp[2] or (np.isnan(s) and np.isnan(exp[2])), (tag, i, s, exp[2])
            assert z == exp[3] or (np.isnan(z) and np.isnan(exp[3])), (tag, i, z, exp[3])


def test_manager_flow():
    G = load_golden("manager")
    for tag in G["cases"]:
        bits = int(tag[1])
        sym = tag.endswith("_sym")
        assert int(G[f"{tag}_observer_bits"]) == 8  # quantization_manager.py:42 never forwards bits_width
        run_min = run_max = 0.0
        mean_abs = []
        for i in range(3):
            x = G[f"{tag}_in{i}"]
            st = oracle.minmax_stats(x)
            run_min, run_max = oracle.minmax_update(run_min, run_max, st[0, 0], st[0, 1])
            mean_abs.append(st[0, 2] / x.size)
            n = x.size
            mean = st[0, 3] / n
            var = (st[0, 4] - n * mean * mean) / (n - 1)
            # the reference's stats are torch fp32 reductions: tolerance, not bit-exact
            assert mean_abs[-1] == pytest.approx(G[f"{tag}_mean_abs"][i], rel=2e-6)
            assert mean == pytest.approx(G[f"{tag}_mean"][i], rel=1e-4, abs=1e-6)
            assert np.sqrt(var) == pytest.approx(G[f"{tag}_std"][i], rel=2e-6)
        s, z = oracle.qparams(run_min, run_max, 8, sym)
        exp = G[f"{tag}_minmax_scale_zp"]
        assert (run_min, run_max, s, z) == tuple(exp), (tag, (run_min, run_max, s, z), exp)
        # fixed-qparam quantisation uses the QUANTIZER's bit-width with the 8-bit observer's scale/zp
        qmin, qmax = (-(2 ** (bits - 1)), 2 ** (bits - 1) - 1) if sym else (0, 2 ** bits - 1)
        y = oracle.fake_quant_fwd(G[f"{tag}_in0"], s, z, qmin, qmax)
        assert bits_equal(y, G[f"{tag}_yq_fixed"]), first_mismatch(y, G[f"{tag}_yq_fixed"])
        # LSQ init from the reference's own per-call means is bit-exact; from our fp64 means it is ~1e-7
        assert oracle.lsq_init_scale(G[f"{tag}_mean_abs"], bits) == float(G[f"{tag}_lsq_init"])
        assert oracle.lsq_init_scale(mean_abs, bits) == pytest.approx(float(G[f"{tag}_lsq_init"]), rel=2e-6)
        assert str(G[f"{tag}_param_dtype"]) == "torch.float64"
        assert float(G[f"{tag}_zp_after_learn"]) == 0.0  # quantization_manager.py:50,103
        if sym:
            x, g = G[f"{tag}_in1"], G[f"{tag}_learn_g"]
            s0 = float(G[f"{tag}_lsq_init"])
            y = oracle.fake_quant_fwd(x, s0, 0, qmin, qmax)
            assert bits_equal(y, G[f"{tag}_learn_y"])
            gs = oracle.grad_scale(qmax, x.size)
            dx, ds, _ = oracle.fake_quant_bwd(x, g, s0, 0, qmin, qmax, grad_scale=gs)
            assert bits_equal(dx, G[f"{tag}_learn_dx"])
            mass = gs * sum_mass(x, g, s0, 0, qmin, qmax)
            assert abs(ds[0] - G[f"{tag}_learn_ds"][0]) <= 1e-5 * mass


def test_lsq_per_channel():
    G = load_golden("lsq_per_channel")
    for tag in G["cases"]:
        qmin, qmax, config_act = (int(v) for v in G[f"{tag}_qp"])
        x, g = G[f"{tag}_x"], G[f"{tag}_g"]
        s, zf = G[f"{tag}_scale"], G[f"{tag}_zpf"]
        C = x.shape[1]
        y = oracle.fake_quant_fwd(x, s, zf, qmin, qmax, ch_axis=1, zp_learned=True)
        assert bits_equal(y, G[f"{tag}_y"]), (tag, first_mismatch(y, G[f"{tag}_y"]))
        gs = oracle.grad_scale(qmax, x.size, C) * (5000.0 if config_act else 1.0)  # lsq_module.py:151-152
        dx, ds, dz = oracle.fake_quant_bwd(x, g, s, zf, qmin, qmax, ch_axis=1, zp_learned=True,
                                           grad_scale=gs, want_dz=True)
        assert bits_equal(dx, G[f"{tag}_dx"]), (tag, first_mismatch(dx, G[f"{tag}_dx"]))
        for c in range(C):
            zeff = float(np.clip(np.rint(zf[c]), qmin, qmax))
            mass = gs * sum_mass(x[:, c], g[:, c], s[c], zeff, qmin, qmax)
            assert abs(ds[c] - G[f"{tag}_ds"][c]) <= 1e-5 * mass, (tag, c, ds[c], G[f"{tag}_ds"][c])
            zmass = gs * float(np.sum(np.abs(g[:, c].astype(np.float64) * s[c])))
            assert abs(dz[c] - G[f"{tag}_dz"][c]) <= 1e-5 * zmass, (tag, c, dz[c], G[f"{tag}_dz"][c])


def test_bn_fold_bit_exact():
    G = load_golden("bn_fold")
    for tag in G["cases"]:
        W, b, bn, eps = G[f"{tag}_W"], G[f"{tag}_b"], G[f"{tag}_bn"], float(G[f"{tag}_eps"])
        Wf, bf = oracle.bn_fold(W, b if b.size else None, bn[0], bn[1], bn[2], bn[3], eps)
        assert bits_equal(Wf, G[f"{tag}_Wf"]), (tag, first_mismatch(Wf, G[f"{tag}_Wf"]))
        assert bits_equal(bf, G[f"{tag}_bf"]), (tag, first_mismatch(bf, G[f"{tag}_bf"]))


def test_bn_reestimate():
    G = load_golden("bn_reestimate")
    k = int(G["num_batches"])
    rm, rv = oracle.bn_reestimate(list(G["conv_out"][:k]))
    np.testing.assert_allclose(rm, G["running_mean"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(rv, G["running_var"], rtol=1e-5, atol=1e-7)
    assert float(G["momentum_after"]) == pytest.approx(0.03) and bool(G["training_after"]) is False


def test_tiny_e2e_qparams_from_golden_activations():
    """Calibrated scales of the end-to-end fixture follow from the recorded min/max (bit-exact)."""
    G = load_golden("tiny_e2e")
    for n in G["fused_names"]:
        for kind in ("weight_quantizer", "activation_quantizer"):
            mn, mx, s, z = G[f"calib_{n}_{kind}"]
            s2, z2 = oracle.qparams(mn, mx, 8, True)
            assert (s2, z2) == (s, z)
            assert oracle.lsq_init_scale(G[f"calib_{n}_{kind}_mean_abs"], 8) == float(G[f"init_{n}_{kind}"])


def test_torch_port_matches_golden():
    """The torch-eager port timed as the CPU baseline reproduces the reference's outputs bit for bit."""
    import torch
    from oracle import torch_port
    G = load_golden("uniform_fixed")
    for tag in list(G["cases"])[:4]:
        scale, zp, qmin, qmax, bits, sym = G[f"{tag}_qp"]
        y, dx = torch_port.fwd_bwd(torch.as_tensor(G[f"{tag}_x"]), torch.as_tensor(G[f"{tag}_g"]), float(scale), int(zp),
                                   int(qmin), int(qmax))
        assert bits_equal(y.numpy(), G[f"{tag}_y"]) and bits_equal(dx.numpy(), G[f"{tag}_dx"])
    G = load_golden("uniform_learned")
    scale, zf, qmin, qmax, bits, sym, gs = G["b8_sym_0_qp"]
    y, dx, ds = torch_port.fwd_bwd(torch.as_tensor(G["b8_sym_0_x"]), torch.as_tensor(G["b8_sym_0_g"]), scale, 0,
                                   int(qmin), int(qmax), learn=True)
    assert bits_equal(y.numpy(), G["b8_sym_0_y"]) and bits_equal(dx.numpy(), G["b8_sym_0_dx"])
    assert ds.item() == G["b8_sym_0_ds"][0]


@pytest.mark.parametrize("seed", [0, 11])
def test_oracle_against_the_live_reference(seed):
    """oracle/check_live_reference.py: the C oracle beside the UNMODIFIED reference (torch on CPU) on fresh random
    cases -- UniformQuantizer fixed / learnable, MinMaxObserver traces, LSQFakeQuantize per channel.  Build container
    only (skipped where /root/reference is absent); the golden vectors cover the same ground everywhere else."""
    import subprocess
    import sys
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "oracle", "check_live_reference.py"), "--cases", "120",
                          "--seed", str(seed)], cwd="/tmp", capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-500:])
    assert "0 mismatches" in out.stdout
