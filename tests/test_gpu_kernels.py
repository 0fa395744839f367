"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle and the reference's golden vectors.

Bit-exact: fake-quantised values, integer codes, dx, observer min/max, scales / zero-points, BN fold.
Tolerance (stated per test): order-dependent sums -- ds, dz, observer means, BN moments.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import bits_equal, first_mismatch, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vsiquantization_b200 import ops as _ops
    return _ops


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def sum_mass(x, g, s, z, qmin, qmax):
    s32 = np.float32(s)
    with np.errstate(all="ignore"):
        v = x.astype(np.float32) / s32
        q = np.clip(np.rint(v + np.float32(z)), qmin, qmax)
        return float(np.sum(np.abs(g.astype(np.float64) * (q - z))) + np.sum(np.abs(g.astype(np.float64) * v)))


def assert_sums_close(st, so, rel=1e-6):
    """sum|x|, sum x, sum x^2 against the fp64 oracle; sum x may cancel, so it is judged against sum|x|."""
    st, so = np.asarray(st, np.float64), np.asarray(so, np.float64)
    assert np.all(np.abs(st[:, 2] - so[:, 2]) <= rel * so[:, 2] + 1e-30), (st[:, 2], so[:, 2])
    assert np.all(np.abs(st[:, 3] - so[:, 3]) <= rel * so[:, 2] + 1e-30), (st[:, 3], so[:, 3])
    assert np.all(np.abs(st[:, 4] - so[:, 4]) <= rel * so[:, 4] + 1e-30), (st[:, 4], so[:, 4])


# ---------------------------------------------------------------------------------- golden vectors
def test_golden_uniform_fixed(ops):
    G = load_golden("uniform_fixed")
    for tag in G["cases"]:
        scale, zp, qmin, qmax, bits, sym = G[f"{tag}_qp"]
        spec = ops.QSpec(int(qmin), int(qmax))
        x, g = dev(G[f"{tag}_x"]), dev(G[f"{tag}_g"])
        y, codes = ops.fake_quant_forward(x, float(scale), int(zp), spec, want_codes=True)
        assert bits_equal(y.cpu().numpy(), G[f"{tag}_y"]), (tag, first_mismatch(y.cpu().numpy(), G[f"{tag}_y"]))
        ref_codes = G[f"{tag}_codes"]
        ok = ~np.isnan(ref_codes)
        assert np.array_equal(codes.cpu().numpy().astype(np.float32)[ok], ref_codes[ok]), tag
        dx = ops.fake_quant_backward_ste(x, g, float(scale), int(zp), spec)
        assert bits_equal(dx.cpu().numpy(), G[f"{tag}_dx"]), (tag, first_mismatch(dx.cpu().numpy(), G[f"{tag}_dx"]))
        y2, dx2 = ops.fake_quant_forward_backward(x, g, float(scale), int(zp), spec)
        assert bits_equal(y2.cpu().numpy(), G[f"{tag}_y"]) and bits_equal(dx2.cpu().numpy(), G[f"{tag}_dx"]), tag


def test_golden_uniform_learned(ops):
    G = load_golden("uniform_learned")
    for tag in G["cases"]:
        scale, zf, qmin, qmax, bits, sym, gs = G[f"{tag}_qp"]
        sym = bool(sym)
        x, g = G[f"{tag}_x"], G[f"{tag}_g"]
        spec = ops.QSpec(int(qmin), int(qmax), zp_learned=not sym)
        s_t = torch.tensor(scale, dtype=torch.float64, device="cuda")  # 0-dim fp64 like the reference Parameter
        z_t = torch.tensor(zf, dtype=torch.float32, device="cuda") if not sym else 0
        y = ops.fake_quant_forward(dev(x), s_t, z_t, spec)
        assert bits_equal(y.cpu().numpy(), G[f"{tag}_y"]), (tag, first_mismatch(y.cpu().numpy(), G[f"{tag}_y"]))
        dx, ds, dz = ops.lsq_backward(dev(x), dev(g), s_t, z_t, spec, gs, want_dz=not sym, ds_dtype=torch.float64)
        assert bits_equal(dx.cpu().numpy(), G[f"{tag}_dx"]), (tag, first_mismatch(dx.cpu().numpy(), G[f"{tag}_dx"]))
        zeff = float(np.clip(np.rint(np.float32(zf)), qmin, qmax)) if not sym else 0.0
        mass = gs * sum_mass(x, g, scale, zeff, qmin, qmax)
        assert ds.dtype == torch.float64
        assert abs(ds.item() - G[f"{tag}_ds"][0]) <= 1e-5 * mass, (tag, ds.item(), G[f"{tag}_ds"][0], mass)
        # and against the fp64 oracle ("truth") much tighter
        _, ds_o, dz_o = oracle.fake_quant_bwd(x, g, scale, zf, int(qmin), int(qmax), zp_learned=not sym, grad_scale=gs,
                                              want_dz=not sym)
        assert abs(ds.item() - ds_o[0]) <= 2e-6 * mass, (tag, ds.item(), ds_o[0])
        if not sym:
            zmass = gs * float(np.sum(np.abs(g.astype(np.float64) * np.float32(scale))))
            assert abs(dz.item() - G[f"{tag}_dz"][0]) <= 1e-5 * zmass, (tag, dz.item(), G[f"{tag}_dz"][0])
            assert abs(dz.item() - dz_o[0]) <= 2e-6 * zmass


def test_golden_funlsq(ops):
    G = load_golden("funlsq")
    s, z, qmin, qmax, gsc = G["qp"]
    spec = ops.QSpec(int(qmin), int(qmax), mask_mode=1)
    x, g = G["x"], G["g"]
    dx, ds, _ = ops.lsq_backward(dev(x), dev(g), torch.tensor([s], dtype=torch.float32, device="cuda"), 0, spec, gsc)
    assert bits_equal(dx.cpu().numpy(), G["dx"]), first_mismatch(dx.cpu().numpy(), G["dx"])
    mass = gsc * float(np.sum(np.abs(g.astype(np.float64))) * 0.5 + 1.0)
    assert abs(ds.item() - G["ds"][0]) <= 1e-5 * mass


def test_golden_lsq_per_channel(ops):
    G = load_golden("lsq_per_channel")
    for tag in G["cases"]:
        qmin, qmax, config_act = (int(v) for v in G[f"{tag}_qp"])
        x, g = G[f"{tag}_x"], G[f"{tag}_g"]
        s, zf = G[f"{tag}_scale"], G[f"{tag}_zpf"]
        C = x.shape[1]
        spec = ops.QSpec(qmin, qmax, ch_axis=1, zp_learned=True)
        s_t, z_t = dev(s).view(1, C, 1, 1), dev(zf).view(1, C, 1, 1)
        y = ops.fake_quant_forward(dev(x), s_t, z_t, spec)
        assert bits_equal(y.cpu().numpy(), G[f"{tag}_y"]), (tag, first_mismatch(y.cpu().numpy(), G[f"{tag}_y"]))
        gs = ops.lsq_grad_scale(qmax, x.size, C) * (5000.0 if config_act else 1.0)
        dx, ds, dz = ops.lsq_backward(dev(x), dev(g), s_t, z_t, spec, gs, want_dz=True)
        assert bits_equal(dx.cpu().numpy(), G[f"{tag}_dx"]), (tag, first_mismatch(dx.cpu().numpy(), G[f"{tag}_dx"]))
        ds, dz = ds.cpu().numpy().astype(np.float64), dz.cpu().numpy().astype(np.float64)
        for c in range(C):
            zeff = float(np.clip(np.rint(zf[c]), qmin, qmax))
            mass = gs * sum_mass(x[:, c], g[:, c], s[c], zeff, qmin, qmax)
            assert abs(ds[c] - G[f"{tag}_ds"][c]) <= 1e-5 * mass, (tag, c, ds[c], G[f"{tag}_ds"][c])
            zmass = gs * float(np.sum(np.abs(g[:, c].astype(np.float64) * s[c])))
            assert abs(dz[c] - G[f"{tag}_dz"][c]) <= 1e-5 * zmass, (tag, c, dz[c], G[f"{tag}_dz"][c])


def test_golden_observer(ops):
    G = load_golden("observer")
    for tag in G["cases"]:
        name, symtag, btag = tag.rsplit("_", 2)
        sym, bits = symtag == "sym", int(btag[1:])
        state = ops.new_observer_state(1, "cuda")
        trace = G[f"{tag}_trace"]
        for i in range(int(G[f"{name}_n"])):
            ops.observe(dev(G[f"{name}_in{i}"]), None, state, bits, sym, 1e-8, want_stats=False)
            st = state.cpu().numpy()[0]
            exp = trace[i]
            for k in range(4):
                assert st[k] == exp[k] or (np.isnan(st[k]) and np.isnan(exp[k])), (tag, i, k, st[:4], exp)


def test_golden_manager_stats(ops):
    G = load_golden("manager")
    for tag in G["cases"]:
        bits = int(tag[1])
        sym = tag.endswith("_sym")
        state = ops.new_observer_state(1, "cuda")
        for i in range(3):
            x = G[f"{tag}_in{i}"]
            st = ops.observe(dev(x), None, state, 8, sym).cpu().numpy()[0]
            assert st[2] / x.size == pytest.approx(G[f"{tag}_mean_abs"][i], rel=2e-6)
            assert st[3] / x.size == pytest.approx(G[f"{tag}_mean"][i], rel=1e-4, abs=1e-6)
        s = state.cpu().numpy()[0]
        assert tuple(s[:4]) == tuple(G[f"{tag}_minmax_scale_zp"])
        assert s[4] == 3 and s[5] / 3 == pytest.approx(np.mean(G[f"{tag}_mean_abs"]), rel=2e-6)
        assert s[7] / 3 == pytest.approx(np.mean(G[f"{tag}_std"]), rel=2e-6)
        out = torch.empty(1, dtype=torch.float64, device="cuda")
        ops.lsq_init_scale(state, bits, out)
        # order-dependent fp32 means in the reference: tolerance, not bit-exact
        assert out.item() == pytest.approx(float(G[f"{tag}_lsq_init"]), rel=2e-6)


def test_golden_bn_fold(ops):
    G = load_golden("bn_fold")
    for tag in G["cases"]:
        W, b, bn, eps = G[f"{tag}_W"], G[f"{tag}_b"], G[f"{tag}_bn"], float(G[f"{tag}_eps"])
        Wf, bf, _, _ = ops.bn_fold(dev(W), dev(b) if b.size else None, dev(bn[0]), dev(bn[1]), dev(bn[2]), dev(bn[3]), eps)
        assert bits_equal(Wf.cpu().numpy(), G[f"{tag}_Wf"]), (tag, first_mismatch(Wf.cpu().numpy(), G[f"{tag}_Wf"]))
        assert bits_equal(bf.cpu().numpy(), G[f"{tag}_bf"]), (tag, first_mismatch(bf.cpu().numpy(), G[f"{tag}_bf"]))
        # fused fold + fake-quant + stats == fold, then oracle fake-quant / stats of the folded weight
        spec = ops.QSpec(-8, 7)
        Wf2, bf2, Wq, st = ops.bn_fold(dev(W), dev(b) if b.size else None, dev(bn[0]), dev(bn[1]), dev(bn[2]), dev(bn[3]),
                                       eps, scale=0.11, zero_point=0, spec=spec, want_stats=True)
        assert bits_equal(Wf2.cpu().numpy(), G[f"{tag}_Wf"]) and bits_equal(bf2.cpu().numpy(), G[f"{tag}_bf"])
        assert bits_equal(Wq.cpu().numpy(), oracle.fake_quant_fwd(G[f"{tag}_Wf"], 0.11, 0, -8, 7))
        so = oracle.minmax_stats(G[f"{tag}_Wf"])[0]
        sg = st.cpu().numpy()[0]
        assert sg[0] == so[0] and sg[1] == so[1]
        assert_sums_close(sg[None, :], so[None, :])
        # per-channel (ch_axis 0) weight fake-quant in the fold
        C = W.shape[0]
        sc = np.linspace(0.05, 0.2, C).astype(np.float32)
        _, _, Wq2, _ = ops.bn_fold(dev(W), None, dev(bn[0]), dev(bn[1]), dev(bn[2]), dev(bn[3]), eps, scale=dev(sc),
                                   zero_point=0, spec=ops.QSpec(-8, 7, ch_axis=0))
        assert bits_equal(Wq2.cpu().numpy(), oracle.fake_quant_fwd(G[f"{tag}_Wf"], sc, 0, -8, 7, ch_axis=0))


def test_golden_bn_reestimate(ops):
    G = load_golden("bn_reestimate")
    k = int(G["num_batches"])
    C = G["conv_out"].shape[2]
    mean_sum = torch.zeros(C, device="cuda")
    var_sum = torch.zeros(C, device="cuda")
    for i in range(k):
        ops.bn_batch_moments(dev(G["conv_out"][i]), mean_sum, var_sum)
    rm, rv = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.bn_reestimate_finish(mean_sum, var_sum, k, rm, rv)
    np.testing.assert_allclose(rm.cpu().numpy(), G["running_mean"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(rv.cpu().numpy(), G["running_var"], rtol=1e-5, atol=1e-7)
    rm_o, rv_o = oracle.bn_reestimate(list(G["conv_out"][:k]))
    np.testing.assert_allclose(rm.cpu().numpy(), rm_o, rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(rv.cpu().numpy(), rv_o, rtol=2e-6, atol=1e-8)


# ------------------------------------------------------------------------ seeded inputs vs the oracle
PER_TENSOR_SIZES = [1, 7, 8, 9, 31, 255, 1023, 1024, 1025, 2047, 2048, 2049, 8191, 8192, 8193, 100003, (1 << 20) + 5]


@pytest.mark.parametrize("n", PER_TENSOR_SIZES)
def test_per_tensor_ragged_sizes(ops, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 2).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    for (qmin, qmax, s, z) in ((-128, 127, 3.0 / 127, 0), (0, 15, 0.4, 8), (-2, 1, 0.9, 0)):
        spec = ops.QSpec(qmin, qmax)
        y = ops.fake_quant_forward(dev(x), s, z, spec)
        assert bits_equal(y.cpu().numpy(), oracle.fake_quant_fwd(x, s, z, qmin, qmax)), (n, qmin)
        gs = ops.lsq_grad_scale(qmax, n)
        dx_o, ds_o, _ = oracle.fake_quant_bwd(x, g, s, z, qmin, qmax, grad_scale=gs)
        dx = ops.fake_quant_backward_ste(dev(x), dev(g), s, z, spec)
        assert bits_equal(dx.cpu().numpy(), dx_o), (n, qmin)
        s_t = torch.tensor(s, dtype=torch.float64, device="cuda")
        dx2, ds, _ = ops.lsq_backward(dev(x), dev(g), s_t, z, spec, gs, ds_dtype=torch.float64)
        assert bits_equal(dx2.cpu().numpy(), dx_o)
        mass = gs * sum_mass(x, g, s, z, qmin, qmax)
        assert abs(ds.item() - ds_o[0]) <= 2e-6 * mass + 1e-30, (n, ds.item(), ds_o[0])
        st = ops.observe(dev(x)).cpu().numpy()[0]
        so = oracle.minmax_stats(x)[0]
        assert st[0] == so[0] and st[1] == so[1]
        assert_sums_close(st[None, :], so[None, :])


@pytest.mark.parametrize("shape,ch_axis", [
    ((16, 3, 3, 3), 0), ((64, 32, 3, 3), 0), ((7, 4609), 0), ((5, 2048), 0), ((3, 10007), 0),
    ((2, 16, 10, 10), 1), ((3, 8, 20, 20), 1), ((2, 5, 47, 47), 1), ((2, 3, 160, 160), 1), ((1, 4, 3, 2731), 1),
    ((4, 6), 1), ((9, 1), 0),
])
def test_per_channel_layouts(ops, shape, ch_axis):
    rng = np.random.default_rng(sum(shape))
    C = shape[ch_axis]
    x = (rng.standard_normal(shape) * 2 + 0.3).astype(np.float32)
    g = rng.standard_normal(shape).astype(np.float32)
    s = (0.02 * rng.uniform(0.5, 2.0, C)).astype(np.float32)
    zf = rng.uniform(100, 150, C).astype(np.float32)
    qshape = [1] * len(shape)
    qshape[ch_axis] = C
    spec = ops.QSpec(0, 255, ch_axis=ch_axis, zp_learned=True)
    s_t, z_t = dev(s).view(qshape), dev(zf).view(qshape)
    y = ops.fake_quant_forward(dev(x), s_t, z_t, spec)
    y_o = oracle.fake_quant_fwd(x, s, zf, 0, 255, ch_axis=ch_axis, zp_learned=True)
    assert bits_equal(y.cpu().numpy(), y_o), first_mismatch(y.cpu().numpy(), y_o)
    gs = ops.lsq_grad_scale(255, x.size, C)
    dx_o, ds_o, dz_o = oracle.fake_quant_bwd(x, g, s, zf, 0, 255, ch_axis=ch_axis, zp_learned=True, grad_scale=gs, want_dz=True)
    dx, ds, dz = ops.lsq_backward(dev(x), dev(g), s_t, z_t, spec, gs, want_dz=True)
    assert bits_equal(dx.cpu().numpy(), dx_o), first_mismatch(dx.cpu().numpy(), dx_o)
    xm = np.moveaxis(x, ch_axis, 0).reshape(C, -1)
    gm = np.moveaxis(g, ch_axis, 0).reshape(C, -1)
    ds, dz = ds.cpu().numpy().astype(np.float64), dz.cpu().numpy().astype(np.float64)
    for c in range(C):
        zeff = float(np.clip(np.rint(zf[c]), 0, 255))
        mass = gs * sum_mass(xm[c], gm[c], s[c], zeff, 0, 255)
        assert abs(ds[c] - ds_o[c]) <= 2e-6 * mass + 1e-30, (c, ds[c], ds_o[c])
        zmass = gs * float(np.sum(np.abs(gm[c].astype(np.float64) * s[c])))
        assert abs(dz[c] - dz_o[c]) <= 2e-6 * zmass + 1e-30, (c, dz[c], dz_o[c])
    st = ops.observe(dev(x), ch_axis=ch_axis).cpu().numpy()
    so = oracle.minmax_stats(x, ch_axis=ch_axis)
    assert np.array_equal(st[:, :2], so[:, :2])
    assert_sums_close(st, so)
    # STE backward and the fused sweep agree with the LSQ backward's dx
    dx2 = ops.fake_quant_backward_ste(dev(x), dev(g), s_t, z_t, spec)
    assert bits_equal(dx2.cpu().numpy(), dx_o)


def test_unaligned_views(ops):
    """Tensors whose storage offset breaks 32-byte alignment take the scalar instantiation."""
    rng = np.random.default_rng(5)
    base = torch.as_tensor(rng.standard_normal(40000).astype(np.float32)).cuda()
    gbase = torch.as_tensor(rng.standard_normal(40000).astype(np.float32)).cuda()
    for off in (1, 3, 5):
        x = base[off:off + 30001]
        g = gbase[off:off + 30001]
        assert x.data_ptr() % 32 != 0
        xn, gn = x.cpu().numpy(), g.cpu().numpy()
        spec = ops.QSpec(-8, 7)
        y = ops.fake_quant_forward(x, 0.3, 0, spec)
        assert bits_equal(y.cpu().numpy(), oracle.fake_quant_fwd(xn, 0.3, 0, -8, 7))
        dx = ops.fake_quant_backward_ste(x, g, 0.3, 0, spec)
        assert bits_equal(dx.cpu().numpy(), oracle.fake_quant_bwd(xn, gn, 0.3, 0, -8, 7, want_ds=False)[0])
        st = ops.observe(x).cpu().numpy()[0]
        so = oracle.minmax_stats(xn)[0]
        assert st[0] == so[0] and st[1] == so[1]


def test_special_values_and_odd_scales(ops):
    """denormals / inf / NaN / huge inputs and scales outside the fast-division window stay bit-exact."""
    sp = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-40, 1.17549435e-38, np.inf, -np.inf, np.nan, 3.4e38, -3.4e38, 1e-30,
                   6e-39, 2.0 ** -61, 2.0 ** -60, 2.0 ** 60, 2.0 ** 61, 1.5, -2.5, 0.5, -0.5], dtype=np.float32)
    x = np.tile(sp, 50)
    g = np.random.default_rng(0).standard_normal(x.size).astype(np.float32)
    for s in (0.0236, 2.0 ** -41, 2.0 ** 41, 1e-38, 3e38, -0.05, 1.0):
        for (qmin, qmax, z) in ((-128, 127, 0), (0, 255, 128)):
            spec = ops.QSpec(qmin, qmax)
            y = ops.fake_quant_forward(dev(x), s, z, spec)
            y_o = oracle.fake_quant_fwd(x, s, z, qmin, qmax)
            assert bits_equal(y.cpu().numpy(), y_o), (s, qmin, first_mismatch(y.cpu().numpy(), y_o))
            dx = ops.fake_quant_backward_ste(dev(x), dev(g), s, z, spec)
            dx_o = oracle.fake_quant_bwd(x, g, s, z, qmin, qmax, want_ds=False)[0]
            assert bits_equal(dx.cpu().numpy(), dx_o), (s, qmin, first_mismatch(dx.cpu().numpy(), dx_o))


@pytest.mark.parametrize("scale", [3.0 / 127, 3.0 / 7, 0.0173, 1.0, 2.0 ** -5, 1.9999999, 1.0000001, 0.3333333, 7.7e-4,
                                   123.456, 2.0 ** -40, 2.0 ** 40, 1.1754944e-38, -0.021])
def test_division_exhaustive(ops, scale):
    """The hoisted-reciprocal arithmetic equals the IEEE sequences for ALL 2^32 input values:
    mode 0 = x / s (forward), mode 1 = RN(RN(g*s) / s) (dx)."""
    assert ops.selftest_division(scale, 0) == 0
    assert ops.selftest_division(scale, 1) == 0


def test_division_random_scales(ops):
    rng = np.random.default_rng(123)
    for s in np.exp(rng.uniform(np.log(1e-6), np.log(1e3), 24)).astype(np.float32):
        assert ops.selftest_division(float(s), 0) == 0, s
        assert ops.selftest_division(float(s), 1) == 0, s


def test_empty_and_errors(ops):
    e = torch.empty(0, device="cuda")
    assert ops.fake_quant_forward(e, 0.1, 0, ops.QSpec(-8, 7)).numel() == 0
    assert ops.fake_quant_backward_ste(e, e, 0.1, 0, ops.QSpec(-8, 7)).numel() == 0
    with pytest.raises(RuntimeError):
        ops.fake_quant_forward(torch.zeros(4), 0.1, 0, ops.QSpec(-8, 7))  # CPU tensor: no fallback
    with pytest.raises(TypeError):
        ops.fake_quant_forward(torch.zeros(4, device="cuda", dtype=torch.float16), 0.1, 0, ops.QSpec(-8, 7))
    from vsiquantization_b200._lib import VsiqError
    with pytest.raises(VsiqError):
        ops.fake_quant_forward(torch.zeros(4, device="cuda"), 0.1, 0, ops.QSpec(7, 7))  # qmin >= qmax


# --------------------------------------------------------------- red zones: nothing is written outside the outputs
_SENT32 = 0x7FC0DEAD  # a NaN with a payload no kernel produces
_PAD = 4096           # elements of sentinel on each side of an output


class _RedZone:
    """An output carved out of a sentinel-filled buffer at an element offset (so every alignment class of the stores is
    visited); untouched() is true when every word outside the carved range still holds the sentinel."""

    def __init__(self, n, off, dtype=torch.float32):
        self.n, self.off, self.dtype = n, off, dtype
        if dtype == torch.float32:
            self.raw = torch.full((2 * _PAD + n + 16,), _SENT32, dtype=torch.int32, device="cuda")
            self.buf = self.raw.view(torch.float32)
        else:
            self.raw = torch.full((2 * _PAD + n + 64,), 0xA5, dtype=torch.uint8, device="cuda")
            self.buf = self.raw.view(dtype)
        self.out = self.buf[_PAD + off:_PAD + off + n]

    def untouched(self):
        sent = _SENT32 if self.dtype == torch.float32 else 0xA5
        lo, hi = self.raw[:_PAD + self.off], self.raw[_PAD + self.off + self.n:]
        return bool((lo == sent).all()) and bool((hi == sent).all())


def _same_bits(a, b):
    return torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


@pytest.mark.parametrize("off", [0, 1, 5, 8])
@pytest.mark.parametrize("shape,ch_axis", [((1,), None), ((7,), None), ((31,), None), ((255,), None), ((1027,), None),
                                           ((8191,), None), ((8200,), None), ((65536 + 24,), None), ((300007,), None),
                                           ((5, 37), 0), ((64, 1030), 0), ((3, 70001), 0), ((300, 9), 0),
                                           ((3, 6, 50, 38), 1)])
def test_no_write_outside_the_outputs(ops, shape, ch_axis, off):
    """Forward, STE backward and the fused sweep write exactly their outputs: y / dx live inside sentinel-filled
    buffers at element offsets 0, 1, 5, 8 (32-byte aligned, unaligned, 32-byte aligned again), the input is offset the
    same way, and the result equals the one computed into a fresh, aligned allocation bit for bit."""
    rng = np.random.default_rng(len(shape) * 131 + shape[-1] + off)
    n = int(np.prod(shape))
    C = 1 if ch_axis is None else shape[ch_axis]
    xs, gs_ = _RedZone(n, off), _RedZone(n, off)
    xs.out.copy_(dev((rng.standard_normal(n) * 3).astype(np.float32)))
    gs_.out.copy_(dev(rng.standard_normal(n).astype(np.float32)))
    x, g = xs.out.view(shape), gs_.out.view(shape)
    spec = ops.QSpec(-8, 7, ch_axis=ch_axis)
    s = 0.4 if ch_axis is None else dev((rng.uniform(0.5, 2.0, C) * 0.4).astype(np.float32))
    z = 0 if ch_axis is None else torch.zeros(C, device="cuda")
    y_ref = ops.fake_quant_forward(x.clone(), s, z, spec)
    dx_ref = ops.fake_quant_backward_ste(x.clone(), g.clone(), s, z, spec)
    yz, dz = _RedZone(n, off), _RedZone(n, off)
    y = ops.fake_quant_forward(x, s, z, spec, out=yz.out.view(shape))
    dx = ops.fake_quant_backward_ste(x, g, s, z, spec, out=dz.out.view(shape))
    assert yz.untouched() and dz.untouched() and xs.untouched() and gs_.untouched()
    assert _same_bits(y, y_ref) and _same_bits(dx, dx_ref)
    if ch_axis is None:
        y2z, dx2z = _RedZone(n, off), _RedZone(n, off)
        y2, dx2 = ops.fake_quant_forward_backward(x, g, s, z, spec, y_out=y2z.out.view(shape), dx_out=dx2z.out.view(shape))
        assert y2z.untouched() and dx2z.untouched()
        assert _same_bits(y2, y_ref) and _same_bits(dx2, dx_ref)


@pytest.mark.parametrize("bits", [4, 8, 16])
@pytest.mark.parametrize("shape,ch_axis", [((2,), None), ((30,), None), ((8190,), None), ((8194,), None), ((70002,), None),
                                           (((1 << 21) + 6,), None), ((5, 38), 0), ((3, 70002), 0), ((3, 6, 50, 38), 1)])
def test_code_export_writes_only_its_codes(ops, bits, shape, ch_axis):
    """Packed codes (two per byte, one per byte, one per 16 bits) inside a 0xA5-filled byte buffer: ragged heads and
    tails, per-channel rows and multi-batch tiles leave every byte outside the codes alone."""
    rng = np.random.default_rng(bits + shape[-1])
    n = int(np.prod(shape))
    C = 1 if ch_axis is None else shape[ch_axis]
    x = dev((rng.standard_normal(shape) * 3).astype(np.float32))
    qmin, qmax = (0, 15) if bits == 4 else (-(1 << (bits - 1)), (1 << (bits - 1)) - 1)
    spec = ops.QSpec(qmin, qmax, ch_axis=ch_axis)
    s = 6.0 / (qmax - qmin) if ch_axis is None else dev((rng.uniform(0.5, 2.0, C) * 6.0 / (qmax - qmin)).astype(np.float32))
    z = (8 if bits == 4 else 0) if ch_axis is None else torch.full((C,), 8.0 if bits == 4 else 0.0, device="cuda")
    _, ref = ops.quantize_codes(x, s, z, spec, bits, want_y=False)
    nbytes = n // 2 if bits == 4 else n * (bits // 8)
    rz = _RedZone(nbytes, 16, torch.uint8)
    c_shape = shape[:-1] + (shape[-1] // 2,) if bits == 4 else shape
    out = rz.out.view(ref.dtype).view(c_shape)
    _, codes = ops.quantize_codes(x, s, z, spec, bits, want_y=False, codes_out=out)
    assert rz.untouched()
    assert torch.equal(codes, ref)


@pytest.mark.parametrize("shape", [(2, 16, 9, 7), (1, 48, 33, 5), (2, 512, 6, 5), (5, 4, 41, 40)])
def test_channels_last_epilogue_writes_only_its_output(ops, shape):
    """The NHWC epilogue forward (bias + ReLU + per-channel fake-quant) into a sentinel-padded channels_last output."""
    rng = np.random.default_rng(sum(shape))
    N, C, H, W = shape
    n = N * C * H * W
    x = dev((rng.standard_normal(shape) * 2).astype(np.float32)).contiguous(memory_format=torch.channels_last)
    b = dev(rng.standard_normal(C).astype(np.float32))
    s = dev((rng.uniform(0.5, 2.0, C) * 0.02).astype(np.float32))
    z = torch.zeros(C, device="cuda")
    spec = ops.QSpec(-128, 127, ch_axis=1, pre_relu=True)
    ref = ops.ci_forward(x, b, s, z, spec)
    rz = _RedZone(n, 4)
    out = rz.out.view(N, H, W, C).permute(0, 3, 1, 2)
    y = ops.ci_forward(x, b, s, z, spec, out=out)
    assert rz.untouched()
    assert _same_bits(y.permute(0, 2, 3, 1), ref.permute(0, 2, 3, 1))


# --------------------------------------------------------------- full-size, size-independent properties
@pytest.mark.parametrize("log2n", [26, 28])
def test_full_size_properties(ops, log2n):
    n = 1 << log2n
    torch.manual_seed(0)
    x = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda")
    s, qmin, qmax = 3.0 / 127, -128, 127
    spec = ops.QSpec(qmin, qmax)
    y, codes = ops.fake_quant_forward(x, s, 0, spec, want_codes=True)
    y1 = ops.fake_quant_forward(x, s, 0, spec)
    assert torch.equal(y, y1)
    # idempotence: fq(fq(x)) == fq(x)
    assert torch.equal(ops.fake_quant_forward(y1, s, 0, spec), y1)
    # codes in range, y == codes * s bitwise
    assert int(codes.min()) >= qmin and int(codes.max()) <= qmax
    assert torch.equal(codes.float() * torch.tensor(s, device="cuda", dtype=torch.float32), y1)
    # monotone: sorting commutes with quantisation on a sample
    xs = torch.sort(x[: 1 << 20]).values
    ys = ops.fake_quant_forward(xs, s, 0, spec)
    assert bool((ys[1:] >= ys[:-1]).all())
    # mask <=> in range; dx == g inside, 0 outside (to 1 ulp: dx = (g*s)/s)
    dx = ops.fake_quant_backward_ste(x, g, s, 0, spec)
    inside = (torch.round(x / torch.tensor(s, device="cuda")) >= qmin) & (torch.round(x / torch.tensor(s, device="cuda")) <= qmax)
    assert bool((dx[~inside] == 0).all())
    assert torch.allclose(dx[inside], g[inside], rtol=2e-7, atol=0)
    # the first 2^20 elements against the oracle, bit for bit
    k = 1 << 20
    xo, go = x[:k].cpu().numpy(), g[:k].cpu().numpy()
    assert bits_equal(y1[:k].cpu().numpy(), oracle.fake_quant_fwd(xo, s, 0, qmin, qmax))
    assert bits_equal(dx[:k].cpu().numpy(), oracle.fake_quant_bwd(xo, go, s, 0, qmin, qmax, want_ds=False)[0])
    # LSQ ds: linear in g (ds(2g) == 2 ds(g)) and equal to the chunked oracle sum
    s_t = torch.tensor(s, dtype=torch.float64, device="cuda")
    gs = ops.lsq_grad_scale(qmax, n)
    _, ds1, _ = ops.lsq_backward(x, g, s_t, 0, spec, gs, ds_dtype=torch.float64)
    _, ds2, _ = ops.lsq_backward(x, g * 2, s_t, 0, spec, gs, ds_dtype=torch.float64)
    assert ds2.item() == pytest.approx(2 * ds1.item(), rel=1e-12)
    acc = 0.0
    mass = 0.0
    step = 1 << 24
    for i in range(0, n, step):
        xo, go = x[i:i + step].cpu().numpy(), g[i:i + step].cpu().numpy()
        acc += oracle.fake_quant_bwd(xo, go, s, 0, qmin, qmax, grad_scale=gs)[1][0]
        mass += gs * sum_mass(xo, go, s, 0, qmin, qmax)
    assert abs(ds1.item() - acc) <= 2e-6 * mass
    # observer: exact extrema, sums to fp64 accuracy
    st = ops.observe(x).cpu().numpy()[0]
    assert st[0] == float(x.min()) and st[1] == float(x.max())
    assert abs(st[3] - float(x.double().sum())) <= 1e-6 * st[2]
    assert st[4] == pytest.approx(float((x.double() ** 2).sum()), rel=1e-6)


def test_maximum_size_2p30_windows(ops):
    """BASELINE configs[1] tops out at 2^30 elements (4 GiB per tensor: byte offsets cross 2^32).  Windows at the head,
    across the 2^31- and 2^32-byte marks and at the tail against the oracle, bit for bit (y, codes, STE dx, LSQ dx), plus
    whole-tensor properties: y == codes * s, extrema equal torch's, ds linear in g."""
    n = 1 << 30
    free, _ = torch.cuda.mem_get_info()
    if free < 30 * (1 << 30):
        pytest.skip("needs 30 GiB of free device memory")
    torch.manual_seed(30)
    x = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda")
    s, qmin, qmax = 3.0 / 127, -128, 127
    spec = ops.QSpec(qmin, qmax)
    y, codes = ops.fake_quant_forward(x, s, 0, spec, want_codes=True)
    _, codes_only = ops.quantize_codes(x, s, 0, spec, 8, want_y=False)
    assert torch.equal(codes, codes_only)
    assert torch.equal(codes.float() * torch.tensor(s, device="cuda", dtype=torch.float32), y)
    del codes_only
    dx = ops.fake_quant_backward_ste(x, g, s, 0, spec)
    s_t = torch.tensor(s, dtype=torch.float64, device="cuda")
    gs = ops.lsq_grad_scale(qmax, n)
    dx_l, ds1, _ = ops.lsq_backward(x, g, s_t, 0, spec, gs, ds_dtype=torch.float64)
    k = 1 << 18
    for start in (0, (1 << 29) - k // 2, (1 << 30) - k):  # element 2^29 = byte 2^31; the tail ends at byte 2^32
        xo, go = x[start:start + k].cpu().numpy(), g[start:start + k].cpu().numpy()
        y_o, c_o = oracle.fake_quant_fwd(xo, s, 0, qmin, qmax, want_codes=True)
        assert bits_equal(y[start:start + k].cpu().numpy(), y_o), start
        assert np.array_equal(codes[start:start + k].cpu().numpy().astype(np.float32), c_o), start
        dx_o = oracle.fake_quant_bwd(xo, go, s, 0, qmin, qmax, want_ds=False)[0]
        assert bits_equal(dx[start:start + k].cpu().numpy(), dx_o), start
        assert bits_equal(dx_l[start:start + k].cpu().numpy(), dx_o), start
    del y, codes, dx, dx_l
    g.mul_(2)
    _, ds2, _ = ops.lsq_backward(x, g, s_t, 0, spec, gs, ds_dtype=torch.float64)
    assert ds2.item() == pytest.approx(2 * ds1.item(), rel=1e-12)
    st = ops.observe(x).cpu().numpy()[0]
    assert st[0] == float(x.min()) and st[1] == float(x.max())
    assert st[4] == pytest.approx(float((x.double() ** 2).sum()), rel=1e-6)


# ---------------------------------------------------------- channel-innermost (NHWC / channels_last) kernels
@pytest.mark.parametrize("shape", [(2, 16, 9, 7), (3, 64, 20, 20), (1, 48, 33, 5), (2, 512, 6, 5), (2, 1024, 3, 3), (5, 4, 40, 40),
                                   (64, 32, 64, 64)])
@pytest.mark.parametrize("pcq,bias,relu", [(False, False, False), (True, False, True), (False, True, True), (True, True, True),
                                           (True, True, False)])
def test_channels_inner_matches_oracle(ops, shape, pcq, bias, relu):
    """NHWC kernels == the oracle applied to the NCHW tensor (after bias add / relu in fp32), bit for bit for y and dx;
    dscale / dzp / dbias to the fp64 oracle within summation tolerance."""
    rng = np.random.default_rng(sum(shape) + 7 * pcq + 3 * bias + relu)
    N, C, H, W = shape
    x = (rng.standard_normal(shape) * 2).astype(np.float32)
    g = rng.standard_normal(shape).astype(np.float32)
    b = (rng.standard_normal(C) * 0.5).astype(np.float32) if bias else None
    if pcq:
        s = (0.02 * rng.uniform(0.5, 2.0, C)).astype(np.float32)
        zf = rng.uniform(-3, 40, C).astype(np.float32)
    else:
        s = np.float32(0.023)
        zf = np.float32(3.3)
    spec = ops.QSpec(0, 255, ch_axis=1 if pcq else None, zp_learned=True, pre_relu=relu)
    xt = dev(x).contiguous(memory_format=torch.channels_last)
    gt = dev(g).contiguous(memory_format=torch.channels_last)
    assert ops.ci_supported(xt)
    s_t = dev(s).view(1, C, 1, 1) if pcq else torch.tensor(float(s), device="cuda")
    z_t = dev(zf).view(1, C, 1, 1) if pcq else torch.tensor(float(zf), device="cuda")
    bt = dev(b) if bias else None
    y = ops.ci_forward(xt, bt, s_t, z_t, spec)
    assert y.stride() == xt.stride()
    # oracle on the pre-activation computed op by op in fp32
    pre = x + b.reshape(1, C, 1, 1) if bias else x
    act = np.where(np.isnan(pre), pre, np.maximum(pre, 0)).astype(np.float32) if relu else pre
    y_o = oracle.fake_quant_fwd(act, s, zf, 0, 255, ch_axis=1 if pcq else None, zp_learned=True)
    assert bits_equal(y.cpu().numpy(), y_o), first_mismatch(y.cpu().numpy(), y_o)
    gs = ops.lsq_grad_scale(255, x.size, C if pcq else 1)
    dx, ds, dz, db = ops.ci_backward(xt, bt, gt, s_t, z_t, spec, gs, None, True, True, True)
    dx_o, ds_o, dz_o = oracle.fake_quant_bwd(act, g, s, zf, 0, 255, ch_axis=1 if pcq else None, zp_learned=True,
                                             grad_scale=gs, want_dz=True)
    if relu:
        dx_o = np.where(pre > 0, dx_o, np.float32(0)).astype(np.float32)
    assert bits_equal(dx.cpu().numpy(), dx_o), first_mismatch(dx.cpu().numpy(), dx_o)
    am = np.moveaxis(act, 1, 0).reshape(C, -1) if pcq else act.reshape(1, -1)
    gm = np.moveaxis(g, 1, 0).reshape(C, -1) if pcq else g.reshape(1, -1)
    sv = np.atleast_1d(s).astype(np.float64)
    zv = np.atleast_1d(zf)
    ds, dz = ds.cpu().numpy().astype(np.float64), dz.cpu().numpy().astype(np.float64)
    for c in range(am.shape[0]):
        zeff = float(np.clip(np.rint(zv[c]), 0, 255))
        mass = gs * sum_mass(am[c], gm[c], sv[c], zeff, 0, 255)
        assert abs(ds[c] - ds_o[c]) <= 2e-6 * mass + 1e-30, (c, ds[c], ds_o[c])
        zmass = gs * float(np.sum(np.abs(gm[c].astype(np.float64) * sv[c])))
        assert abs(dz[c] - dz_o[c]) <= 2e-6 * zmass + 1e-30, (c, dz[c], dz_o[c])
    if bias:
        db_o = np.moveaxis(dx_o.astype(np.float64), 1, 0).reshape(C, -1).sum(1)
        db_mass = np.moveaxis(np.abs(dx_o.astype(np.float64)), 1, 0).reshape(C, -1).sum(1)
        assert np.all(np.abs(db.cpu().numpy() - db_o) <= 2e-6 * db_mass + 1e-30)
    # STE-only form (no dscale) gives the same dx and dbias
    dx2, ds2, _, db2 = ops.ci_backward(xt, bt, gt, s_t, z_t, spec, want_ds=False, want_dbias=True)
    assert ds2 is None and bits_equal(dx2.cpu().numpy(), dx_o)
    if bias:
        assert torch.equal(db2, db)


def test_channels_inner_special_values_and_determinism(ops):
    x = torch.randn(4, 32, 17, 13, device="cuda")
    x.view(-1)[:9] = torch.tensor([0.0, -0.0, float("nan"), 1e-40, -1e-40, float("inf"), -float("inf"), 3e38, 1e-30], device="cuda")
    x = x.contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(x)
    b = torch.randn(32, device="cuda")
    s = torch.full((1, 32, 1, 1), 0.02, device="cuda")
    z = torch.full((1, 32, 1, 1), 7.6, device="cuda")
    spec = ops.QSpec(0, 255, ch_axis=1, zp_learned=True, pre_relu=True)
    y = ops.ci_forward(x, b, s, z, spec)
    xn = x.cpu().numpy() + b.cpu().numpy().reshape(1, 32, 1, 1)
    act = np.where(np.isnan(xn), xn, np.maximum(xn, 0)).astype(np.float32)
    y_o = oracle.fake_quant_fwd(act, s.cpu().numpy().reshape(-1), z.cpu().numpy().reshape(-1), 0, 255, ch_axis=1, zp_learned=True)
    assert bits_equal(y.cpu().numpy(), y_o), first_mismatch(y.cpu().numpy(), y_o)
    a = ops.ci_backward(x, b, g, s, z, spec, 0.01, None, True, True, True)
    c = ops.ci_backward(x, b, g, s, z, spec, 0.01, None, True, True, True)
    for u, w in zip(a, c):  # dynamic tile scheduling, yet bit-reproducible: fixed-order combination
        assert torch.equal(torch.nan_to_num(u), torch.nan_to_num(w))


def test_channels_inner_strided_grad_output(ops):
    """grad_output that is a channel slice of a wider NHWC tensor (what torch.cat's backward hands out) is read in place."""
    torch.manual_seed(2)
    x = torch.randn(3, 32, 11, 9, device="cuda").contiguous(memory_format=torch.channels_last)
    wide = torch.randn(3, 96, 11, 9, device="cuda").contiguous(memory_format=torch.channels_last)
    g_view = wide[:, 32:64]
    assert g_view.stride() != x.stride() and g_view.stride(1) == 1
    b = torch.randn(32, device="cuda")
    s = torch.full((1, 32, 1, 1), 0.02, device="cuda")
    z = torch.full((1, 32, 1, 1), 3.3, device="cuda")
    spec = ops.QSpec(0, 255, ch_axis=1, zp_learned=True, pre_relu=True)
    a = ops.ci_backward(x, b, g_view, s, z, spec, 0.01, None, True, True, True)
    c = ops.ci_backward(x, b, g_view.contiguous(memory_format=torch.channels_last), s, z, spec, 0.01, None, True, True, True)
    for u, w in zip(a, c):
        assert torch.equal(u, w)


@pytest.mark.parametrize("shape", [(2, 8, 5, 7), (3, 64, 33, 17), (2, 300, 9, 9), (1, 1024, 3, 5), (64, 32, 40, 40)])
def test_ci_observe_matches_nchw_observer(shape):
    """vsiq_ci_observe (per-channel statistics on channels_last memory) against vsiq_observe on the NCHW tensor and the
    oracle: extrema bit-exact (NaN channel included), sums to 2e-6 relative (both kernels add <= 16 elements in fp32
    before the fp64 accumulation, in different groupings; the BN-moment bar is 1e-5)."""
    from vsiquantization_b200 import ops
    torch.manual_seed(11)
    x = torch.randn(shape, device="cuda") * 3 + 0.5
    if shape[1] >= 8:
        x[0, 3, 0, 0] = float("nan")
    xc = x.contiguous(memory_format=torch.channels_last)
    assert ops.ci_supported(xc)
    st0 = ops.new_observer_state(shape[1], "cuda")
    st1 = st0.clone()
    a = ops.observe(x, ch_axis=1, state=st0, bits=8, symmetric=False)
    l0 = ops._lib.launch_count
    b = ops.observe(xc, ch_axis=1, state=st1, bits=8, symmetric=False)
    assert ops._lib.launch_count - l0 == 2  # the channel-innermost pass + its combine; no layout conversion
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert np.array_equal(a[:, :2], b[:, :2], equal_nan=True)
    mass = np.abs(x.cpu().numpy()).sum(axis=(0, 2, 3))[:, None]
    np.testing.assert_allclose(b[:, 2:4], a[:, 2:4], rtol=0, atol=float(2e-6 * np.nanmax(mass)))
    np.testing.assert_allclose(b[:, 4], a[:, 4], rtol=2e-6)
    so = oracle.minmax_stats(x.cpu().numpy(), ch_axis=1)
    assert np.array_equal(b[:, :2], so[:, :2], equal_nan=True)
    s0, s1 = st0.cpu().numpy(), st1.cpu().numpy()
    assert np.array_equal(s0[:, :5], s1[:, :5], equal_nan=True)  # running extrema, scale, zero-point, call count
    np.testing.assert_allclose(s1[:, 5:], s0[:, 5:], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,qmin,qmax,z", [(8, -128, 127, 0.0), (4, 0, 15, 8.0), (16, -32768, 32767, 0.0)])
def test_integer_code_export_long_tiles(ops, bits, qmin, qmax, z):
    """Large tensors walk tiles of up to 8 batches (reduce_tile_mult) and take the code straight from the mantissa of
    t + 1.5 * 2^23: every code in the range, ties, specials and a ragged tail against the oracle, codes only and with y."""
    rng = np.random.default_rng(bits)
    n = (1 << 24) + 40
    s = np.float32(6.0 / (qmax - qmin))
    x = (rng.standard_normal(n) * 2).astype(np.float32)
    k = np.arange(qmin - 3, qmax + 4, dtype=np.float64)
    ties = np.concatenate([(k - z) * s, (k + 0.5 - z) * s]).astype(np.float32)  # exact codes and half-way points
    x[1000:1000 + ties.size] = ties
    x[::9973] = np.nan
    x[1::9967] = np.inf
    x[2::9949] = -1e30
    x[3::9941] = 1e-42
    x[4::9931] = -0.0
    y_o, c_o = oracle.fake_quant_fwd(x, s, z, qmin, qmax, want_codes=True)
    want = np.where(np.isnan(c_o), 0, c_o).astype(np.int64)
    spec = ops.QSpec(qmin, qmax)
    xt = dev(x)
    for want_y in (False, True):
        y, codes = ops.quantize_codes(xt, float(s), float(z), spec, bits, want_y=want_y)
        c = codes.cpu().numpy()
        if bits == 4:
            got = np.stack([(c & 0xf), (c >> 4)], axis=-1).reshape(-1).astype(np.int64)
        else:
            got = c.astype(np.int64)
        assert np.array_equal(got, want), (bits, want_y, np.flatnonzero(got != want)[:5])
        if want_y:
            assert bits_equal(y.cpu().numpy(), y_o)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,sym", [(8, True), (8, False), (4, True), (4, False), (16, True), (12, False), (3, True)])
@pytest.mark.parametrize("shape,ch_axis", [(((1 << 16) + 24,), None), ((37, 1030), 0), ((3, 6, 50, 38), 1), ((5, 8), None)])
def test_integer_code_export(ops, bits, sym, shape, ch_axis):
    """vsiq_quantize_codes: codes equal the oracle's x_int exactly in every width (int4 packed two per byte, int8, int16),
    signed and unsigned, per tensor and per channel, ragged rows; y from the same pass is the forward's y bit for bit;
    wider-than-8-bit ranges are never wrapped into bytes."""
    rng = np.random.default_rng(bits * 7 + len(shape))
    x = (rng.standard_normal(shape) * 3).astype(np.float32)
    x.reshape(-1)[::97] = np.nan
    x.reshape(-1)[1::113] = 1e30
    qmin, qmax = (-(1 << (bits - 1)), (1 << (bits - 1)) - 1) if sym else (0, (1 << bits) - 1)
    C = 1 if ch_axis is None else shape[ch_axis]
    s = (rng.uniform(0.5, 2.0, C) * 6.0 / (qmax - qmin)).astype(np.float32)
    z = np.zeros(C, np.float32) if sym else np.rint(rng.uniform(qmax / 3, qmax / 2, C)).astype(np.float32)
    spec = ops.QSpec(qmin, qmax, ch_axis=ch_axis)
    xt = dev(x)
    st, zt = (float(s[0]), float(z[0])) if ch_axis is None else (dev(s), dev(z))
    y_o, c_o = oracle.fake_quant_fwd(x, s if ch_axis is not None else s[0], z if ch_axis is not None else z[0], qmin, qmax,
                                     ch_axis=ch_axis, want_codes=True)
    want = np.where(np.isnan(c_o), 0, c_o).astype(np.int64)
    for code_bits in sorted({4 if bits <= 4 else None, 8 if bits <= 8 else None, 16} - {None}):
        y, codes = ops.quantize_codes(xt, st, zt, spec, code_bits)
        assert bits_equal(y.cpu().numpy(), y_o), (code_bits, first_mismatch(y.cpu().numpy(), y_o))
        c = codes.cpu().numpy()
        if code_bits == 4:
            assert c.dtype == np.uint8 and c.shape == shape[:-1] + (shape[-1] // 2,)
            lo, hi = (c & 0xf).astype(np.int64), (c >> 4).astype(np.int64)
            if sym:
                lo, hi = np.where(lo > 7, lo - 16, lo), np.where(hi > 7, hi - 16, hi)
            got = np.stack([lo, hi], axis=-1).reshape(shape)
        else:
            got = c.astype(np.int64)
            assert c.shape == shape and c.dtype.itemsize * 8 == code_bits and (c.dtype.kind == "i") == sym
        assert np.array_equal(got, want), code_bits
        # the codes-only instantiation (no y written: its own kernel) emits the same bytes
        none, codes_only = ops.quantize_codes(xt, st, zt, spec, code_bits, want_y=False)
        assert none is None and np.array_equal(codes_only.cpu().numpy(), c), code_bits
    # the default width never wraps: 8 bits when the range fits, else 16
    y2, c2 = ops.fake_quant_forward(xt, st, zt, spec, want_codes=True)
    assert c2.dtype.itemsize == (1 if bits <= 8 else 2) and np.array_equal(c2.cpu().numpy().astype(np.int64), want)
    if bits > 8:
        with pytest.raises(Exception):
            ops.quantize_codes(xt, st, zt, spec, 8)
