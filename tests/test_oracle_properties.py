"""Property tests (hypothesis) of the oracle's fake-quant semantics -- the size-independent invariants the GPU tests
re-check at full size: idempotence, integer codes inside [qmin, qmax], monotonicity, mask <=> in range, dx in {0, ~g}."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from hypothesis.extra import numpy as hnp

import oracle

finite32 = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False, width=32)
arrays = hnp.arrays(np.float32, st.integers(1, 300), elements=finite32)
scales = st.floats(min_value=float(np.float32(1e-4)), max_value=10.0, allow_nan=False, width=32)
bits = st.integers(2, 8)


def qrange(b, sym):
    return (-(2 ** (b - 1)), 2 ** (b - 1) - 1) if sym else (0, 2 ** b - 1)


@settings(max_examples=150, deadline=None)
@given(arrays, scales, bits, st.booleans(), st.integers(-3, 20))
def test_idempotent_and_codes_in_range(x, s, b, sym, zp):
    qmin, qmax = qrange(b, sym)
    z = 0 if sym else zp
    y, codes = oracle.fake_quant_fwd(x, s, z, qmin, qmax, want_codes=True)
    assert np.all(codes == np.rint(codes)) and codes.min() >= qmin and codes.max() <= qmax
    y2 = oracle.fake_quant_fwd(y, s, z, qmin, qmax)
    assert np.array_equal(y2, y)  # fq(fq(x)) == fq(x)
    # y is exactly (code - z) * s in fp32
    assert np.array_equal(y, ((codes - np.float32(z)) * np.float32(s)).astype(np.float32))


@settings(max_examples=100, deadline=None)
@given(arrays, scales, bits, st.booleans())
def test_monotone(x, s, b, sym):
    qmin, qmax = qrange(b, sym)
    xs = np.sort(x)
    y = oracle.fake_quant_fwd(xs, s, 0, qmin, qmax)
    assert np.all(np.diff(y.astype(np.float64)) >= 0)


@settings(max_examples=100, deadline=None)
@given(arrays, scales, bits, st.booleans())
def test_mask_iff_in_range_and_dx_is_masked_g(x, s, b, sym):
    qmin, qmax = qrange(b, sym)
    g = np.linspace(-1.0, 1.0, x.size).astype(np.float32) + np.float32(0.25)
    dx, ds, _ = oracle.fake_quant_bwd(x, g, s, 0, qmin, qmax, grad_scale=1.0)
    r = np.rint(x.astype(np.float32) / np.float32(s))
    inside = (r >= qmin) & (r <= qmax)
    assert np.all(dx[~inside] == 0)
    np.testing.assert_allclose(dx[inside], g[inside], rtol=2e-7, atol=0)  # ((g*s)/s): within one ulp of g
    # ds is linear in g
    _, ds2, _ = oracle.fake_quant_bwd(x, 2 * g, s, 0, qmin, qmax, grad_scale=1.0)
    assert abs(ds2[0] - 2 * ds[0]) <= 1e-9 * (abs(ds[0]) + 1)


@settings(max_examples=60, deadline=None)
@given(hnp.arrays(np.float32, st.tuples(st.integers(1, 4), st.integers(1, 6), st.integers(1, 9)), elements=finite32), bits)
def test_per_channel_equals_per_tensor_on_each_channel(x, b):
    qmin, qmax = qrange(b, True)
    C = x.shape[1]
    s = (0.01 * (1 + np.arange(C))).astype(np.float32)
    y = oracle.fake_quant_fwd(x, s, np.zeros(C), qmin, qmax, ch_axis=1)
    for c in range(C):
        assert np.array_equal(y[:, c], oracle.fake_quant_fwd(x[:, c], s[c], 0, qmin, qmax))


@settings(max_examples=100, deadline=None)
@given(st.floats(-1e3, 1e3, allow_nan=False), st.floats(-1e3, 1e3, allow_nan=False), bits, st.booleans())
def test_qparams_formula(a, c, b, sym):
    mn, mx = min(a, c, 0.0), max(a, c, 0.0)  # the observer's state always contains 0
    s, z = oracle.qparams(mn, mx, b, sym)
    if sym:
        assert z == 0 and s == max(abs(mn), abs(mx)) / (2 ** (b - 1) - 1 + 1e-8)
    else:
        assert s == (mx - mn) / (2 ** b - 1 + 1e-8) and z == round(-mn / (s + 1e-8))


@pytest.mark.parametrize("scale", [3.0 / 127, 3.0 / 7, 1.0 / 3.0, -0.05, 2.0 ** -40, 2.0 ** 40])
def test_division_free_arithmetic_is_exact_for_every_input(scale):
    """The kernels' hoisted-reciprocal + exact-residual-FMA division (csrc/common.cuh div_fast / dx_fast), emulated on
    the host (oracle/fastpath_proof.c: fmaf and float division are the same IEEE operations as on the GPU), against
    x / s and RN(RN(g*s) / s) for ALL 2^32 bit patterns that take the fast path: zero mismatches.  The GPU repeats this
    on the device (vsiq_selftest_division, tests/test_gpu_kernels.py::test_division_exhaustive)."""
    for mode in (0, 1):
        wrong, covered = oracle.proof_division(scale, mode)
        assert wrong == 0, (scale, mode, wrong)
        assert covered == 2030043138  # 2 * (121 binades * 2^23) + the two zeros


def test_division_free_arithmetic_declines_out_of_range_scales():
    for scale in (2.0 ** -41, 2.0 ** 41, 0.0, float("inf"), float("nan")):
        assert oracle.proof_division(scale, 0, 0, 65537) == (0, 0)  # such tiles always run the IEEE sequence


@pytest.mark.parametrize("qmin,qmax,bits", [(-8, 7, 4), (0, 15, 4), (-128, 127, 8), (0, 255, 8), (-32768, 32767, 16),
                                            (0, 65535, 16), (-2, 1, 4), (0, 3, 8)])
def test_code_export_mantissa_conversion_is_exact_for_every_float(qmin, qmax, bits):
    """The code-export kernel (csrc/fake_quant.cu: codes_vec) reads the integer code from the low mantissa bits of
    RN(t + 1.5 * 2^23) instead of converting rint(t).  Emulated on the host over EVERY float32 t the pre-rounding clamp
    can leave ([qmin - 0.5, qmax + 0.5]: both zeros, all denormals, every tie): identical to (int)rint(t) in two's
    complement, for each code width and range the ABI admits."""
    wrong, covered = oracle.proof_code_magic(qmin - 0.5, qmax + 0.5, bits)
    assert wrong == 0, (qmin, qmax, bits, wrong)
    assert covered > 2 ** 30  # more than a billion floats live in such an interval (denormals and tiny values included)
