"""CPU oracle for the fake-quantization hot path -- TEST INFRASTRUCTURE ONLY.

NumPy-facing wrapper around ``oracle/vsiq_oracle.c`` (a plain-C restatement of
the reference's algorithm; every function there cites the reference file:line
it follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only
as the checker or the timed CPU baseline.  The product package
``vsiquantization_b200`` never imports it.

Parity pin: the reference has no tests or golden vectors of its own; the
oracle is pinned against outputs of the reference's own Python code run on CPU
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``, checked by
``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvsiq_oracle.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc, seconds).  Returns the path of the .so."""
    src = os.path.join(_HERE, "vsiq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.vsiq_oracle_fake_quant_fwd.argtypes = [
            _f32p, _f32p, _f32p, _i64, _i64, _i64, _f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.vsiq_oracle_fake_quant_fwd.restype = None
        L.vsiq_oracle_fake_quant_bwd.argtypes = [
            _f32p, _f32p, _f32p, _f64p, _f64p, _i64, _i64, _i64, _f32p, _f32p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, ctypes.c_int]
        L.vsiq_oracle_fake_quant_bwd.restype = None
        L.vsiq_oracle_minmax_stats.argtypes = [_f32p, _i64, _i64, _i64, _f64p]
        L.vsiq_oracle_minmax_stats.restype = None
        L.vsiq_oracle_minmax_update.argtypes = [_f64p, _f64p, ctypes.c_double, ctypes.c_double]
        L.vsiq_oracle_minmax_update.restype = None
        L.vsiq_oracle_qparams.argtypes = [
            ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_double, _f64p, _f64p]
        L.vsiq_oracle_qparams.restype = None
        L.vsiq_oracle_lsq_init_scale.argtypes = [_f64p, _i64, ctypes.c_int]
        L.vsiq_oracle_lsq_init_scale.restype = ctypes.c_double
        L.vsiq_oracle_grad_scale.argtypes = [ctypes.c_int, _i64, _i64]
        L.vsiq_oracle_grad_scale.restype = ctypes.c_double
        L.vsiq_oracle_bn_fold.argtypes = [
            _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, ctypes.c_float, _i64, _i64, _f32p, _f32p]
        L.vsiq_oracle_bn_fold.restype = None
        L.vsiq_oracle_bn_moments.argtypes = [_f32p, _i64, _i64, _i64, _f64p, _f64p, _f64p]
        L.vsiq_oracle_bn_moments.restype = None
        L.vsiq_oracle_bn_reestimate_accumulate.argtypes = [_f32p, _f32p, _f32p, _f32p, _i64]
        L.vsiq_oracle_bn_reestimate_accumulate.restype = None
        L.vsiq_oracle_bn_reestimate_finish.argtypes = [_f32p, _f32p, _i64, _f32p, _f32p, _i64]
        L.vsiq_oracle_bn_reestimate_finish.restype = None
        L.vsiq_oracle_fake_quant_fwd_bwd.argtypes = [
            _f32p, _f32p, _f32p, _f32p, _i64, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int]
        L.vsiq_oracle_fake_quant_fwd_bwd.restype = None
        L.vsiq_oracle_set_threads.argtypes = [ctypes.c_int]
        L.vsiq_oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p32(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_f32p)


def _p64(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_f64p)


def _c32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def layout(shape, ch_axis: Optional[int]) -> Tuple[int, int, int]:
    """(outer, C, inner) of a contiguous tensor quantised along ``ch_axis`` (None = per tensor)."""
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if ch_axis is None:
        return 1, 1, n
    outer = int(np.prod(shape[:ch_axis], dtype=np.int64)) if ch_axis > 0 else 1
    C = int(shape[ch_axis])
    inner = n // (outer * C) if outer * C else 0
    return outer, C, inner


def _qp(v, C) -> np.ndarray:
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size == 1 and C > 1:
        a = np.repeat(a, C)
    assert a.size == C, (a.size, C)
    return np.ascontiguousarray(a.astype(np.float32))  # reference rounds scale/zp to fp32 before use


def set_threads(n: int) -> None:
    lib().vsiq_oracle_set_threads(int(n))


def max_threads() -> int:
    return int(lib().vsiq_oracle_max_threads())


def fake_quant_fwd(x, scale, zero_point, qmin, qmax, ch_axis=None, zp_learned=False, want_codes=False):
    """y (and fp32 codes) of quantizers/uniform.py:54-55,95 / lsq_module.py:254-274."""
    x = _c32(x)
    outer, C, inner = layout(x.shape, ch_axis)
    s, z = _qp(scale, C), _qp(zero_point, C)
    y = np.empty_like(x)
    codes = np.empty_like(x) if want_codes else None
    lib().vsiq_oracle_fake_quant_fwd(_p32(x), _p32(y), _p32(codes), outer, C, inner, _p32(s), _p32(z),
                                     int(bool(zp_learned)), int(qmin), int(qmax))
    return (y, codes) if want_codes else y


def fake_quant_bwd(x, g, scale, zero_point, qmin, qmax, ch_axis=None, zp_learned=False,
                   grad_scale=None, want_ds=True, want_dz=False, mask_mode=0):
    """dx (fp32, op-for-op), ds[C], dz[C] (double accumulation) -- see vsiq_oracle.c."""
    x, g = _c32(x), _c32(g)
    outer, C, inner = layout(x.shape, ch_axis)
    s, z = _qp(scale, C), _qp(zero_point, C)
    dx = np.empty_like(x)
    ds = np.zeros(C, dtype=np.float64) if want_ds else None
    dz = np.zeros(C, dtype=np.float64) if want_dz else None
    gs = None
    if grad_scale is not None:
        gs = np.asarray(grad_scale, dtype=np.float64).reshape(-1)
        if gs.size == 1 and C > 1:
            gs = np.repeat(gs, C)
        gs = np.ascontiguousarray(gs)
    lib().vsiq_oracle_fake_quant_bwd(_p32(x), _p32(g), _p32(dx), _p64(ds), _p64(dz), outer, C, inner,
                                     _p32(s), _p32(z), int(bool(zp_learned)), int(qmin), int(qmax),
                                     _p64(gs), int(mask_mode))
    return dx, ds, dz


def minmax_stats(x, ch_axis=None) -> np.ndarray:
    """[C,5] = min, max, sum|x|, sum x, sum x^2 of one observer call."""
    x = _c32(x)
    outer, C, inner = layout(x.shape, ch_axis)
    out = np.zeros((C, 5), dtype=np.float64)
    lib().vsiq_oracle_minmax_stats(_p32(x), outer, C, inner, _p64(out))
    return out


def minmax_update(run_min: float, run_max: float, call_min: float, call_max: float):
    a, b = ctypes.c_double(run_min), ctypes.c_double(run_max)
    lib().vsiq_oracle_minmax_update(ctypes.byref(a), ctypes.byref(b), float(call_min), float(call_max))
    return a.value, b.value


def qparams(mn: float, mx: float, bits: int = 8, symmetric: bool = True, eps: float = 1e-8):
    s, z = ctypes.c_double(), ctypes.c_double()
    lib().vsiq_oracle_qparams(float(mn), float(mx), int(bits), int(bool(symmetric)), float(eps),
                              ctypes.byref(s), ctypes.byref(z))
    return s.value, z.value


def lsq_init_scale(mean_abs, bits: int) -> float:
    a = np.ascontiguousarray(np.asarray(mean_abs, dtype=np.float64).reshape(-1))
    return float(lib().vsiq_oracle_lsq_init_scale(_p64(a), a.size, int(bits)))


def grad_scale(qmax: int, numel: int, C: int = 1) -> float:
    return float(lib().vsiq_oracle_grad_scale(int(qmax), int(numel), int(C)))


def bn_fold(W, bias, gamma, beta, mean, var, eps):
    W = _c32(W)
    C = W.shape[0]
    inner = W.size // C
    bias_a = None if bias is None else _c32(bias)
    gamma, beta, mean, var = _c32(gamma), _c32(beta), _c32(mean), _c32(var)
    Wo = np.empty_like(W)
    bo = np.empty(C, dtype=np.float32)
    lib().vsiq_oracle_bn_fold(_p32(W), _p32(bias_a), _p32(gamma), _p32(beta), _p32(mean), _p32(var),
                              ctypes.c_float(eps), C, inner, _p32(Wo), _p32(bo))
    return Wo, bo


def bn_moments(x):
    """Per-channel (mean, unbiased var, biased var) of an [N,C,...] batch, in double."""
    x = _c32(x)
    N, C = x.shape[0], x.shape[1]
    HW = x.size // (N * C)
    m = np.zeros(C, dtype=np.float64)
    vu = np.zeros(C, dtype=np.float64)
    vb = np.zeros(C, dtype=np.float64)
    lib().vsiq_oracle_bn_moments(_p32(x), N, C, HW, _p64(m), _p64(vu), _p64(vb))
    return m, vu, vb


def bn_reestimate(batches):
    """utils/estimate_bn.py:56-99 over an iterable of [N,C,H,W] conv outputs -> (running_mean, running_var) fp32."""
    mean_sum = var_sum = None
    k = 0
    for x in batches:
        m, vu, _ = bn_moments(x)
        m32, v32 = m.astype(np.float32), vu.astype(np.float32)
        if mean_sum is None:
            mean_sum = np.zeros_like(m32)
            var_sum = np.zeros_like(v32)
        lib().vsiq_oracle_bn_reestimate_accumulate(_p32(mean_sum), _p32(var_sum), _p32(m32), _p32(v32), m32.size)
        k += 1
    rm = np.empty_like(mean_sum)
    rv = np.empty_like(var_sum)
    lib().vsiq_oracle_bn_reestimate_finish(_p32(mean_sum), _p32(var_sum), k, _p32(rm), _p32(rv), rm.size)
    return rm, rv


def fake_quant_fwd_bwd(x, g, scale, zero_point, qmin, qmax, y=None, dx=None):
    """Single-sweep fwd + STE bwd (per tensor) -- the C port timed as a CPU baseline."""
    x, g = _c32(x), _c32(g)
    y = np.empty_like(x) if y is None else y
    dx = np.empty_like(x) if dx is None else dx
    lib().vsiq_oracle_fake_quant_fwd_bwd(_p32(x), _p32(g), _p32(y), _p32(dx), x.size,
                                         ctypes.c_float(np.float32(scale)), ctypes.c_float(np.float32(zero_point)),
                                         int(qmin), int(qmax))
    return y, dx


_proof = None


def proof_division(scale: float, mode: int = 0, first: int = 0, stride: int = 1):
    """(mismatches, inputs on the fast path) of the kernels' division-free arithmetic against IEEE division over the bit
    patterns first, first+stride, ... of the input, emulated on the host (oracle/fastpath_proof.c).  mode 0: x / s;
    mode 1: RN(RN(g*s) / s)."""
    global _proof
    if _proof is None:
        so = os.path.join(_HERE, "_build", "libvsiq_fastpath_proof.so")
        src = os.path.join(_HERE, "fastpath_proof.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE, "-s"], check=True)
        _proof = ctypes.CDLL(so)
        _proof.vsiq_proof_division.restype = ctypes.c_ulonglong
        _proof.vsiq_proof_division.argtypes = [ctypes.c_float, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                               ctypes.POINTER(ctypes.c_ulonglong)]
    covered = ctypes.c_ulonglong(0)
    wrong = _proof.vsiq_proof_division(float(scale), int(mode), int(first), int(stride), ctypes.byref(covered))
    return int(wrong), int(covered.value)


def proof_code_magic(lo: float, hi: float, bits: int):
    """(mismatches, floats visited) of the code-export conversion -- low `bits` mantissa bits of RN(t + 1.5 * 2^23)
    against (int)rint(t) -- over EVERY float32 in [lo, hi], emulated on the host (oracle/fastpath_proof.c)."""
    proof_division(1.0, 0, 0, 1 << 31)  # loads (and if needed builds) the library
    _proof.vsiq_proof_code_magic.restype = ctypes.c_ulonglong
    _proof.vsiq_proof_code_magic.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
    covered = ctypes.c_ulonglong(0)
    wrong = _proof.vsiq_proof_code_magic(float(lo), float(hi), int(bits), ctypes.byref(covered))
    return int(wrong), int(covered.value)
