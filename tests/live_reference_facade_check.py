"""CPU check of vsiquantization_b200.quantizers.lsq_module.LSQFakeQuantize against the LIVE reference's LSQFakeQuantize
(quantizers/lsq_module.py:73-173) -- build container only; run as a subprocess by tests/test_abi_and_host.py.

The module's two kernel entry points (the observer pass and LSQQuantizer.quantize) are replaced by oracle-backed stand-ins
(this host has no GPU; the kernels behind them have their own parity tests), so what is compared is everything the module
itself does: phases, buffers, lazily created parameters and their shapes, state_dict keys, the gradient scale handed to
the quantiser (x 5000 for activations), outputs and gradients -- per tensor / per channel x symmetric / affine x weight /
activation."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import ref_shim
ref_shim.install()
from quantizers.lsq_module import LSQFakeQuantize as RefFQ
from torch.quantization import MovingAveragePerChannelMinMaxObserver, MovingAverageMinMaxObserver
from vsiquantization_b200.quantizers.lsq_module import LSQFakeQuantize as OurFQ


class OracleFQ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, zp, qmin, qmax, ch_axis, zp_learned, gs):
        ctx.args = (qmin, qmax, ch_axis, zp_learned, gs)
        ctx.save_for_backward(x, scale, zp)
        y = oracle.fake_quant_fwd(x.detach().numpy(), scale.detach().numpy().reshape(-1).astype(np.float64),
                                  zp.detach().numpy().reshape(-1).astype(np.float64), qmin, qmax, ch_axis=ch_axis,
                                  zp_learned=zp_learned)
        return torch.from_numpy(y)

    @staticmethod
    def backward(ctx, g):
        qmin, qmax, ch_axis, zp_learned, gs = ctx.args
        x, scale, zp = ctx.saved_tensors
        dx, ds, dz = oracle.fake_quant_bwd(x.detach().numpy(), g.numpy(), scale.detach().numpy().reshape(-1).astype(np.float64),
                                           zp.detach().numpy().reshape(-1).astype(np.float64), qmin, qmax, ch_axis=ch_axis,
                                           zp_learned=zp_learned, grad_scale=gs, want_ds=True, want_dz=True)
        return (torch.from_numpy(dx), torch.from_numpy(ds).to(scale.dtype).view(scale.shape),
                torch.from_numpy(dz).to(zp.dtype).view(zp.shape), None, None, None, None, None)


def stand_in(fq):
    q = fq._quantizer

    def quantize(x, scale, zero_point, is_learning_scale=False):
        ch_axis = q._resolve_axis(x, scale)
        C = scale.numel()
        zp_round = bool(is_learning_scale and zero_point.is_floating_point())
        gs = q.calculate_grad_scale(x, C) * q.grad_boost if is_learning_scale else 1.0
        return OracleFQ.apply(x, scale, zero_point, q.qmin, q.qmax, ch_axis, zp_round, gs)
    q.quantize = quantize

    def extrema(X):
        st = oracle.minmax_stats(X.detach().numpy(), ch_axis=1 if fq.is_per_channel else None)
        return torch.from_numpy(st[:, 0]).float(), torch.from_numpy(st[:, 1]).float()
    fq._batch_extrema = extrema


def run(per_channel, affine, config_act, seed):
    rng = np.random.default_rng(seed)
    kw = dict(quant_min=0 if affine else -128, quant_max=255 if affine else 127,
              dtype=torch.quint8 if affine else torch.qint8)
    if per_channel:
        kw.update(observer=MovingAveragePerChannelMinMaxObserver, ch_axis=1,
                  qscheme=torch.per_channel_affine if affine else torch.per_channel_symmetric)
    else:
        kw.update(observer=MovingAverageMinMaxObserver,
                  qscheme=torch.per_tensor_affine if affine else torch.per_tensor_symmetric)
    a, b = RefFQ(learn_scale=True, config_act=config_act, **kw), OurFQ(learn_scale=True, config_act=config_act, **kw)
    stand_in(b)
    shape = (2, 5, 4, 3)
    for i in range(3):  # observer phase (fake-quant on, like the reference's defaults)
        x = (rng.standard_normal(shape) * (1 + i) + (0.8 if affine else 0.0)).astype(np.float32)
        ya, yb = a(torch.from_numpy(x.copy())), b(torch.from_numpy(x.copy()))
        assert torch.equal(ya, yb), ("observer-phase output", per_channel, affine, i)
        assert torch.equal(a.scale, b.scale) and torch.equal(a.zero_point, b.zero_point), ("buffers", i)
        assert torch.equal(a.activation_post_process.min_val, b.activation_post_process.min_val)
        assert torch.equal(a.scale_param.detach(), b.scale_param.detach()), i
        assert torch.equal(a.zero_point_param_float.detach(), b.zero_point_param_float.detach()), i
    assert set(a.state_dict()) == set(b.state_dict()), (set(a.state_dict()) ^ set(b.state_dict()))
    a.disable_observer(); b.disable_observer()
    for i in range(2):  # learn phase
        x = (rng.standard_normal(shape) * 2.5 + (0.8 if affine else 0.0)).astype(np.float32)
        g = rng.standard_normal(shape).astype(np.float32)
        outs = []
        for m in (a, b):
            m.zero_grad()
            xt = torch.from_numpy(x.copy()).requires_grad_(True)
            y = m(xt)
            y.backward(torch.from_numpy(g))
            outs.append((y.detach(), xt.grad, m.scale_param.grad.clone(), m.zero_point_param_float.grad.clone()))
        (ya, dxa, dsa, dza), (yb, dxb, dsb, dzb) = outs
        assert torch.equal(ya, yb) and torch.equal(dxa, dxb), ("learn y/dx", per_channel, affine, i)
        assert dsa.shape == dsb.shape and dza.shape == dzb.shape and dsa.dtype == dsb.dtype
        tol = 2e-5 * float(dsa.abs().max() + dsb.abs().max() + 1e-30) + 1e-4 * float(np.abs(g).sum()) * b.calculate_grad_scale(torch.from_numpy(x)) * (5000 if config_act else 1) * 1e-2
        assert torch.allclose(dsa, dsb, rtol=1e-3, atol=tol), ("ds", dsa.flatten()[:4], dsb.flatten()[:4])
        assert torch.allclose(dza, dzb, rtol=1e-3, atol=tol), ("dz", dza.flatten()[:4], dzb.flatten()[:4])
    # fixed phase: learn_scale off -> buffers as constants
    a.learn_scale = b.learn_scale = False
    x = (rng.standard_normal(shape) * 2).astype(np.float32)
    assert torch.equal(a(torch.from_numpy(x.copy())), b(torch.from_numpy(x.copy())))


if __name__ == "__main__":
    n = 0
    for per_channel in (False, True):
        for affine in (False, True):
            for config_act in (False, True):
                run(per_channel, affine, config_act, 100 + n)
                n += 1
    print("facade matches the live reference in", n, "configurations")
