"""Calibration forward of a fused layer as ONE pass over the conv output: bias add (or the inference-mode BatchNorm the
layer kept) + ReLU / SiLU + the output observer (vsiq_ci_epilogue_observe).  Reference: modules/fused.py:124-134, then
quantizers/quantization_manager.py:55-71 -> observers/minmax.py:32-47.  The written tensor and the running min / max /
scale / zero-point must equal the separate passes bit for bit; the three sums are fp64 sums of fp32 partials in another
order (1e-6 relative, the bar the all-reduced LSQ-init statistics already have)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _close(a, b, rel=1e-6):
    a, b = a.double().cpu(), b.double().cpu()
    return bool(((a - b).abs() <= rel * b.abs().clamp_min(1e-30)).all())


@pytest.mark.parametrize("shape", [(2, 8, 5, 7), (3, 24, 9, 7), (4, 64, 20, 20), (2, 512, 6, 6), (1, 4, 1, 3)])
@pytest.mark.parametrize("act", [None, "relu", "silu"])
@pytest.mark.parametrize("pre", ["none", "bias", "bn"])
def test_epilogue_observe_equals_separate_passes(shape, act, pre):
    from vsiquantization_b200 import _lib, ops
    torch.manual_seed(3)
    N, C, H, W = shape
    x = _cl(torch.randn(shape, device="cuda") * 1.7)
    bias = torch.randn(C, device="cuda") * 0.3
    mean, var = torch.randn(C, device="cuda") * 0.2, torch.rand(C, device="cuda") + 0.5
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
    fn = {None: lambda t: t, "relu": torch.relu, "silu": torch.nn.functional.silu}[act]
    if pre == "bias":
        want = fn(x + bias.view(1, -1, 1, 1))
        kw = {"bias": bias}
    elif pre == "bn":
        want = fn(ops.ci_bn_normalize(x, mean, var, gamma, beta, 1e-3, relu=False)) if ops.ci_supported(x) else None
        kw = {"bn": (mean, var, gamma, beta, 1e-3)}
    else:
        want = fn(x)
        kw = {}
    for symmetric in (True, False):
        st_f = ops.new_observer_state(1, x.device)
        st_s = ops.new_observer_state(1, x.device)
        n0 = _lib.launch_count
        y, stats = ops.ci_epilogue_observe(x, st_f, 8, symmetric, 1e-8, act, **kw)
        assert _lib.launch_count - n0 == 1
        assert y.stride() == x.stride() and torch.equal(y, want)
        stats_s = ops.observe(want, None, st_s, 8, symmetric, 1e-8)
        assert torch.equal(stats[:, :2], stats_s[:, :2])                  # min, max
        assert _close(stats[:, 2:5], stats_s[:, 2:5]), (stats, stats_s)  # sum|x|, sum x, sum x^2
        assert torch.equal(st_f[:, :5], st_s[:, :5])                      # running min / max, scale, zero-point, n_calls
        assert _close(st_f[:, 5:], st_s[:, 5:], 2e-6)
        # against the CPU oracle's observer over the written tensor
        ref = oracle.minmax_stats(want.permute(0, 2, 3, 1).contiguous().cpu().numpy().reshape(-1))
        got = stats.cpu().numpy().reshape(-1)
        assert got[0] == ref.reshape(-1)[0] and got[1] == ref.reshape(-1)[1]
        # a second call keeps accumulating like the separate observer does
        x2 = _cl(x * 2.0)
        y2, _ = ops.ci_epilogue_observe(x2, st_f, 8, symmetric, 1e-8, act, **kw)
        ops.observe(y2, None, st_s, 8, symmetric, 1e-8)
        assert torch.equal(st_f[:, :5], st_s[:, :5]) and float(st_f[0, 4]) == 2.0


def test_nan_never_updates_the_running_extrema():
    from vsiquantization_b200 import ops
    x = _cl(torch.randn(2, 8, 4, 4, device="cuda"))
    st = ops.new_observer_state(1, x.device)
    ops.ci_epilogue_observe(x, st, 8, True, 1e-8, None)
    before = st.clone()
    x2 = _cl(x * 3)
    x2.permute(0, 2, 3, 1).reshape(-1)[5] = float("nan")   # a view of the channels_last memory
    assert bool(torch.isnan(x2).any())
    y, stats = ops.ci_epilogue_observe(x2, st, 8, True, 1e-8, "relu")
    assert torch.isnan(stats[0, 0]) and torch.isnan(stats[0, 1]) and bool(torch.isnan(y).any())
    assert torch.equal(st[:, :2], before[:, :2])     # observers/minmax.py:44-47: NaN compares false


def _layer(is_fuse_bn, act_cls, cout=16):
    from vsiquantization_b200.modules.fused import ConvBnReLU
    cv, bn = torch.nn.Conv2d(8, cout, 3, padding=1, bias=False), torch.nn.BatchNorm2d(cout, eps=1e-3)
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    return ConvBnReLU(cv, bn, act_cls(), "LSQObserver", "LSQQuantizer", "LSQObserver", "LSQQuantizer", True, False,
                      is_fuse_bn, 4, 8).cuda().to(memory_format=torch.channels_last)


@pytest.mark.parametrize("is_fuse_bn", [True, False], ids=["bn-folded", "bn-kept"])
@pytest.mark.parametrize("act_cls", [torch.nn.ReLU, torch.nn.SiLU], ids=["relu", "silu"])
def test_calibrating_layer_is_one_epilogue_pass_and_matches_the_separate_passes(is_fuse_bn, act_cls, monkeypatch):
    import copy
    from vsiquantization_b200 import _lib
    from vsiquantization_b200.utils.quantize_manager import activate_learning_qparam, calibrate_qat_model
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(9)
    fused = _layer(is_fuse_bn, act_cls)
    plain = copy.deepcopy(fused)
    plain.fuse_observer_into_epilogue = False
    batches = [_cl(torch.randn(4, 8, 20, 20, device="cuda") * (1 + i)) for i in range(3)]

    def calib(m, loader, dev):
        with torch.no_grad():
            return [m(b) for b in loader]

    n0 = _lib.launch_count
    calibrate_qat_model(fused, batches, calib)
    n_fused = _lib.launch_count - n0
    calibrate_qat_model(plain, batches, calib)
    n_plain = _lib.launch_count - n0 - n_fused
    assert n_fused == 3 * 2, n_fused                 # per batch: weight observer + the epilogue pass
    # the separate passes: weight observer + output observer, plus vsiq_ci_bn_normalize for a kept BN + ReLU; the
    # bias add / batch_norm / activation passes of the other forms are ATen's and not counted here
    assert n_plain == 3 * (3 if (not is_fuse_bn and act_cls is torch.nn.ReLU) else 2), n_plain
    with torch.no_grad():
        y_f, y_p = fused(batches[0]), plain(batches[0])
    if is_fuse_bn or act_cls is torch.nn.ReLU:
        assert torch.equal(y_f, y_p)                 # bias + act: ATen's values; kept BN + ReLU: vsiq_ci_bn_normalize's
    else:
        assert torch.allclose(y_f, y_p, rtol=2e-6, atol=2e-6)   # ATen's batch_norm vs x * a[c] + b[c]
    sf, sp = fused.activation_quantizer.observer.state, plain.activation_quantizer.observer.state
    if is_fuse_bn or act_cls is torch.nn.ReLU:
        assert torch.equal(sf[:, :5], sp[:, :5])     # post-calibration min / max / scale / zero-point: bit-exact
    else:
        assert _close(sf[:, :4], sp[:, :4], 1e-5)
    assert _close(sf[:, 5:], sp[:, 5:], 1e-5 if not is_fuse_bn else 2e-6)
    activate_learning_qparam(fused, use_init=True)
    activate_learning_qparam(plain, use_init=True)
    assert _close(fused.activation_quantizer.scale.detach(), plain.activation_quantizer.scale.detach(), 1e-5)
    # with autograd on (the reference's data_calib, yolov8_qat.py:42-52, has no no_grad) the same pass runs behind an
    # autograd node; a backward through it -- nobody's hot path -- gives the ATen composition's gradients
    twin = _layer(is_fuse_bn, act_cls)
    ref = copy.deepcopy(twin)
    ref.fuse_observer_into_epilogue = False
    grads = []
    for m in (twin, ref):
        outs = []
        n0 = _lib.launch_count
        calibrate_qat_model(m, batches[:1], lambda mm, loader, dev: outs.extend(mm(b) for b in loader))
        if m is twin:
            assert _lib.launch_count - n0 == 2
        assert outs[0].requires_grad
        (outs[0] * torch.linspace(-1, 1, outs[0].numel(), device="cuda").view_as(outs[0])).sum().backward()
        grads.append({n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None})
    assert set(grads[0]) == set(grads[1]) and len(grads[0]) >= 2
    for n in grads[1]:
        assert torch.allclose(grads[0][n], grads[1][n], rtol=1e-5, atol=1e-6 * float(grads[1][n].abs().max())), n
    st_t, st_r = twin.activation_quantizer.observer.state, ref.activation_quantizer.observer.state
    if is_fuse_bn:
        assert torch.equal(st_t[:, :4], st_r[:, :4])
    else:   # with autograd on the separate passes use ATen's batch_norm, an ulp away from x * a[c] + b[c]
        assert _close(st_t[:, :4], st_r[:, :4], 1e-5)


def test_bn_reestimation_while_observing_folds_the_observer_into_the_normalise_pass(monkeypatch):
    """reestimate_BN_stats right after calibrate_qat_model (observers still collecting, utils/estimate_bn.py:79-91): per
    layer and batch the moments pass, then normalise + ReLU + output observer as ONE pass -- same running statistics and
    observer state as the separate passes, one launch less."""
    import copy
    from vsiquantization_b200 import _lib
    from vsiquantization_b200.utils.estimate_bn import reestimate_BN_stats
    from vsiquantization_b200.utils.quantize_manager import calibrate_qat_model
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(4)
    fused = _layer(False, torch.nn.ReLU)
    plain = copy.deepcopy(fused)
    plain.fuse_observer_into_epilogue = False
    imgs = [(torch.randint(0, 256, (4, 8, 12, 12), dtype=torch.uint8, device="cuda")
             .contiguous(memory_format=torch.channels_last),) for _ in range(3)]

    def calib(m, loader, dev):
        with torch.no_grad():
            return [m(_cl(b[0].float() / 255.0)) for b in loader]

    counts = []
    for m in (fused, plain):
        calibrate_qat_model(m, imgs, calib)
        n0 = _lib.launch_count
        reestimate_BN_stats(m, imgs, num_batches=3, sync=False)
        counts.append(_lib.launch_count - n0)
    assert counts[1] - counts[0] == 3, counts          # the separate output-observer launch of every batch is gone
    assert torch.equal(fused.bn.running_mean, plain.bn.running_mean)
    assert torch.equal(fused.bn.running_var, plain.bn.running_var)
    sf, sp = fused.activation_quantizer.observer.state, plain.activation_quantizer.observer.state
    assert torch.equal(sf[:, :5], sp[:, :5]) and _close(sf[:, 5:], sp[:, 5:], 2e-6)
