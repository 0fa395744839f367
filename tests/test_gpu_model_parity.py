"""Model-scale parity on ONE GPU, bit for bit: the native plugins against the reference's eager composition
(oracle/torch_port.EagerQuantizer -- golden-pinned in test_oracle_golden.py, and the same arithmetic
tests/test_gpu_reference.py checks against the unmodified reference) through the SAME fused layers, cuDNN algorithms and
calibrated parameters.  YOLOv8n, 57 fused layers; BASELINE configs[0]-style (W8A8 symmetric per tensor) and
configs[2]-style (W4A8 per-channel asymmetric LSQ with learnable zero-points); NCHW and channels_last; per-layer launches,
weight bank, CUDA graph.  What must be bit-identical: every fused layer's output (hence every integer code), the input
gradient, every weight gradient; what is a sum (dscale, dzero_point, dbias from the fused epilogue) must agree to
1e-5 x the sum of |terms| (fp32 sums in ATen's order on one side, fixed-order fp64 on the other)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
IMG = 96


def _bits(t):
    return t.detach().float().cpu().numpy().view(np.uint32)


def _build(per_channel_lsq: bool):
    from vsiquantization_b200.modules.fuse import fuse_modules_unified
    from vsiquantization_b200.modules.fuse_config import FuseConfig, create_fuse_config_manager
    from vsiquantization_b200.nets import yolov8
    from vsiquantization_b200.utils.quantize_manager import activate_learning_qparam, activate_quantizer, calibrate_qat_model
    torch.manual_seed(0)
    model = yolov8.yolo_v8_n(num_classes=20)
    g = torch.Generator().manual_seed(1)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    model = model.cuda()
    if per_channel_lsq:
        cfg = FuseConfig(observer_w_name="LSQObserver", quantizer_w_name="LSQQuantizer", observer_a_name="LSQObserver",
                         quantizer_a_name="LSQQuantizer", w_symmetric=False, a_symmetric=False, bits_w=4, bits_a=8,
                         w_ch_axis=0, a_ch_axis=1)
    else:
        cfg = FuseConfig(bits_w=8, bits_a=8)
    model = fuse_modules_unified(model, [["conv", "bn", "relu"]], config_manager=create_fuse_config_manager(cfg))
    gc = torch.Generator().manual_seed(2)
    calib = [(torch.randint(0, 256, (2, 3, IMG, IMG), generator=gc, dtype=torch.uint8), None) for _ in range(2)]

    def data_calib(m, loader, dev):
        m.eval()
        with torch.no_grad():
            for imgs, _ in loader:
                m(imgs.to(dev).float() / 255.0)

    calibrate_qat_model(model, calib, data_calib, torch.device("cuda"))
    activate_learning_qparam(model, use_init=True)
    activate_quantizer(model)
    return model.cuda().train()


def _fused(model):
    return [(n, m) for n, m in model.named_modules() if hasattr(m, "weight_quantizer")]


def _make_eager(model, decomposed_bias: bool):
    """Same model, same Parameters; only the quantizer plugin objects are swapped for the eager composition.  With
    ``decomposed_bias`` the layer computes conv(x, Wq) + b as two roundings, which is what the fused NHWC epilogue does
    (cuDNN's own bias add rounds once)."""
    from oracle.torch_port import EagerQuantizer
    from vsiquantization_b200.modules.fused import ConvBnReLU
    e = copy.deepcopy(model)
    for _, m in _fused(e):
        for mgr in (m.weight_quantizer, m.activation_quantizer):
            mgr.quantizer = EagerQuantizer.like(mgr.quantizer)
        if decomposed_bias and isinstance(m, ConvBnReLU):
            def fwd(x, m=m):
                w, b = m.get_weight_bias()
                pre = m._conv(x, m.quantize_weights(w), None) + b.view(1, -1, 1, 1)
                return m.quantize_activation(F.relu(pre))
            m.forward = fwd
    return e


def _step(model, x, mass=None):
    """One forward + backward; returns per-layer outputs, loss, input gradient, parameter gradients."""
    acts, hooks = {}, []
    for n, m in _fused(model):
        hooks.append(m.register_forward_hook(lambda mod, i, o, n=n: acts.__setitem__(n, o.detach())))
    if mass is not None:  # sum of |terms| per quantiser: the yardstick for the reduced gradients
        for n, m in _fused(model):
            for attr in ("weight_quantizer", "activation_quantizer"):
                mgr = getattr(m, attr)
                orig = mgr.quantizer.quantize

                def wrapped(xq, scale, zp, learn=False, _o=orig, _k=f"{n}.{attr}", _q=mgr.quantizer, **kw):
                    yq = _o(xq, scale, zp, learn, **kw)
                    if yq.requires_grad:
                        xa = xq.detach()
                        if kw.get("bias") is not None:
                            xa = xa + kw["bias"].detach().view(1, -1, 1, 1)
                        if kw.get("pre_relu"):
                            xa = F.relu(xa)

                        def grab(gq, xq=xa, yq=yq.detach()):
                            C = scale.numel()
                            shape = [1] * xq.dim()
                            if C > 1:
                                shape[_q.ch_axis] = C
                            s = scale.detach().float().reshape(shape if C > 1 else ()).double().abs()
                            gs = (_q.qmax * xq.numel() / C) ** -0.5
                            t = gq.double().abs() * (yq.double().abs() + xq.double().abs()) / s
                            dims = [d for d in range(xq.dim()) if C == 1 or d != _q.ch_axis]
                            mass[_k + ".scale"] = (gs * t.sum(dims)).reshape(-1)
                            mass[_k + ".zero_point"] = (gs * (gq.double().abs() * s).sum(dims)).reshape(-1)
                        yq.register_hook(grab)
                    return yq
                mgr.quantizer.quantize = wrapped
    x = x.clone().requires_grad_(True)
    model.zero_grad(set_to_none=True)
    outs = model(x)
    loss = sum((o.float() ** 2).mean() for o in outs)
    loss.backward()
    for h in hooks:
        h.remove()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return acts, float(loss.detach()), x.grad.detach(), grads


def _compare(a, b, mass, what, bias_is_sum=False):
    acts_a, loss_a, dx_a, g_a = a
    acts_b, loss_b, dx_b, g_b = b
    assert set(acts_a) == set(acts_b) and len(acts_a) == 57
    bad = [n for n in acts_a if not np.array_equal(_bits(acts_a[n]), _bits(acts_b[n]))]
    assert bad == [], f"{what}: fused-layer outputs differ in {bad[:4]}"
    assert loss_a == loss_b, (what, loss_a, loss_b)
    assert np.array_equal(_bits(dx_a), _bits(dx_b)), f"{what}: input gradient differs"
    assert set(g_a) == set(g_b)
    n_sum = 0
    for n in g_a:
        ga, gb = g_a[n], g_b[n]
        if n.endswith(("quantizer.scale", "quantizer.zero_point")):
            m = mass[n].to(ga.device)
            err = (ga.double() - gb.double()).abs().reshape(-1)
            assert bool((err <= 1e-5 * m + 1e-30).all()), (what, n, float((err / (m + 1e-30)).max()))
            n_sum += 1
        elif bias_is_sum and n.endswith(".bias"):
            assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-6 * float(gb.abs().max() + 1e-30)), (what, n)
        else:
            assert np.array_equal(_bits(ga), _bits(gb)), f"{what}: gradient of {n} differs"
    return n_sum


@pytest.mark.parametrize("per_channel_lsq", [False, True], ids=["w8a8_per_tensor", "w4a8_per_channel_asym_lsq"])
def test_native_equals_eager_composition_at_model_scale(per_channel_lsq, monkeypatch):
    from vsiquantization_b200.bank import WeightBank
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)
    monkeypatch.setattr(torch.backends.cudnn, "benchmark", False)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    model = _build(per_channel_lsq)
    gx = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (2, 3, IMG, IMG), generator=gx, dtype=torch.uint8).cuda().float() / 255.0
    n_q = 4 * 57 if per_channel_lsq else 2 * 57  # learnable scales (+ zero-points when asymmetric)

    # ---- NCHW: per-layer native launches (ReLU fused into the quantiser) vs the eager composition
    mass = {}
    native = _step(copy.deepcopy(model), x, mass)
    eager = _step(_make_eager(model, decomposed_bias=False), x)
    assert _compare(native, eager, mass, "NCHW native vs eager") == n_q

    # ---- channels_last: fused bias + ReLU + quantiser epilogue (TMA-staged backward) vs the eager composition
    xcl = x.contiguous(memory_format=torch.channels_last)
    mcl = copy.deepcopy(model).to(memory_format=torch.channels_last)
    mass_cl = {}
    native_cl = _step(copy.deepcopy(mcl), xcl, mass_cl)
    eager_cl = _step(_make_eager(mcl, decomposed_bias=True), xcl)
    assert _compare(native_cl, eager_cl, mass_cl, "channels_last native vs eager", bias_is_sum=True) == n_q

    # ---- weight bank (one multi-tensor launch each way) vs per-layer launches: same values
    mb = copy.deepcopy(mcl)
    bank = WeightBank(mb).install()
    banked = _step(mb, xcl)
    assert bank.last_used
    assert _compare(banked, native_cl, mass_cl, "weight bank vs per-layer") == n_q
    bank.remove()

    # ---- three SGD steps, captured as one CUDA graph, vs the same three steps launched eagerly: identical trajectories
    from vsiquantization_b200.graph import GraphedQATStep
    loss_fn = lambda outs: sum((o.float() ** 2).mean() for o in outs)  # noqa: E731
    batches = [torch.randint(0, 256, (2, 3, IMG, IMG), generator=gx, dtype=torch.uint8).cuda().float().div(255.0)
               .contiguous(memory_format=torch.channels_last) for _ in range(3)]
    m1, m2 = copy.deepcopy(mcl), copy.deepcopy(mcl)
    o1 = torch.optim.SGD(m1.parameters(), lr=1e-3, momentum=0.9, nesterov=True)
    o2 = torch.optim.SGD(m2.parameters(), lr=1e-3, momentum=0.9, nesterov=True)
    WeightBank(m1).install()
    WeightBank(m2).install()
    for _ in range(3):  # the warm-up steps GraphedQATStep runs on its example input
        o1.zero_grad(set_to_none=True)
        loss_fn(m1(batches[0])).backward()
        o1.step()
    step = GraphedQATStep(m2, o2, loss_fn, batches[0], warmup=3)
    for b in batches:
        o1.zero_grad(set_to_none=True)
        l1 = loss_fn(m1(b))
        l1.backward()
        o1.step()
        l2 = step(b)
        assert float(l1.detach()) == float(l2), "graphed and eagerly launched steps diverge"
    for (n1, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert np.array_equal(_bits(p1), _bits(p2)), n1


@pytest.mark.gpu
@pytest.mark.parametrize("phase", ["train", "calibrate"])
def test_channel_slice_input_keeps_the_channels_last_epilogue(phase, monkeypatch):
    """C2f blocks feed a fused layer ``chunk(2, dim=1)[1]`` of a channels_last tensor: stride(1) == 1 but pitched rows.
    The layer densifies that view once and stays on the NHWC epilogue (bias add, ReLU, quantiser -- or observer while
    calibrating -- in one pass; dx + dbias + dscale in one backward pass) instead of falling back to a convolution with
    its own bias-add / bias-gradient passes and an input copy in each direction.  Outputs and every gradient equal those
    of the same layer fed a dense copy of the slice, bit for bit, with the same number of native launches."""
    from vsiquantization_b200 import _lib
    from vsiquantization_b200.modules.fused import ConvBnReLU
    from vsiquantization_b200.utils.quantize_manager import activate_learning_qparam, activate_quantizer, calibrate_qat_model
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(3)
    cv, bn = torch.nn.Conv2d(16, 32, 3, padding=1, bias=False), torch.nn.BatchNorm2d(32)
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    layer = ConvBnReLU(cv, bn, torch.nn.ReLU(), "LSQObserver", "LSQQuantizer", "LSQObserver", "LSQQuantizer", True, False,
                       True, 4, 8).cuda().to(memory_format=torch.channels_last)
    wide = torch.randn(4, 32, 20, 20, device="cuda").contiguous(memory_format=torch.channels_last)
    calib = lambda m, loader, dev: [m(b) for b in loader]  # noqa: E731
    if phase == "calibrate":
        import copy
        twin = copy.deepcopy(layer)
        n0 = _lib.launch_count
        calibrate_qat_model(layer, [wide.chunk(2, 1)[1]], calib)
        n_slice = _lib.launch_count - n0
        n0 = _lib.launch_count
        calibrate_qat_model(twin, [wide.chunk(2, 1)[1].contiguous(memory_format=torch.channels_last)], calib)
        assert _lib.launch_count - n0 == n_slice
        a, b = layer.activation_quantizer, twin.activation_quantizer
        assert a.observer.get_scale_zero_point() == b.observer.get_scale_zero_point()
        return
    calibrate_qat_model(layer, [wide.chunk(2, 1)[1].contiguous(memory_format=torch.channels_last)], calib)
    activate_learning_qparam(layer, use_init=True)
    activate_quantizer(layer)
    layer.train()

    def run(dense):
        layer.zero_grad(set_to_none=True)
        w = wide.clone().requires_grad_(True)
        x = w.chunk(2, 1)[1]
        assert x.stride(1) == 1 and not x.is_contiguous(memory_format=torch.channels_last)
        if dense:
            x = x.contiguous(memory_format=torch.channels_last)
        n0 = _lib.launch_count
        y = layer(x)
        (y * torch.linspace(-1, 1, y.numel(), device="cuda").view_as(y)).sum().backward()
        grads = {n: p.grad.detach().clone() for n, p in layer.named_parameters() if p.grad is not None}
        return y.detach().clone(), w.grad.detach().clone(), grads, _lib.launch_count - n0

    y0, dw0, g0, n_dense = run(True)
    y1, dw1, g1, n_slice = run(False)
    assert n_slice == n_dense
    assert y1.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y1, y0) and torch.equal(dw1, dw0)
    assert bool((dw1[:, :16] == 0).all())  # the other half of the chunk gets no gradient
    assert set(g1) == set(g0) and any(n.endswith("conv_fuse.bias") for n in g0)
    for n in g0:
        assert torch.equal(g1[n], g0[n]), n
