"""``quantize_inp`` fused into the producer's epilogue (reference: quantizers/fake_quantize.py:44-45,
``if self.quantize_inp: x = self.quantize_activation(x)``): one pass over the conv output writes this layer's quantised
output AND the next layer's input quantisation.  The bar is the unlinked model on the same GPU, bit for bit: the second
stage quantises the very value that is stored, and the backward is the same pair of kernels either way."""
import numpy as np
import pytest
import torch

import oracle
from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("act", ["none", "relu", "silu"])
@pytest.mark.parametrize("pc1,pc2", [(False, False), (True, False), (False, True), (True, True)])
def test_two_output_epilogue_equals_two_launches(act, pc1, pc2):
    from vsiquantization_b200 import _lib, ops
    torch.manual_seed(11)
    N, C, H, W = 3, 24, 9, 7
    x = _cl(torch.randn(N, C, H, W, device="cuda") * 2.0)
    flat = x.permute(0, 2, 3, 1).reshape(-1)
    # every special the single-launch kernels are tested on: the per-vector guard must route them to the IEEE path
    flat[:8] = torch.tensor([float("inf"), -float("inf"), float("nan"), 0.0, -0.0, 1e-42, -3e38, 5e-20], device="cuda")
    bias = torch.randn(C, device="cuda") * 0.1
    s1 = (torch.rand(C, device="cuda") * 0.05 + 0.01) if pc1 else 0.037
    z1 = torch.randint(0, 16, (C,), device="cuda").float() if pc1 else 3
    s2 = (torch.rand(C, device="cuda") * 0.2 + 0.05) if pc2 else torch.tensor(0.11, device="cuda")
    z2 = torch.randint(-3, 4, (C,), device="cuda").float() if pc2 else 0
    spec1 = ops.QSpec(0, 255, ch_axis=1 if pc1 else None, pre_relu=act == "relu", pre_silu=act == "silu")
    spec2 = ops.QSpec(-8, 7, ch_axis=1 if pc2 else None)
    for b in (bias, None):
        n0 = _lib.launch_count
        y, y2 = ops.ci_forward(x, b, s1, z1, spec1, second=(s2, z2, spec2))
        assert _lib.launch_count - n0 == 1
        y_ref = ops.ci_forward(x, b, s1, z1, spec1)
        y2_ref = ops.fake_quant_forward(y_ref, s2, z2, spec2)
        assert y.stride() == x.stride() and y2.stride() == x.stride()
        assert bits_equal(y.cpu().numpy(), y_ref.cpu().numpy())
        assert bits_equal(y2.cpu().numpy(), y2_ref.cpu().numpy())


def test_two_output_epilogue_second_stage_against_the_oracle():
    """The second output against the CPU oracle's forward over the first output (W8 activations -> W4 input codes)."""
    from vsiquantization_b200 import ops
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 16, 12, 12), dtype=np.float32)
    xt = _cl(torch.from_numpy(x).cuda())
    y, y2 = ops.ci_forward(xt, None, 3.0 / 127, 0, ops.QSpec(-128, 127, pre_relu=True),
                           second=(3.0 / 7, 8, ops.QSpec(0, 15)))
    y_np = y.cpu().numpy()
    want = np.asarray(oracle.fake_quant_fwd(np.ascontiguousarray(y_np).reshape(-1), 3.0 / 7, 8, 0, 15)).reshape(y_np.shape)
    assert bits_equal(y2.cpu().numpy(), want)


def test_odd_scales_take_the_ieee_path_in_both_stages():
    from vsiquantization_b200 import ops
    torch.manual_seed(2)
    x = _cl(torch.randn(2, 8, 6, 6, device="cuda"))
    for s1, s2 in ((1e-30, 0.1), (0.05, 3e25), (float(np.float32(2.0) ** -45), float(np.float32(2.0) ** 50))):
        y, y2 = ops.ci_forward(x, None, s1, 0, ops.QSpec(-128, 127), second=(s2, 0, ops.QSpec(-128, 127)))
        y_ref = ops.ci_forward(x, None, s1, 0, ops.QSpec(-128, 127))
        assert bits_equal(y.cpu().numpy(), y_ref.cpu().numpy())
        assert bits_equal(y2.cpu().numpy(), ops.fake_quant_forward(y_ref, s2, 0, ops.QSpec(-128, 127)).cpu().numpy())


def _two_layers(per_channel: bool, act_cls=torch.nn.ReLU):
    from vsiquantization_b200.modules.fused import ConvBnReLU
    torch.manual_seed(7)
    layers = []
    for cin, cout in ((8, 16), (16, 16)):
        cv, bn = torch.nn.Conv2d(cin, cout, 3, padding=1, bias=False), torch.nn.BatchNorm2d(cout)
        with torch.no_grad():
            bn.running_mean.normal_(0, 0.1)
            bn.running_var.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.2)
        kw = {"a_ch_axis": 1} if per_channel else {}
        layers.append(ConvBnReLU(cv, bn, act_cls(), "LSQObserver", "LSQQuantizer", "LSQObserver", "LSQQuantizer", True,
                                 False, True, 4, 8, **kw))
    layers[1].quantize_inp = True
    return torch.nn.Sequential(*layers).cuda().to(memory_format=torch.channels_last)


@pytest.mark.parametrize("learn", [False, True], ids=["fixed", "lsq"])
@pytest.mark.parametrize("per_channel", [False, True], ids=["per-tensor", "per-channel"])
def test_linked_layers_match_the_unlinked_model_bit_for_bit(learn, per_channel, monkeypatch):
    from vsiquantization_b200 import _lib
    from vsiquantization_b200.utils.quantize_manager import (activate_learning_qparam, activate_quantizer,
                                                             calibrate_qat_model, link_quantize_inp)
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    model = _two_layers(per_channel)
    x = _cl(torch.randn(4, 8, 20, 20, device="cuda"))
    calibrate_qat_model(model, [x], lambda m, loader, dev: [m(b) for b in loader])
    if learn:
        activate_learning_qparam(model, use_init=True)
    else:
        for m in model:
            for q in (m.weight_quantizer, m.activation_quantizer):
                q.is_learning_scale = False
                q.is_observer_qparam = False
    activate_quantizer(model)
    model.train()

    def run():
        model.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        n0 = _lib.launch_count
        y = model(xi)
        fwd = _lib.launch_count - n0
        (y * torch.linspace(-1, 1, y.numel(), device="cuda").view_as(y)).sum().backward()
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        return y.detach().clone(), xi.grad.detach().clone(), grads, fwd, _lib.launch_count - n0 - fwd

    y0, dx0, g0, f0, b0 = run()
    assert link_quantize_inp(model, x) == 1
    assert model[0].__dict__["_inp_consumer"] is model[1]
    y1, dx1, g1, f1, b1 = run()
    assert f1 == f0 - 1 and b1 == b0, (f0, f1, b0, b1)   # one forward launch less, the same backward kernels
    assert model[1].__dict__.get("_prequant") is None     # consumed
    assert torch.equal(y1, y0) and torch.equal(dx1, dx0)
    assert set(g1) == set(g0) and len(g0) >= 4
    for n in g0:
        assert torch.equal(g1[n], g0[n]), n
    # a deep copy keeps the link inside the copy
    import copy
    twin = copy.deepcopy(model)
    assert twin[0].__dict__["_inp_consumer"] is twin[1]
    # consumer not quantising (or still observing): nothing is offered, the producer runs its ordinary epilogue
    model[1].activation_quantizer.is_quantize = False
    n0 = _lib.launch_count
    with torch.no_grad():
        model(x)
    assert model[1].__dict__.get("_prequant") is None and _lib.launch_count - n0 == f0 - 2
    model[1].activation_quantizer.is_quantize = True
    # unlinking restores the separate launch
    model[0].feed_input_quantizer_of(None)
    y2, _, _, f2, _ = run()
    assert f2 == f0 and torch.equal(y2, y0)


def test_link_survives_cuda_graph_capture(monkeypatch):
    """The linked pair inside a captured forward + backward: replays reproduce the eager result."""
    from vsiquantization_b200.utils.quantize_manager import (activate_learning_qparam, activate_quantizer,
                                                             calibrate_qat_model, link_quantize_inp)
    monkeypatch.setattr(torch.backends.cudnn, "deterministic", True)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    model = _two_layers(False, torch.nn.SiLU)
    x = _cl(torch.randn(2, 8, 16, 16, device="cuda"))
    calibrate_qat_model(model, [x], lambda m, loader, dev: [m(b) for b in loader])
    activate_learning_qparam(model, use_init=True)
    activate_quantizer(model)
    model.train()
    assert link_quantize_inp(model, x) == 1
    w = torch.linspace(-1, 1, 2 * 16 * 16 * 16, device="cuda").view(2, 16, 16, 16)

    def step(inp):
        model.zero_grad(set_to_none=False)
        y = model(inp)
        (y * w).sum().backward()
        return y

    y_eager = step(x).detach().clone()
    g_eager = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    static_x = x.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step(static_x)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_static = step(static_x)
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y_static, y_eager)
    for n, p in model.named_parameters():
        if n in g_eager:
            assert torch.equal(p.grad, g_eager[n]), n
