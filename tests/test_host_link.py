"""Host logic of the quantize_inp link (no GPU): discovery by tensor identity, one consumer per producer, unlinking, and
the consumer-side pick-up keyed on the exact tensor object."""
import weakref

import torch

from vsiquantization_b200.quantizers.fake_quantize import FakeQuantize
from vsiquantization_b200.utils.quantize_manager import link_quantize_inp


class _Stub(FakeQuantize):
    """A fused layer without kernels: forward adds one, so every output is a fresh tensor object."""

    def __init__(self, quantize_inp=False):
        torch.nn.Module.__init__(self)
        self.activation_quantizer = torch.nn.Identity()
        self.quantize_inp = quantize_inp
        self.seen_pre = None

    def forward(self, x):
        if self.quantize_inp:
            x = self.quantize_input(x)
        return x + 1

    def quantize_activation(self, x):
        return x * 1.0


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b, self.c, self.d = _Stub(), _Stub(True), _Stub(True), _Stub(True)

    def forward(self, x):
        y = self.a(x)
        y1 = self.b(y)                 # a -> b: direct
        y2 = self.c(y)                 # a -> c: second consumer of the same producer, not linked
        return self.d(torch.cat([y1, y2], 1))   # re-packed on the way: no link


def test_discovery_links_direct_pairs_only():
    net = _Net()
    assert link_quantize_inp(net, torch.zeros(1, 2, 2, 2)) == 1
    assert net.a.__dict__["_inp_consumer"] is net.b
    for m in (net.b, net.c, net.d):
        assert m.__dict__.get("_inp_consumer") is None
    net.b.quantize_inp = False
    assert link_quantize_inp(net, torch.zeros(1, 2, 2, 2)) == 1      # now c is the first quantize_inp consumer of a
    assert net.a.__dict__["_inp_consumer"] is net.c
    net.c.quantize_inp = False
    assert link_quantize_inp(net, torch.zeros(1, 2, 2, 2)) == 0 and net.a.__dict__["_inp_consumer"] is None


def test_consumer_picks_up_only_the_tensor_it_was_offered():
    m = _Stub(True)
    x, other, pre = torch.ones(2), torch.ones(2), torch.full((2,), 7.0)
    m.__dict__["_prequant"] = (weakref.ref(x), pre)
    assert m.quantize_input(x) is pre and m.__dict__["_prequant"] is None      # consumed once
    assert torch.equal(m.quantize_input(x), x)                                  # then the ordinary step again
    m.__dict__["_prequant"] = (weakref.ref(x), pre)
    assert torch.equal(m.quantize_input(other), other) and m.__dict__["_prequant"] is None   # stale offers are dropped


def test_link_rejects_non_layers():
    import pytest
    with pytest.raises(TypeError):
        _Stub().feed_input_quantizer_of(torch.nn.ReLU())


def test_observers_with_their_own_update_rule_keep_their_separate_pass():
    """The calibration epilogue folds MinMaxObserver.observe into the activation pass; a subclass that overrides observe()
    (the moving-average observers) must not inherit that shortcut."""
    from vsiquantization_b200.observers.minmax import LSQObserver, MinMaxObserver
    from vsiquantization_b200.observers.moving_average import MovingAverageMinMaxObserver
    x = torch.zeros(1, 4, 2, 2)
    assert MovingAverageMinMaxObserver().observe_epilogue(x) is None
    assert type(LSQObserver(True)).observe is MinMaxObserver.observe
    assert MinMaxObserver(True).observe_epilogue(x) is None      # a CPU / NCHW tensor: not applicable either, no launch
