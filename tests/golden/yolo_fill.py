"""Deterministic weights for the YOLOv8 fixture, shared by oracle/gen_golden.py (which fills the REFERENCE's
nets/yolov8.py model) and tests/ (which fills vsiquantization_b200/nets/yolov8.py -- the same network: same parameter and
buffer names, shapes and order, tests/test_abi_and_host.py::test_yolov8_fixture_is_the_reference_network).
Variance-preserving so that activations neither vanish nor overflow through ~60 layers."""
import torch


def fill_(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / fan_in) ** 0.5)
            elif name.endswith("norm.weight") or ".norm." in name and name.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
    return model
