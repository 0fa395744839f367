"""Tiny float CNN shared by oracle/gen_golden.py (fused with the REFERENCE's modules)
and tests/ (fused with this repo's modules).  Child names follow the reference's
nets/yolov8.py block convention (conv / norm / relu) so the same fuse pattern applies."""
import torch


class Block(torch.nn.Module):
    def __init__(self, cin, cout, k=3, s=1, act="relu", bias=False):
        super().__init__()
        self.conv = torch.nn.Conv2d(cin, cout, k, s, (k - 1) // 2, bias=bias)
        self.norm = torch.nn.BatchNorm2d(cout, eps=0.001, momentum=0.03)
        self.relu = torch.nn.ReLU() if act == "relu" else torch.nn.SiLU()

    def forward(self, x):
        return self.relu(self.norm(self.conv(x)))


class TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = Block(3, 8, 3, 2)
        self.backbone = torch.nn.Sequential(Block(8, 16, 3, 2, bias=True), Block(16, 16, 1, 1, act="silu"))
        self.head = torch.nn.Conv2d(16, 5, 1)

    def forward(self, x):
        return self.head(self.backbone(self.stem(x)))


def make_tiny(seed=0):
    g = torch.Generator().manual_seed(seed)
    m = TinyNet()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.2)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
    return m
