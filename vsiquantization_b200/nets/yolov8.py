"""YOLOv8 n/t/s/m/l/x in plain PyTorch -- the fixture the QAT path is exercised on.

Architecture and parameter names follow the reference's nets/yolov8.py (so its float checkpoints load and its fuse
pattern ["conv","bn","relu"] applies: every block is conv / norm / relu children, ReLU not SiLU, nets/yolov8.py:13-30;
the last 1x1 convolutions of the head are bare Conv2d and stay float, :170-175).  Written from the published YOLOv8
topology (CSP-Darknet backbone -> PAN-FPN neck -> decoupled head), not copied: builders are table-driven.

Fused-layer counts (conv+bn+relu blocks): n/t/s 57, m 77, l/x 97.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

# depth (repeats of the residual unit in the three CSP stage kinds) and channel widths per variant
VARIANTS = {
    "n": ([1, 2, 2], [3, 16, 32, 64, 128, 256]),
    "t": ([1, 2, 2], [3, 24, 48, 96, 192, 384]),
    "s": ([1, 2, 2], [3, 32, 64, 128, 256, 512]),
    "m": ([2, 4, 4], [3, 48, 96, 192, 384, 576]),
    "l": ([3, 6, 6], [3, 64, 128, 256, 512, 512]),
    "x": ([3, 6, 6], [3, 80, 160, 320, 640, 640]),
}


class Conv(nn.Module):
    """conv (no bias) -> BatchNorm(eps 1e-3, momentum 0.03) -> ReLU; children named conv / norm / relu."""

    def __init__(self, in_ch: int, out_ch: int, k: int = 1, s: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, k, s, (k - 1) // 2, bias=False)
        self.norm = nn.BatchNorm2d(out_ch, eps=0.001, momentum=0.03)
        self.relu = nn.ReLU()

    def forward(self, x):
        return self.relu(self.norm(self.conv(x)))


class Residual(nn.Module):
    def __init__(self, ch: int, add: bool = True):
        super().__init__()
        self.add_m = add
        self.conv1 = Conv(ch, ch, 3)
        self.conv2 = Conv(ch, ch, 3)

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return x + y if self.add_m else y


class CSP(nn.Module):
    """C2f: split, n residual units chained on the running half, concatenate everything, 1x1 merge."""

    def __init__(self, in_ch: int, out_ch: int, n: int = 1, add: bool = True):
        super().__init__()
        self.conv1 = Conv(in_ch, out_ch)
        self.conv2 = Conv((2 + n) * out_ch // 2, out_ch)
        self.res_m = nn.ModuleList(Residual(out_ch // 2, add) for _ in range(n))

    def forward(self, x):
        parts: List[torch.Tensor] = list(self.conv1(x).chunk(2, dim=1))
        for unit in self.res_m:
            parts.append(unit(parts[-1]))
        return self.conv2(torch.cat(parts, dim=1))


class SPP(nn.Module):
    """SPPF: three chained 5x5 max-pools concatenated with the input."""

    def __init__(self, in_ch: int, out_ch: int, k: int = 5):
        super().__init__()
        self.conv1 = Conv(in_ch, in_ch // 2)
        self.conv2 = Conv(in_ch * 2, out_ch)
        self.res_m = nn.MaxPool2d(k, 1, k // 2)

    def forward(self, x):
        x = self.conv1(x)
        pools = [x]
        for _ in range(3):
            pools.append(self.res_m(pools[-1]))
        return self.conv2(torch.cat(pools, 1))


class DarkNet(nn.Module):
    def __init__(self, width: Sequence[int], depth: Sequence[int]):
        super().__init__()
        w, d = width, depth
        self.p1 = nn.Sequential(Conv(w[0], w[1], 3, 2))
        self.p2 = nn.Sequential(Conv(w[1], w[2], 3, 2), CSP(w[2], w[2], d[0]))
        self.p3 = nn.Sequential(Conv(w[2], w[3], 3, 2), CSP(w[3], w[3], d[1]))
        self.p4 = nn.Sequential(Conv(w[3], w[4], 3, 2), CSP(w[4], w[4], d[2]))
        self.p5 = nn.Sequential(Conv(w[4], w[5], 3, 2), CSP(w[5], w[5], d[0]), SPP(w[5], w[5]))

    def forward(self, x):
        p3 = self.p3(self.p2(self.p1(x)))
        p4 = self.p4(p3)
        return p3, p4, self.p5(p4)


class DarkFPN(nn.Module):
    def __init__(self, width: Sequence[int], depth: Sequence[int]):
        super().__init__()
        w, n = width, depth[0]
        self.up = nn.Upsample(size=None, scale_factor=2)
        self.h1 = CSP(w[4] + w[5], w[4], n, False)
        self.h2 = CSP(w[3] + w[4], w[3], n, False)
        self.h3 = Conv(w[3], w[3], 3, 2)
        self.h4 = CSP(w[3] + w[4], w[4], n, False)
        self.h5 = Conv(w[4], w[4], 3, 2)
        self.h6 = CSP(w[4] + w[5], w[5], n, False)

    def forward(self, p3, p4, p5):
        p4 = self.h1(torch.cat([self.up(p5), p4], 1))
        p3 = self.h2(torch.cat([self.up(p4), p3], 1))
        p4 = self.h4(torch.cat([self.h3(p3), p4], 1))
        p5 = self.h6(torch.cat([self.h5(p4), p5], 1))
        return p3, p4, p5


class Head(nn.Module):
    """Decoupled head: per scale a box branch (4 outputs) and a class branch (nc outputs), concatenated."""

    def __init__(self, nc: int = 80, ch: Sequence[int] = ()):
        super().__init__()
        self.nc = nc
        self.no = nc + 4
        self.stride = torch.zeros(len(ch))
        box = max(64, ch[0] // 4)
        cls = max(80, ch[0], nc)
        self.box = nn.ModuleList(nn.Sequential(Conv(c, box, 3), Conv(box, box, 3), nn.Conv2d(box, 4, 1)) for c in ch)
        self.cls = nn.ModuleList(nn.Sequential(Conv(c, cls, 3), Conv(cls, cls, 3), nn.Conv2d(cls, nc, 1)) for c in ch)

    def forward(self, p3, p4, p5):
        return [torch.cat((b(f), c(f)), 1) for f, b, c in zip((p3, p4, p5), self.box, self.cls)]


class YOLO(nn.Module):
    def __init__(self, width: Sequence[int], depth: Sequence[int], num_classes: int):
        super().__init__()
        self.net = DarkNet(width, depth)
        self.fpn = DarkFPN(width, depth)
        self.head = Head(num_classes, (width[3], width[4], width[5]))
        self.head.stride = torch.tensor([8.0, 16.0, 32.0])  # p3 / p4 / p5 (the reference derives them from a dummy pass)
        self.stride = self.head.stride

    def forward(self, x):
        return self.head(*self.fpn(*self.net(x)))


def _variant(name: str, num_classes: int) -> YOLO:
    depth, width = VARIANTS[name]
    return YOLO(width, depth, num_classes)


def yolo_v8_n(num_classes: int = 80): return _variant("n", num_classes)  # noqa: E704
def yolo_v8_t(num_classes: int = 80): return _variant("t", num_classes)  # noqa: E704
def yolo_v8_s(num_classes: int = 80): return _variant("s", num_classes)  # noqa: E704
def yolo_v8_m(num_classes: int = 80): return _variant("m", num_classes)  # noqa: E704
def yolo_v8_l(num_classes: int = 80): return _variant("l", num_classes)  # noqa: E704
def yolo_v8_x(num_classes: int = 80): return _variant("x", num_classes)  # noqa: E704
