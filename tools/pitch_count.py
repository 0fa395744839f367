"""How many channel-innermost backward launches of a YOLOv8 QAT step get a dense / pitched / copied grad_output?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks import yolo_qat
from vsiquantization_b200 import ops
args = yolo_qat.parse(["--model", "s", "--batch", "16", "--imgsz", "640", "--steps", "1", "--channels-last"] + sys.argv[1:])
dev = torch.device("cuda")
model, _, _ = yolo_qat.build_model(args, dev)
model.train().to(memory_format=torch.channels_last)
stats = {}
orig = ops.ci_backward
def counting(x, bias, g, *a, **k):
    kind = "dense" if g.stride() == x.stride() else ("pitched" if (g.stride(1) == 1 and g.stride(3) % 4 == 0 and g.stride(3) >= x.shape[1]) else "other")
    key = (kind, tuple(x.shape))
    stats[key] = stats.get(key, 0) + 1
    return orig(x, bias, g, *a, **k)
ops.ci_backward = counting
x = torch.rand(16, 3, 640, 640, device=dev).contiguous(memory_format=torch.channels_last)
loss = sum((o.float() ** 2).mean() for o in model(x)); loss.backward()
torch.cuda.synchronize()
tot = {}
for (kind, shp), n in sorted(stats.items()):
    el = n * shp[0] * shp[1] * shp[2] * shp[3]
    tot[kind] = tot.get(kind, 0) + el
    print(kind, shp, n)
print({k: v / sum(tot.values()) for k, v in tot.items()})
