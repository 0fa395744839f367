"""Kernel-time breakdown of one YOLOv8s QAT step (torch profiler, CUDA activity)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks import yolo_qat
from torch.profiler import profile, ProfilerActivity
args = yolo_qat.parse(["--model", "s", "--batch", "64", "--imgsz", "640", "--steps", "2"] + sys.argv[1:])
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda")
model, n_fused, _ = yolo_qat.build_model(args, dev)
model.train()
if args.channels_last:
    model.to(memory_format=torch.channels_last)
if args.weight_bank:
    from vsiquantization_b200.bank import WeightBank
    WeightBank(model).install()
opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9, nesterov=True)
x = torch.rand(64, 3, 640, 640, device=dev)
if args.channels_last:
    x = x.contiguous(memory_format=torch.channels_last)
def step():
    outs = model(x); loss = sum((o.float() ** 2).mean() for o in outs); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print("total device ms per step", tot / 3e3)
for e in rows[:28]:
    print(f"{e.device_time_total/3e3:8.3f} ms {100*e.device_time_total/tot:5.1f}%  x{e.count//3:4d}  {e.key[:110]}")
