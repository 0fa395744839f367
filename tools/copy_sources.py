"""Where do the aten::copy_ / contiguous kernels of a YOLOv8s QAT step come from?  (torch profiler with stacks + shapes)"""
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
opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9, nesterov=True)
x = torch.rand(64, 3, 640, 640, device=dev)
if args.channels_last:
    x = x.contiguous(memory_format=torch.channels_last)
def step():
    outs = model(x); loss = sum((o.float() ** 2).mean() for o in outs); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=12) if e.key in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::add_", "aten::add")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:24]:
    print(f"{e.device_time_total/1e3:8.3f} ms x{e.count:3d} {e.key} {str(e.input_shapes)[:90]}")
    for fr in e.stack[:12]:
        if "site-packages/torch" not in fr:
            print("        ", fr[:150])
