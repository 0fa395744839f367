"""Where do the copy / add / cat kernels of a YOLOv8s QAT step come from?

torch profiler with shapes, Python stacks and the autograd node each op runs under; copies are attributed to the Python
line that issued them (forward) or to the backward node (`autograd::engine::evaluate_function: ...`) they run inside.
Only leaf ops are counted (aten::copy_ under aten::contiguous / clone is one copy), device time per step.

    python tools/copy_sources.py --channels-last --weight-bank --w-bits 4 --a-bits 8 --asym --per-channel --lsq
    python tools/copy_sources.py --toy        # CPU self-test of the attribution logic (no GPU, no package)
"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TARGETS = ("aten::copy_", "aten::add", "aten::add_", "aten::cat", "aten::_cat", "aten::upsample_nearest2d",
           "aten::upsample_nearest2d_backward", "aten::max_pool2d_with_indices", "aten::max_pool2d_with_indices_backward",
           "aten::mul", "aten::mul_", "aten::fill_", "aten::zero_", "aten::sum", "aten::mean", "aten::div", "aten::pow")


def dev_time(e):
    for a in ("self_device_time_total", "self_cuda_time_total"):
        v = getattr(e, a, None)
        if v:
            return float(v)
    return 0.0


def context(e):
    """The backward node this op runs under, else the innermost Python frame outside torch."""
    p = e
    chain = []
    while p is not None:
        chain.append(p.name)
        if p.name.startswith("autograd::engine::evaluate_function"):
            return p.name.replace("autograd::engine::evaluate_function: ", "bwd ")
        p = getattr(p, "cpu_parent", None)
    for fr in (e.stack or []):
        if "site-packages/torch" not in fr and "<built-in" not in fr and "copy_sources.py" not in fr:
            return fr.strip()[-110:]
    outer = [n for n in chain[1:] if n.startswith("aten::") or n.startswith("nn.Module")]
    return "under " + (outer[0] if outer else "?")


def report(prof, steps, use_cpu_time=False):
    agg = collections.defaultdict(lambda: [0.0, 0])
    per_op = collections.defaultdict(lambda: [0.0, 0])
    for e in prof.events():
        if e.name not in TARGETS:
            continue
        t = float(e.self_cpu_time_total) if use_cpu_time else dev_time(e)
        if t <= 0:
            continue
        shapes = str(getattr(e, "input_shapes", ""))[:70]
        key = (e.name, shapes, context(e))
        agg[key][0] += t
        agg[key][1] += 1
        per_op[e.name][0] += t
        per_op[e.name][1] += 1
    print("per op (ms per step, launches per step):")
    for k, (t, n) in sorted(per_op.items(), key=lambda kv: -kv[1][0]):
        print(f"  {t / 1e3 / steps:8.3f} ms x{n / steps:6.1f}  {k}")
    print("by source:")
    for (name, shapes, ctx), (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        print(f"  {t / 1e3 / steps:8.3f} ms x{n / steps:5.1f}  {name:28s} {shapes:70s} {ctx}")


def toy():
    lin = torch.nn.Conv2d(4, 8, 1)
    x = torch.randn(2, 4, 8, 8, requires_grad=True)

    def step():
        y = lin(x)
        a, b = y.chunk(2, 1)
        z = torch.cat([a, b.contiguous() + 1, a], 1)
        z.transpose(1, 2).contiguous().sum().backward()
    step()
    with profile(activities=[ProfilerActivity.CPU], record_shapes=True, with_stack=True,
                 experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
        step()
    report(prof, 1, use_cpu_time=True)


def main():
    if "--toy" in sys.argv:
        return toy()
    from benchmarks import yolo_qat
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
        outs = model(x)
        loss = sum((o.float() ** 2).mean() for o in outs)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True,
                 experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
        step()
        torch.cuda.synchronize()
    report(prof, 1)


if __name__ == "__main__":
    main()
