"""Why is fwd slower inside the fwd+bwd sequence?  Per-kernel events over alternating launches, preallocated outputs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import ops
n = 1 << 28
x = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda")
y = torch.empty_like(x); dx = torch.empty_like(x)
spec = ops.QSpec(-128, 127)
def run(seq, reps=10, label=""):
    for _ in range(3):
        for f in seq: f()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps * len(seq) + 1)]
    evs[0].record()
    k = 1
    for _ in range(reps):
        for f in seq:
            f(); evs[k].record(); k += 1
    torch.cuda.synchronize()
    per = [0.0] * len(seq)
    for r in range(reps):
        for j in range(len(seq)):
            i = r * len(seq) + j
            per[j] += evs[i].elapsed_time(evs[i + 1]) / reps
    print(label, " ".join(f"{p:.3f}" for p in per), "total/step %.3f" % sum(per))
fwd = lambda: ops.fake_quant_forward(x, 3.0 / 127, 0, spec, out=y)
bwd = lambda: ops.fake_quant_backward_ste(x, g, 3.0 / 127, 0, spec, out=dx)
fwd_alloc = lambda: ops.fake_quant_forward(x, 3.0 / 127, 0, spec)
cp = lambda: y.copy_(x)
run([fwd], label="fwd only            ")
run([bwd], label="bwd only            ")
run([fwd, bwd], label="fwd,bwd (prealloc)  ")
run([fwd_alloc, bwd], label="fwd(alloc),bwd      ")
run([cp, bwd], label="copy,bwd            ")
run([cp], label="copy only           ")
run([fwd, fwd, bwd, bwd], label="fwd,fwd,bwd,bwd     ")
