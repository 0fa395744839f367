"""Host-pipeline sweep (the PCIe-bound end-to-end path): copy ceilings of this box, then chunk size x slots, with the
host time spent submitting each call (vsiq_host_pipeline_last_enqueue_ns).  profiles/r02_e2e_pipeline_sweep.log was
written while the library still had both stream layouts (mode 0 = one stream per slot, mode 1 = the shipped one)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import _lib, ops  # noqa: E402

n = 1 << 28
xh = torch.randn(n).pin_memory()
gh = torch.randn(n).pin_memory()
yh = torch.empty(n).pin_memory()
dh = torch.empty(n).pin_memory()
d, d2 = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
print("host cpus", len(os.sched_getaffinity(0)), flush=True)
for name, fn in (("H2D", lambda: d.copy_(xh, non_blocking=True)), ("D2H", lambda: yh.copy_(d, non_blocking=True))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(name, "GB/s", round(3 * 4 * n / (time.perf_counter() - t0) / 1e9, 2), flush=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for label, m in (("1 GiB buffers", n), ("first 256 MiB", 1 << 26)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            d[:m].copy_(xh[:m], non_blocking=True)
        with torch.cuda.stream(s2):
            yh[:m].copy_(d2[:m], non_blocking=True)
    torch.cuda.synchronize()
    print(f"bidirectional ({label}) GB/s each way", round(3 * 4 * m / (time.perf_counter() - t0) / 1e9, 2), flush=True)
del d, d2
for mode in ("1",):
    for lg in (21, 22, 23, 24):
        for slots in (2, 4, 8):
            p = ops.HostPipeline(1 << lg, slots)
            p.fwd_bwd(xh, gh, 0.02, 0, -128, 127, yh, dh)
            ts, qs = [], []
            for _ in range(4):
                t0 = time.perf_counter()
                p.fwd_bwd(xh, gh, 0.02, 0, -128, 127, yh, dh)
                ts.append(time.perf_counter() - t0)
                qs.append(_lib.lib.vsiq_host_pipeline_last_enqueue_ns(p._h) / 1e6)
            dt = sorted(ts)[1]
            print(f"mode {mode} chunk 2^{lg} slots {slots}: {dt * 1e3:7.2f} ms (min {min(ts) * 1e3:6.2f} max {max(ts) * 1e3:6.2f}) "
                  f"{20 * n / dt / 1e9:6.1f} GB/s metric {8 * n / dt / 1e9:5.1f} GB/s each way, submit {sorted(qs)[1]:6.2f} ms",
                  flush=True)
            p.close()
