"""Host-pipeline chunk/slot sweep (PCIe-bound end-to-end path)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import ops
n = 1 << 28
xh = torch.randn(n).pin_memory(); gh = torch.randn(n).pin_memory()
yh = torch.empty(n).pin_memory(); dh = torch.empty(n).pin_memory()
# raw PCIe ceilings
d = torch.empty(n, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(xh, non_blocking=True)), ("D2H", lambda: yh.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(name, "GB/s", 3 * 4 * n / (time.perf_counter() - t0) / 1e9)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): d.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2): yh.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("bidirectional GB/s each way", 3 * 4 * n / (time.perf_counter() - t0) / 1e9)
for chunk in (1 << 20, 1 << 22, 1 << 24, 1 << 25):
    for slots in (2, 3, 4, 8):
        p = ops.HostPipeline(chunk, slots)
        p.fwd_bwd(xh, gh, 0.02, 0, -128, 127, yh, dh)
        t0 = time.perf_counter()
        for _ in range(3): p.fwd_bwd(xh, gh, 0.02, 0, -128, 127, yh, dh)
        dt = (time.perf_counter() - t0) / 3
        print(f"chunk 2^{chunk.bit_length()-1} slots {slots}: {dt*1e3:7.2f} ms  {20*n/dt/1e9:6.1f} GB/s metric, {8*n/dt/1e9:5.1f} GB/s each way")
        p.close()
