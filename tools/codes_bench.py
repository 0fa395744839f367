"""Integer-code export (vsiq_quantize_codes, codes only) timed with CUDA events, median of 10, inputs larger than L2.

The schedule A/B this script ran while the kernel had experiment knobs is kept as profiles/r02_ab_code_export.log."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsiquantization_b200 import ops  # noqa: E402

PEAK = 6531.9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda:0")
CASES = (("int8", 8, ops.QSpec(-128, 127), 3.0 / 127, 0, 5.0), ("int4", 4, ops.QSpec(0, 15), 3.0 / 7, 8, 4.5))


def timeit(x, codes, bits, spec, s, z, iters=10):
    for _ in range(3):
        ops.quantize_codes(x, s, z, spec, bits, want_y=False, codes_out=codes)
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.quantize_codes(x, s, z, spec, bits, want_y=False, codes_out=codes)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    if "--once" in sys.argv:  # the command ncu profiles: one int8 and one packed-int4 launch at 2^28 after a warm-up each
        n = 1 << 28
        x = torch.randn(n, device=dev)
        for name, bits, spec, s, z, bpe in CASES:
            codes = torch.empty(n // 2 if bits == 4 else n, dtype=torch.uint8 if spec.qmin >= 0 else torch.int8, device=dev)
            for _ in range(2):
                ops.quantize_codes(x, s, z, spec, bits, want_y=False, codes_out=codes)
        torch.cuda.synchronize()
        return
    for log2n in (26, 28):
        n = 1 << log2n
        torch.manual_seed(0)
        x = torch.randn(n, device=dev)
        for name, bits, spec, s, z, bpe in CASES:
            codes = torch.empty(n // 2 if bits == 4 else n, dtype=torch.uint8 if spec.qmin >= 0 else torch.int8, device=dev)
            ms = timeit(x, codes, bits, spec, s, z)
            gbs = bpe * n / (ms * 1e-3) / 1e9
            print(f"n=2^{log2n} {name}: {ms * 1e3:8.1f} us {gbs:7.1f} GB/s {gbs / PEAK:.3f} of peak", flush=True)


if __name__ == "__main__":
    main()
