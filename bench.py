#!/usr/bin/env python
"""bench.py -- the contract benchmark of the fake-quantization hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--log2n L]

Contract workload (BASELINE.json configs[1], the configuration the first half of the metric is quoted on): one per-tensor
W8 symmetric UniformQuantizer over a 2^28-element fp32 tensor; one step = the forward kernel + the STE backward kernel
(20 algorithmic bytes per element: read x / write y, then read g, read x / write dx).  x is 1 GiB, larger than the
126 MB L2, so nothing survives between launches.

Prints ONE JSON line (rank 0).  `value` is whole-job GB/s with inputs resident in HBM; `e2e` is the same metric
through the host-buffer entry of the C ABI (pinned host x, g -> y, dx; both PCIe directions inside the timed region);
`roofline` is the dominant kernel (STE backward, 12 B/element) against the measured HBM copy bandwidth; `cpu_baseline`
is the reference's own CPU code (baseline/_ref, kind "reference"; the torch-eager port oracle/torch_port.py when the
reference tree is absent) on the host cores over a bounded sample; `gpu_eager_baseline` is the same reference code on
the same B200 (CUDA tensors, ATen eager kernels): the GPU-vs-GPU yardstick.

The second half of BASELINE.json's metric -- YOLOv8 QAT images/s at 1/2/4/8 B200 -- rides in the same line:
  `yolo_qat`     BASELINE configs[2]: YOLOv8s, batch 64 per GPU @640, W4A8 per-channel asymmetric LSQ, channels_last,
                 weight bank, one CUDA graph per rank with the flat NCCL gradient all-reduce captured inside, input
                 prefetch; images/s, ms/step (max over ranks), the captured all-reduce's device time, the H2D time, the
                 fake-quant kernels' device time against their HBM floor, the loss;
  `calibration`  BASELINE configs[3]: YOLOv8m calibrate_qat_model over 50 global batches of 64 @640 sharded over the
                 ranks + sync_observers (one packed all_reduce(MIN) + one SUM) + reestimate_BN_stats (SyncBN-style sums);
                 images/s, the sync time, scale digests, and whether every rank ended with bit-identical scales.

`--impl reference`: the reference arm -- the reference's own UniformQuantizer on the host CPU, all host threads, the SAME
2^28-element workload (same `config`), rank 0 only.
Under torchrun (N > 1) the microbench runs the same per-GPU workload on every rank (weak scaling, no data-path collective);
every time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fake-quant fwd+bwd GB/s (20 algorithmic bytes/element)"
QMIN, QMAX, SCALE, ZP = -128, 127, 3.0 / 127, 0  # W8 symmetric, ~0.3 % of randn clipped


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--log2n", type=int, default=28, help="log2 of the tensor size (elements) per GPU")
    ap.add_argument("--cpu-log2n", type=int, default=25, help="log2 of the bounded CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 2^20..2^30 x quantiser-kind sweep (N=1 only)")
    ap.add_argument("--sweep-log2n", type=int, nargs="+", default=[20, 22, 24, 26, 28, 30])
    ap.add_argument("--no-yolo", action="store_true", help="skip the YOLOv8s data-parallel QAT step section")
    ap.add_argument("--no-calibration", action="store_true", help="skip the YOLOv8m calibration section")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the same-GPU eager yardstick")
    ap.add_argument("--yolo-steps", type=int, default=0, help="timed steps of the YOLO section (default: 16..32 from --steps)")
    ap.add_argument("--yolo-batch", type=int, default=64)
    ap.add_argument("--yolo-model", default="s")
    ap.add_argument("--calib-batches", type=int, default=50)
    ap.add_argument("--calib-model", default="m")
    return ap.parse_args()


def workload_config(args, world: int) -> dict:
    """The `config` object -- identical for the native and the reference arm (the driver compares them)."""
    n = 1 << args.log2n
    return {"workload": f"fake-quant microbench (BASELINE configs[1]): per-tensor W8 symmetric UniformQuantizer, "
                        f"2^{args.log2n} fp32 elements per GPU, step = forward + STE backward",
            "elements_per_gpu": n, "qmin": QMIN, "qmax": QMAX, "scale": SCALE,
            "l2_policy": "inputs larger than L2 (x = %d MiB > 126 MB); no flush needed" % (4 * n >> 20),
            "parallelism": f"dp{world} (independent per-GPU tensors, no data-path collective)"}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def reference_quantizer():
    """(UniformQuantizer instance of the UNMODIFIED reference, "reference") when the tree is available (baseline/_ref on
    the GPU box, /root/reference in the build container), else (None, "port")."""
    try:
        from oracle import ref_shim
        if not ref_shim.available():
            return None, "port"
        ref_shim.install()
        import importlib
        mod = importlib.import_module("quantizers.uniform")
        if not (getattr(mod, "__file__", "") or "").startswith(ref_shim.REFERENCE_ROOT):
            return None, "port"
        return mod.UniformQuantizer(8, True), "reference"
    except Exception:
        return None, "port"


def reference_step(q, x, g, scale):
    """One forward + backward of the reference's fake-quant (quantizers/uniform.py:34-56 + autograd) -- or of its
    torch-eager port when `q` is None.  Works on CPU and CUDA tensors alike."""
    if q is None:
        from oracle import torch_port
        return torch_port.fwd_bwd(x, g, scale, ZP, QMIN, QMAX)
    xr = x.detach().requires_grad_(True)
    y = q.quantize(xr, scale, ZP, False)
    y.backward(g)
    return y.detach(), xr.grad


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel: str, log2n: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, if one matches."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(f"{kernel}@2^{log2n}")
        except Exception:
            return None
    return None


def cpu_baseline(log2n: int, iters: int = 5, warm: int = 2):
    """The reference's fake-quant on the host cores (its own code when baseline/_ref is there, else the torch-eager
    port; all threads) over a bounded sample of 2^log2n elements."""
    import torch
    torch.set_num_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1
    q, kind = reference_quantizer()
    n = 1 << log2n
    torch.manual_seed(0)
    x, g = torch.randn(n), torch.randn(n)
    ts = []
    for i in range(warm + iters):
        t0 = time.perf_counter()
        reference_step(q, x, g, SCALE)
        if i >= warm:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    what = ("the reference's UniformQuantizer.quantize + autograd (baseline/_ref/quantizers/uniform.py:34-56)" if kind == "reference"
            else "torch-eager port of quantizers/uniform.py:54-55,95 + autograd (oracle/torch_port.py)")
    out = {"value": 20.0 * n / med / 1e9, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": kind,
           "sample": f"2^{log2n} elements per iteration, median of {iters} after {warm} warm-ups; {what}",
           "ms_per_step": med * 1e3, "host_cpus": os.cpu_count()}
    # the single-sweep C port (OpenMP) as a second, stronger yardstick
    try:
        import numpy as np
        import oracle
        xn, gn = x.numpy(), g.numpy()
        yo, dxo = np.empty_like(xn), np.empty_like(xn)
        tc = []
        for i in range(warm + iters):
            t0 = time.perf_counter()
            oracle.fake_quant_fwd_bwd(xn, gn, SCALE, ZP, QMIN, QMAX, yo, dxo)
            if i >= warm:
                tc.append(time.perf_counter() - t0)
        tc.sort()
        out["c_port_fused_gbs"] = 20.0 * n / tc[len(tc) // 2] / 1e9
        out["c_port_threads"] = oracle.max_threads()
    except Exception as e:  # the C port is optional
        out["c_port_fused_gbs"] = None
        out["c_port_note"] = str(e)[:80]
    return out


def gpu_eager_baseline(dev, x, g, iters: int = 5):
    """The reference's code on the SAME B200: ATen eager kernels composed in Python, autograd backward, scale as a CUDA
    tensor (ATen then divides exactly like the CPU path; a Python-float scale would multiply by 1/s).  SURVEY 8(d)."""
    import torch
    q, kind = reference_quantizer()
    scale = torch.tensor(SCALE, dtype=torch.float64, device=dev)
    try:
        for _ in range(2):
            reference_step(q, x, g, scale)
        torch.cuda.synchronize()
        evs = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            reference_step(q, x, g, scale)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        med = ts[len(ts) // 2]
        return {"value": 20.0 * x.numel() / (med * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": med, "kind": kind,
                "what": "the reference's fake-quant forward + autograd backward as ATen eager kernels on this GPU, "
                        f"2^{x.numel().bit_length() - 1} elements, CUDA-tensor scale, median of {iters}"}
    except Exception as e:
        return {"value": None, "error": str(e)[:120], "kind": kind}
    finally:
        torch.cuda.empty_cache()


def size_sweep(dev, log2ns, peak, iters: int = 10):
    """BASELINE configs[1] in full: per-tensor and per-channel W8 / W4 quantisers over 2^20 .. 2^30 fp32 elements,
    forward + backward timed separately with CUDA events (median of `iters`); tensors that fit the 126 MB L2 get a
    1 GiB flush write before every timed launch.  Extra information beside the contract keys, N = 1 only."""
    import torch
    from vsiquantization_b200 import ops

    def timed(fn, flush):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(iters):
            if flush is not None:
                flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        return ts[len(ts) // 2]

    rows = []
    flush = torch.empty(256 << 20, dtype=torch.float32, device=dev)
    C = 512
    for lg in log2ns:
        n = 1 << lg
        try:
            torch.manual_seed(0)
            x, g = torch.randn(n, device=dev), torch.randn(n, device=dev)
            y, dx = torch.empty_like(x), torch.empty_like(x)
        except torch.OutOfMemoryError:
            rows.append({"log2n": lg, "skipped": "out of memory"})
            continue
        fl = flush if 4 * n <= (256 << 20) else None
        xc, gc, yc, dxc = x.view(C, n // C), g.view(C, n // C), y.view(C, n // C), dx.view(C, n // C)
        s8 = (torch.rand(C, device=dev) * 1.5 + 0.5) * (3.0 / 127)
        s4 = (torch.rand(C, device=dev) * 1.5 + 0.5) * (3.0 / 7)
        z0, z8 = torch.zeros(C, device=dev), torch.full((C,), 8.3, device=dev)
        ds, dz = torch.empty(C, device=dev), torch.empty(C, device=dev)
        pt8, pt4 = ops.QSpec(-128, 127), ops.QSpec(0, 15)
        pc8, pc4 = ops.QSpec(-128, 127, ch_axis=0), ops.QSpec(0, 15, ch_axis=0, zp_learned=True)
        gs4 = ops.lsq_grad_scale(15, n, C)
        kinds = [
            ("per-tensor W8 symmetric UniformQuantizer, fwd + STE bwd",
             lambda: ops.fake_quant_forward(x, SCALE, ZP, pt8, out=y),
             lambda: ops.fake_quant_backward_ste(x, g, SCALE, ZP, pt8, out=dx)),
            ("per-tensor W4 asymmetric UniformQuantizer (z=8), fwd + STE bwd",
             lambda: ops.fake_quant_forward(x, 3.0 / 7, 8, pt4, out=y),
             lambda: ops.fake_quant_backward_ste(x, g, 3.0 / 7, 8, pt4, out=dx)),
            (f"per-channel (C={C}, ch_axis 0) W8 symmetric UniformQuantizer, fwd + STE bwd",
             lambda: ops.fake_quant_forward(xc, s8, z0, pc8, out=yc),
             lambda: ops.fake_quant_backward_ste(xc, gc, s8, z0, pc8, out=dxc)),
            (f"per-channel (C={C}, ch_axis 0) W4 asymmetric LSQQuantizer, fwd + LSQ bwd (dx, dscale[C], dzp[C])",
             lambda: ops.fake_quant_forward(xc, s4, z8, pc4, out=yc),
             lambda: ops.lsq_backward(xc, gc, s4, z8, pc4, gs4, ds_out=ds, dz_out=dz, want_dz=True, dx_out=dxc)),
        ]
        for name, f, b in kinds:
            tf, tb = timed(f, fl), timed(b, fl)
            gbs = 20.0 * n / ((tf + tb) * 1e-3) / 1e9
            rows.append({"log2n": lg, "quantizer": name, "fwd_ms": round(tf, 5), "bwd_ms": round(tb, 5),
                         "fwd_bwd_gbs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4),
                         "l2_flushed": fl is not None})
        del x, g, y, dx, xc, gc, yc, dxc
        torch.cuda.empty_cache()
    return rows


def code_export_bench(dev, log2n: int, peak: float, iters: int = 10):
    """Integer-code export (vsiq_quantize_codes) over 2^log2n elements: packed int4 (4.5 B/element) and int8 (5 B/element),
    codes only, CUDA events, median of `iters`."""
    import torch
    from vsiquantization_b200 import ops
    n = 1 << log2n
    torch.manual_seed(0)
    x = torch.randn(n, device=dev)
    out = {}
    for name, bits, spec, scale, zp, bpe in (("int8", 8, ops.QSpec(-128, 127), SCALE, 0, 5.0),
                                             ("int4_packed", 4, ops.QSpec(0, 15), 3.0 / 7, 8, 4.5)):
        codes = torch.empty(n // 2 if bits == 4 else n, dtype=torch.uint8 if spec.qmin >= 0 else torch.int8, device=dev)
        for _ in range(3):
            ops.quantize_codes(x, scale, zp, spec, bits, want_y=False, codes_out=codes)
        torch.cuda.synchronize()
        evs = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.quantize_codes(x, scale, zp, spec, bits, want_y=False, codes_out=codes)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        med = ts[len(ts) // 2]
        gbs = bpe * n / (med * 1e-3) / 1e9
        out[name] = {"ms": round(med, 5), "bytes_per_element": bpe, "gbs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4),
                     "note": "codes only, preallocated output"}
    out["elements"] = n
    return out


# ---------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # the reference arm runs on rank 0 alone
    import torch
    torch.set_num_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1; this arm may use every host core
    q, kind = reference_quantizer()
    log2n = args.log2n
    try:  # the full 2^28 workload needs ~12 GB of host memory for the autograd intermediates
        import psutil
        while log2n > 20 and psutil.virtual_memory().available < 14 * 4 * (1 << log2n):
            log2n -= 1
    except Exception:
        pass
    n = 1 << log2n
    torch.manual_seed(0)
    x, g = torch.randn(n), torch.randn(n)
    for _ in range(args.warmup):
        reference_step(q, x, g, SCALE)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_step(q, x, g, SCALE)
    dt = time.perf_counter() - t0
    value = 20.0 * n * args.steps / dt / 1e9
    what = ("the unmodified reference's UniformQuantizer.quantize + autograd backward (baseline/_ref)" if kind == "reference"
            else "torch-eager port of the reference's op sequence (oracle/torch_port.py; reference tree not available)")
    sample = (f"each step = fwd+bwd over {'the whole' if log2n == args.log2n else 'a bounded sample of'} 2^{log2n} elements of "
              f"the 2^{args.log2n}-element workload; {what}, all host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": kind,
                             "sample": sample, "host_cpus": os.cpu_count(), "elements_per_step": n},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ model sections
def yolo_section(args, world):
    """BASELINE configs[2] through benchmarks/yolo_qat.py (the same code path as its command line)."""
    from benchmarks import yolo_qat
    steps = args.yolo_steps or min(max(args.steps, 16), 32)
    argv = ["--model", args.yolo_model, "--batch", str(args.yolo_batch), "--imgsz", "640", "--steps", str(steps),
            "--warmup", "3", "--w-bits", "4", "--a-bits", "8", "--asym", "--per-channel", "--lsq", "--channels-last",
            "--weight-bank", "--cuda-graph", "--profile-steps", "2"]
    r = yolo_qat.run(yolo_qat.parse(argv))
    fq = r.get("fake_quant", {})
    out = {"config": f"YOLOv8{args.yolo_model} QAT, batch {args.yolo_batch}/GPU @640, W4A8 per-channel asymmetric LSQQuantizer "
                     f"(learnable step + zero-point), channels_last, weight bank, CUDA graph + captured flat NCCL all-reduce, "
                     f"input prefetch; synthetic uint8 images, loss = sum of mean(out^2)",
           "images_per_s": r["images_per_s"], "ms_per_step": r["ms_per_step"], "steps": r["steps"], "n_gpus": r["n_gpus"],
           "allreduce_ms": r.get("allreduce_ms"), "allreduce_alone": r.get("allreduce_alone"),
           "h2d_ms": r.get("h2d_ms"), "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": r["d2h_bytes_per_step"],
           "fake_quant_ms": fq.get("kernel_ms_per_step"), "hbm_floor_ms": fq.get("hbm_floor_ms_per_step"),
           "fake_quant_frac_of_floor": fq.get("fraction_of_hbm_floor"), "fake_quant_kernels": fq.get("kernels"),
           "device_ms_per_step": fq.get("device_ms_per_step"), "fake_quant_gb_per_step": fq.get("algorithmic_gb_per_step_per_gpu"),
           "loss": r["loss"], "ms_per_step_by_rank": r.get("ms_per_step_by_rank"), "peak_mem_gb": r["peak_mem_gb"],
           "fused_layers": r["fused_layers"], "prefetch": r["prefetch"], "weight_bank": r["weight_bank"],
           "slow_paths": r.get("slow_paths")}
    if fq.get("profile_error"):
        out["profile_error"] = fq["profile_error"]
    return out


def calibration_section(args, world):
    """BASELINE configs[3] through benchmarks/calibration.py."""
    from benchmarks import calibration
    batches = max(args.calib_batches, world)
    r = calibration.run(calibration.parse(["--model", args.calib_model, "--batch", "64", "--imgsz", "640", "--batches",
                                           str(batches), "--channels-last"]))
    r["config"] = (f"YOLOv8{args.calib_model}, {batches} global batches of 64 @640 sharded over {world} rank(s), is_fuse_bn=False, "
                   f"calibrate_qat_model + sync_observers (packed all_reduce MIN + SUM) + reestimate_BN_stats (SyncBN-style sums)")
    return r


# -------------------------------------------------------------------------------------- native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- vsiquantization_b200 has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from vsiquantization_b200 import _lib, ops

    n = 1 << args.log2n
    torch.manual_seed(rank)
    x = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    spec = ops.QSpec(QMIN, QMAX)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # outputs are allocated once: a QAT step reuses its activation / gradient buffers through the caching allocator
    y = torch.empty_like(x)
    dx = torch.empty_like(x)

    def step():
        ops.fake_quant_forward(x, SCALE, ZP, spec, out=y)
        ops.fake_quant_backward_ste(x, g, SCALE, ZP, spec, out=dx)

    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clocks:
        sampler.start()  # started before the warm-up so nvidia-smi's own start-up is outside the timed region
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if rank == 0 and not args.no_clocks:
        time.sleep(0.3)
    launches0 = _lib.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        ops.fake_quant_forward(x, SCALE, ZP, spec, out=y)
        ev[3 * i + 1].record()
        ops.fake_quant_backward_ste(x, g, SCALE, ZP, spec, out=dx)
        ev[3 * i + 2].record()
        ev[3 * i + 3].record()
    barrier()
    launches = _lib.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[3 * args.steps])
    fwd_ms = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.steps)) / args.steps
    bwd_ms = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.steps)) / args.steps
    step_ms = sorted(ev[3 * i].elapsed_time(ev[3 * i + 3]) for i in range(args.steps))
    clocks = sampler.stop() if (rank == 0 and not args.no_clocks) else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = 20.0 * n * args.steps * world / (total_ms * 1e-3) / 1e9

    # correctness spot check inside the bench (first 2^16 elements against the oracle), rank 0
    check = None
    if rank == 0:
        try:
            import oracle
            k = 1 << 16
            yo = oracle.fake_quant_fwd(x[:k].cpu().numpy(), SCALE, ZP, QMIN, QMAX)
            dxo = oracle.fake_quant_bwd(x[:k].cpu().numpy(), g[:k].cpu().numpy(), SCALE, ZP, QMIN, QMAX, want_ds=False)[0]
            check = bool((y[:k].cpu().numpy().view("uint32") == yo.view("uint32")).all() and
                         (dx[:k].cpu().numpy().view("uint32") == dxo.view("uint32")).all())
        except Exception as e:
            check = f"oracle unavailable: {str(e)[:60]}"
    del y, dx

    # ---- end to end through the host-buffer entry point (pinned host memory, H2D + D2H inside the timed region)
    e2e = None
    e2e_ok, e2e_err = not args.no_e2e, None
    if e2e_ok:
        e2e_n = n
        xh = gh = yh = dh = None
        try:  # 16 bytes of page-locked host memory per element and rank
            xh = torch.empty(e2e_n, dtype=torch.float32, pin_memory=True)
            gh = torch.empty(e2e_n, dtype=torch.float32, pin_memory=True)
            yh = torch.empty(e2e_n, dtype=torch.float32, pin_memory=True)
            dh = torch.empty(e2e_n, dtype=torch.float32, pin_memory=True)
        except Exception as e:  # noqa: BLE001 -- reported in the line instead of losing it
            e2e_ok, e2e_err = False, str(e)[:120]
        if world > 1:  # every rank or none: the timed region below holds collectives
            flag = torch.tensor([1.0 if e2e_ok else 0.0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            e2e_ok = bool(flag.item() == 1.0)
        if not e2e_ok:
            e2e = {"error": "pinned host buffers unavailable on some rank: " + (e2e_err or "another rank failed")}
            del xh, gh, yh, dh
    if e2e_ok:
        xh.copy_(x)
        gh.copy_(g)
        pipe = ops.HostPipeline(chunk_elems=1 << 23, n_slots=4, device=dev)
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            pipe.fwd_bwd(xh, gh, SCALE, ZP, QMIN, QMAX, yh, dh)
        barrier()
        l0 = _lib.launch_count
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.fwd_bwd(xh, gh, SCALE, ZP, QMIN, QMAX, yh, dh)  # returns when y, dx have landed on the host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e_launches = _lib.launch_count - l0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": 20.0 * e2e_n * e2e_steps * world / dt / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": 8 * e2e_n, "d2h_bytes_per_step": 8 * e2e_n, "steps": e2e_steps,
               "ms_per_step": dt / e2e_steps * 1e3, "gpu_launches": e2e_launches,
               "pcie_gbs_each_way": 8.0 * e2e_n * e2e_steps / dt / 1e9,
               "api": "vsiq_host_pipeline_fwd_bwd (ops.HostPipeline): 8 Mi-element chunks through 4 staging slots, H2D / fused "
                      "fwd+bwd kernel / D2H on three event-linked streams",
               "host_submit_ms_per_step": round(pipe.last_submit_ms, 3)}
        pipe.close()
        # what the host side can deliver by itself: concurrent H2D + D2H copies of the same pinned buffers, all ranks at
        # once, no kernel -- the PCIe / host-DRAM ceiling the e2e number sits under (8 ranks share one socket's memory)
        try:
            m = min(e2e_n, 1 << 26)
            dbuf, dbuf2 = torch.empty(m, device=dev), torch.empty(m, device=dev)
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
            barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                with torch.cuda.stream(s1):
                    dbuf.copy_(xh[:m], non_blocking=True)
                with torch.cuda.stream(s2):
                    yh[:m].copy_(dbuf2, non_blocking=True)
            torch.cuda.synchronize()
            dtc = time.perf_counter() - t0
            tc_ = torch.tensor([dtc], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tc_, op=dist.ReduceOp.MAX)
            e2e["copy_only_gbs_each_way_per_gpu"] = 4.0 * m * 4 / float(tc_.item()) / 1e9
            e2e["host_cpus_visible"] = host_threads()
            del dbuf, dbuf2
        except Exception as e:
            e2e["copy_only_note"] = str(e)[:80]
        del xh, gh, yh, dh

    # ---- the same reference code on the same GPU (eager ATen kernels): GPU-vs-GPU yardstick, rank 0
    gpu_eager = None
    if rank == 0 and not args.no_gpu_eager:
        gpu_eager = gpu_eager_baseline(dev, x, g)
    del x, g
    torch.cuda.empty_cache()

    # ---- BASELINE configs[2] and [3]: every rank takes part (NCCL collectives inside)
    def guarded(fn):
        try:
            return fn(args, world)
        except Exception as e:  # never lose the contract line to a model-section failure: say what failed
            import traceback
            return {"error": f"{type(e).__name__}: {str(e)[:300]}", "where": traceback.format_exc(limit=3)[-400:]}
        finally:
            torch.cuda.empty_cache()
    yolo = None if args.no_yolo else guarded(yolo_section)
    calib = None if args.no_calibration else guarded(calibration_section)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    achieved = 12.0 * n / (bwd_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "step_ms": {"median": step_ms[len(step_ms) // 2], "min": step_ms[0], "max": step_ms[-1]},
        "config": workload_config(args, world),
        "roofline": {"bound": "hbm", "kernel": "fq_bwd_ste_kernel<256,8> (read g, read x, write dx: 12 B/element)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": ncu_traffic("fq_bwd_ste_kernel", args.log2n),
                     "algorithmic_bytes_per_launch": 12 * n, "avg_launch_ms": bwd_ms,
                     "fwd_kernel": {"achieved": 8.0 * n / (fwd_ms * 1e-3) / 1e9, "avg_launch_ms": fwd_ms,
                                    "frac": 8.0 * n / (fwd_ms * 1e-3) / 1e9 / peak}},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "oracle_spot_check_bit_exact": check,
        "yolo_qat": yolo, "calibration": calib, "gpu_eager_baseline": gpu_eager,
    }
    if gpu_eager and gpu_eager.get("value"):
        line["vs_gpu_eager"] = value / world / gpu_eager["value"]
    if world == 1 and not args.no_sweep:
        line["sweep"] = size_sweep(dev, args.sweep_log2n, peak)
        big = [r["frac_of_peak"] for r in line["sweep"] if r.get("log2n", 0) >= 26 and "frac_of_peak" in r]
        line["sweep_worst_frac_of_peak_ge_2p26"] = min(big) if big else None
        try:
            line["code_export"] = code_export_bench(dev, min(args.log2n, 28), peak)
        except Exception as e:
            line["code_export"] = {"error": str(e)[:120]}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.cpu_log2n)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
