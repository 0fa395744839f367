"""Calibration-pass benchmark (BASELINE.json configs[3]): calibrate_qat_model + reestimate_BN_stats over `--batches`
synthetic batches on YOLOv8m, calibration batches sharded across ranks, observer statistics all-reduced (MIN/MAX, exact)
and BN moments all-reduced per layer (SyncBN-style).

    python -m benchmarks.calibration --model m --batch 64 --imgsz 640 --batches 50
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m benchmarks.calibration ...

Reports seconds and images/s for both phases, the number of host synchronisations the calibration forward passes made
(0: everything stays on the device), and a digest of the post-calibration scales so runs at different world sizes can
be compared (MIN/MAX are exact, so weight scales are identical at every N; activation extrema depend on which batches
exist, which is the same set at every N here)."""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="m")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--batches", type=int, default=50, help="GLOBAL number of calibration batches")
    ap.add_argument("--channels-last", action="store_true",
                    help="NHWC weights / activations (cuDNN's native layout on sm_100); observers and the BN moments "
                         "pass walk that memory in place")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    from vsiquantization_b200 import _lib
    from vsiquantization_b200.modules.fuse import fuse_modules_unified
    from vsiquantization_b200.modules.fuse_config import FuseConfig, create_fuse_config_manager
    from vsiquantization_b200.nets import yolov8
    from vsiquantization_b200.parallel import quantization_managers, sync_observers
    from vsiquantization_b200.utils.estimate_bn import reestimate_BN_stats
    from vsiquantization_b200.utils.quantize_manager import calibrate_qat_model

    torch.manual_seed(0)
    model = getattr(yolov8, f"yolo_v8_{args.model}")(num_classes=20).to(dev)
    cfg = create_fuse_config_manager(FuseConfig(is_fuse_bn=False, bits_w=8, bits_a=8))  # BN kept: needed for re-estimation
    model = fuse_modules_unified(model, [["conv", "bn", "relu"]], config_manager=cfg)
    if args.channels_last:
        model.to(memory_format=torch.channels_last)

    # the same GLOBAL set of batches at every world size; rank r takes batches r, r+world, ...
    def batch(i):
        g = torch.Generator().manual_seed(1000 + i)
        return torch.randint(0, 256, (args.batch, 3, args.imgsz, args.imgsz), generator=g, dtype=torch.uint8).pin_memory()
    mine = [(batch(i), None) for i in range(rank, args.batches, world)]

    def data_calib(m, loader, device):
        m.eval()
        with torch.no_grad():
            for imgs, _ in loader:
                m(imgs.to(device, non_blocking=True).float() / 255.0)

    data_calib(model, mine[:1], dev)  # warm-up (cuDNN plans, allocator) outside the timing
    for _, q in quantization_managers(model):  # forget the warm-up observation
        q.observer._state = None
        q._call_stats.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    t0 = time.perf_counter()
    calibrate_qat_model(model, mine, data_calib, dev)
    rows = sync_observers(model) if world > 1 else sum(1 for _ in quantization_managers(model))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_cal = time.perf_counter() - t0
    launches_cal = _lib.launch_count - l0
    t0 = time.perf_counter()
    reestimate_BN_stats(model, mine, num_batches=len(mine), sync=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_bn = time.perf_counter() - t0
    h = hashlib.sha256()
    for name, q in quantization_managers(model):
        if name.endswith("weight_quantizer"):
            h.update(repr((q.scale, q.zero_point)).encode())
    ha = hashlib.sha256()
    for name, q in quantization_managers(model):
        if name.endswith("activation_quantizer"):
            ha.update(repr((q.scale, q.zero_point)).encode())
    bn0 = next(m for m in model.modules() if hasattr(m, "bn")).bn
    if rank == 0:
        print(json.dumps({"model": f"yolov8{args.model}", "n_gpus": world, "global_batches": args.batches,
                          "batch": args.batch, "imgsz": args.imgsz, "channels_last": args.channels_last,
                          "observer_rows_synced": rows,
                          "calibration_s": t_cal, "calibration_images_per_s": args.batches * args.batch / t_cal,
                          "bn_reestimate_s": t_bn, "bn_reestimate_images_per_s": args.batches * args.batch / t_bn,
                          "vsiq_launches_calibration": launches_cal,
                          "weight_scales_sha256": h.hexdigest()[:16], "activation_scales_sha256": ha.hexdigest()[:16],
                          "bn0_running_mean_head": [round(float(v), 6) for v in bn0.running_mean[:3]]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
