"""Subprocess body of tests/test_gpu_reference.py: the UNMODIFIED reference's own code path

    fuse_modules_unified -> calibrate_qat_model -> activate_learning_qparam -> activate_quantizer -> model.to(cuda)
    -> forward + backward                                      (yolov8_qat.py:63-92,225-263)

run twice on the same GPU with the same seeds: once with the reference's own plugins, once after
``vsiquantization_b200.dropin.install_plugins()`` replaced the two plugin classes in the reference's CLASS_REGISTRY
(quantizers/quantization_manager.py:41-42 then builds the native kernels by name).  Everything else -- manager, fused
layers, fuse pass, control API, network -- is the reference's file in both runs.  Prints one JSON object.
TEST INFRASTRUCTURE (imports oracle.ref_shim)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ref_shim  # noqa: E402

ref_shim.install()
import importlib  # noqa: E402

registry = importlib.import_module("utils.registry")
ry = importlib.import_module("nets.yolov8")
rfuse = importlib.import_module("modules.fuse")
rcfg = importlib.import_module("modules.fuse_config")
rctl = importlib.import_module("utils.quantize_manager")
runi = importlib.import_module("quantizers.uniform")
robs = importlib.import_module("observers.minmax")
for m in (registry, ry, rfuse, rcfg, rctl, runi, robs):
    assert m.__file__.startswith(ref_shim.REFERENCE_ROOT), m.__file__

torch.backends.cudnn.deterministic = True
torch.backends.cudnn.benchmark = False
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
IMG = int(os.environ.get("TIER1_IMG", "96"))


def bits(t):
    return t.detach().float().cpu().numpy().view(np.uint32)


def run(native: bool, bits_w: int, bits_a: int):
    if native:
        import vsiquantization_b200.dropin as dropin
        dropin.install_plugins()
    else:
        registry.CLASS_REGISTRY["UniformQuantizer"] = runi.UniformQuantizer
        registry.CLASS_REGISTRY["MinMaxObserver"] = robs.MinMaxObserver
    torch.manual_seed(0)
    model = ry.yolo_v8_n(20)
    # non-trivial BatchNorm statistics so the fold matters
    g = torch.Generator().manual_seed(1)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    cfg = rcfg.create_fuse_config_manager(rcfg.FuseConfig(bits_w=bits_w, bits_a=bits_a))
    model = rfuse.fuse_modules_unified(model, [["conv", "bn", "relu"]], is_trace=False, config_manager=cfg)
    fused = [(n, m) for n, m in model.named_modules() if hasattr(m, "weight_quantizer")]
    plug = type(fused[0][1].weight_quantizer.quantizer).__module__
    gcal = torch.Generator().manual_seed(2)
    calib = [(torch.randint(0, 256, (2, 3, IMG, IMG), generator=gcal, dtype=torch.uint8), None) for _ in range(2)]

    def data_calib(m, loader, device):  # yolov8_qat.py:42-52
        m.eval()
        m.to(device)
        for imgs, _ in loader:
            m(imgs.to(device, non_blocking=True).float() / 255.0)
        m.train()

    rctl.calibrate_qat_model(model, calib, data_calib, "cuda")
    calibrated = {n: [float(m.weight_quantizer.scale), int(m.weight_quantizer.zero_point),
                      float(m.activation_quantizer.scale), int(m.activation_quantizer.zero_point)] for n, m in fused}
    rctl.activate_learning_qparam(model, use_init=True)
    rctl.activate_quantizer(model)
    model.to(dev)  # the learnable scales become CUDA tensors (yolov8_qat.py:111)
    model.train()
    init = {n: [float(m.weight_quantizer.scale.detach()), float(m.activation_quantizer.scale.detach())] for n, m in fused}
    # |terms| of every scale gradient's sums, gs * sum |g| (|q - z| + |x/s|): the yardstick its fp32 error is judged by
    mass = {}
    if native:
        for n, m in fused:
            for attr in ("weight_quantizer", "activation_quantizer"):
                mgr = getattr(m, attr)
                orig = mgr.quantizer.quantize

                def wrapped(xq, scale, zp, learn=False, _orig=orig, _key=f"{n}.{attr}.scale", _q=mgr.quantizer):
                    yq = _orig(xq, scale, zp, learn)
                    if yq.requires_grad and _key not in mass:
                        def grab(gq, xq=xq.detach(), yq=yq.detach(), scale=scale):
                            s32 = float(scale.detach().float())
                            gs = (_q.qmax * xq.numel()) ** -0.5
                            mass[_key] = gs / abs(s32) * float((gq.double().abs() * (yq.double().abs() + xq.double().abs())).sum())
                        yq.register_hook(grab)
                    return yq
                mgr.quantizer.quantize = wrapped
    acts = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, n=n: acts.__setitem__(n, o.detach())) for n, m in fused]
    gx = torch.Generator().manual_seed(3)
    x = (torch.randint(0, 256, (2, 3, IMG, IMG), generator=gx, dtype=torch.uint8).to(dev).float() / 255.0).requires_grad_(True)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True)
    losses = []
    for step in range(3):
        opt.zero_grad(set_to_none=True)
        outs = model(x)
        loss = sum((o.float() ** 2).mean() for o in outs)
        loss.backward()
        losses.append(float(loss.detach()))
        if step == 0:
            first = {"acts": {n: bits(a) for n, a in acts.items()}, "dx": bits(x.grad),
                     "grads": {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}}
        opt.step()
    for h in hooks:
        h.remove()
    return {"mass": mass, "plugin_module": plug, "calibrated": calibrated, "init": init, "losses": losses, "first": first,
            "n_fused": len(fused), "scale_devices": sorted({str(m.weight_quantizer.scale.device) for _, m in fused})}


def compare(bits_w, bits_a):
    a = run(False, bits_w, bits_a)
    b = run(True, bits_w, bits_a)
    res = {"bits": [bits_w, bits_a], "reference_plugin": a["plugin_module"], "native_plugin": b["plugin_module"],
           "n_fused": a["n_fused"], "scale_devices": b["scale_devices"],
           "calibrated_equal": a["calibrated"] == b["calibrated"], "init_equal": a["init"] == b["init"],
           "losses_reference": a["losses"], "losses_native": b["losses"]}
    fa, fb = a["first"], b["first"]
    bad_layers = [n for n in fa["acts"] if not np.array_equal(fa["acts"][n], fb["acts"][n])]
    res["layers"] = len(fa["acts"])
    res["layers_with_output_mismatch"] = bad_layers[:5]
    res["dx_equal"] = bool(np.array_equal(fa["dx"], fb["dx"]))
    wbad, sworst = [], 0.0
    for n, ga in fa["grads"].items():
        gb = fb["grads"][n]
        if n.endswith(("quantizer.scale", "quantizer.zero_point")):
            da, db = float(ga), float(gb)
            sworst = max(sworst, abs(da - db) / b["mass"][n])
        elif not np.array_equal(bits(ga), bits(gb)):
            wbad.append(n)
    res["weight_bias_grads_with_mismatch"] = wbad[:5]
    res["n_grads"] = len(fa["grads"])
    res["scale_grad_worst_err_over_mass"] = sworst
    res["n_scale_grads"] = len(b["mass"])
    return res


if __name__ == "__main__":
    out = [compare(8, 8), compare(4, 8)]
    print("TIER1_JSON " + json.dumps(out))
