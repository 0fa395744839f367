"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE -- build-container only (needs /root/reference).  The
reference ships no tests or golden vectors, so these fixtures ARE the parity
pin: each file holds seeded inputs plus the outputs the reference's own
classes produced for them (torch CPU, see the ``meta`` entry of every file).

    python -m oracle.gen_golden            # rewrites tests/golden/*.npz

Reference entry points exercised (file:line):
  UniformQuantizer.quantize            quantizers/uniform.py:34-56
  ScaleGradient / RoundStraightThrough quantizers/uniform.py:242-271
  MinMaxObserver                       observers/minmax.py:25-88
  QuantizationManager                  quantizers/quantization_manager.py:33-114
  LSQFakeQuantize (per channel)        quantizers/lsq_module.py:73-384
  FunLSQ                               quantizers/uniform.py:105-151
  ConvBnReLU / LinearBnReLU BN fold    modules/fused.py:92-108, 286-300
  reestimate_BN_stats                  utils/estimate_bn.py:38-101
  fuse_modules_unified + calibrate_qat_model + activate_learning_qparam
                                       modules/fuse.py:254-277, utils/quantize_manager.py:4-66
  the same sequence on YOLOv8n         nets/yolov8.py (57 fused layers, 64x64 input)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from quantizers.uniform import UniformQuantizer, FunLSQ  # noqa: E402  (reference)
from observers.minmax import MinMaxObserver  # noqa: E402
from quantizers.quantization_manager import QuantizationManager  # noqa: E402
from quantizers.lsq_module import LSQFakeQuantize  # noqa: E402
from modules.fused import ConvBnReLU, LinearBnReLU  # noqa: E402
from modules.fuse import fuse_modules_unified  # noqa: E402
from modules.fuse_config import FuseConfig, create_fuse_config_manager  # noqa: E402
from utils.quantize_manager import calibrate_qat_model, activate_learning_qparam, activate_quantizer  # noqa: E402
from utils.estimate_bn import reestimate_BN_stats  # noqa: E402

sys.path.insert(0, GOLD)
from tiny_model import make_tiny  # noqa: E402

META = np.array(f"reference=tranngocduvnvp/VSIQuantization torch={torch.__version__} device=cpu "
                f"threads={torch.get_num_threads()}")


def save(name, **arrays):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, meta=META, **arrays)
    print(f"  {name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


def adversarial(scale, zp, qmin, qmax, n_rand, gen, dtype=np.float32):
    """Ties at k+1/2, both clamp bounds, signed zeros, denormals, inf, nan + random bulk."""
    s = np.float32(scale)
    ks = np.arange(qmin - 3, qmax + 4, dtype=np.float64)
    ties = ((ks + 0.5 - zp) * float(s)).astype(np.float32)
    ints = ((ks - zp) * float(s)).astype(np.float32)
    near = np.concatenate([np.nextafter(ties, np.float32(np.inf)), np.nextafter(ties, np.float32(-np.inf))])
    special = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-40, -1e-40, 1.17549435e-38, -1.17549435e-38,
                        np.inf, -np.inf, np.nan, 3.4e38, -3.4e38, 1e-30, -1e-30,
                        -0.3 * float(s), 0.3 * float(s), -0.5 * float(s), 0.5 * float(s)], dtype=np.float32)
    bulk = torch.randn(n_rand, generator=gen).numpy() * np.float32(3.0 * 127 * float(s) / 3.0 if qmax >= 127 else
                                                                  float(s) * (qmax - qmin) / 4.0)
    return np.concatenate([ties, ints, near, special, bulk.astype(np.float32)]).astype(dtype)


def gen_uniform_fixed(gen):
    """Non-learning path: Python-float scale, Python-int zero_point (quantization_manager.py:69-71,88)."""
    out = {}
    cases = []
    for bits in (8, 4, 2):
        for sym in (True, False):
            q = UniformQuantizer(bits, sym)
            for si, scale in enumerate((3.0 / (2 ** (bits - 1) - 1 if sym else 2 ** bits - 1) * 0.37, 2.0 ** -5)):
                zp = 0 if sym else (2 ** (bits - 1)) - (1 if si else 0)
                tag = f"b{bits}_{'sym' if sym else 'asym'}_{si}"
                x = adversarial(scale, zp, q.qmin, q.qmax, 3000, gen)
                xt = torch.tensor(x, requires_grad=True)
                y = q.quantize(xt, float(scale), int(zp), False)
                g = torch.randn(x.shape, generator=gen)
                y.backward(g)
                with torch.no_grad():
                    codes = q.discreate_tensor(xt, float(scale), int(zp), q.qmin, q.qmax)
                out[f"{tag}_x"] = x
                out[f"{tag}_g"] = g.numpy()
                out[f"{tag}_y"] = y.detach().numpy()
                out[f"{tag}_codes"] = codes.numpy()
                out[f"{tag}_dx"] = xt.grad.numpy()
                out[f"{tag}_qp"] = np.array([scale, zp, q.qmin, q.qmax, bits, int(sym)], dtype=np.float64)
                cases.append(tag)
    out["cases"] = np.array(cases)
    save("uniform_fixed", **out)


def gen_uniform_learned(gen):
    """Learning path: 0-dim fp64 Parameter scale (quantization_manager.py:99,112), zp = 0 (sym),
    and the intended-but-unreachable asymmetric learnable zero-point (uniform.py:50-52) driven directly."""
    out = {}
    cases = []
    for bits in (8, 4):
        for sym in (True, False):
            for zi, zf in enumerate(((0.0,) if sym else (float(2 ** (bits - 1)) + 1e-9, 3.4, -7.25))):
                q = UniformQuantizer(bits, sym)
                scale = np.float64(0.0173 if bits == 8 else 0.21)
                tag = f"b{bits}_{'sym' if sym else 'asym'}_{zi}"
                x = adversarial(scale, round(zf), q.qmin, q.qmax, 5000, gen)
                # the reference's sums see inf/nan; keep them out of the learned-path fixture
                x = x[np.isfinite(x) & (np.abs(x) < 1e30)]
                xt = torch.tensor(x, requires_grad=True)
                sp = torch.nn.Parameter(torch.tensor(scale))  # float64, 0-dim
                assert sp.dtype == torch.float64
                if sym:
                    zpar = 0
                else:
                    zpar = torch.nn.Parameter(torch.tensor(zf, dtype=torch.float32))
                y = q.quantize(xt, sp, zpar, True)
                g = torch.randn(x.shape, generator=gen)
                y.backward(g)
                out[f"{tag}_x"] = x
                out[f"{tag}_g"] = g.numpy()
                out[f"{tag}_y"] = y.detach().numpy()
                out[f"{tag}_dx"] = xt.grad.numpy()
                out[f"{tag}_ds"] = sp.grad.numpy().astype(np.float64).reshape(1)
                out[f"{tag}_dz"] = (zpar.grad.numpy().astype(np.float64).reshape(1) if not sym else np.zeros(1))
                out[f"{tag}_qp"] = np.array([scale, zf, q.qmin, q.qmax, bits, int(sym),
                                             q.calculate_grad_scale(xt)], dtype=np.float64)
                cases.append(tag)
    # calib_grad_scale as a [C] tensor (utils/estimate_bn.py:136): autograd sum-reduces it onto the 0-dim scale
    q = UniformQuantizer(8, True)
    q.calib_grad_scale = torch.tensor([0.5, 2.0, 1.25])
    x = torch.randn(2, 3, 4, 5, generator=gen).permute(0, 2, 3, 1).contiguous()  # last dim broadcasts against [3]
    xt = x.clone().requires_grad_(True)
    sp = torch.nn.Parameter(torch.tensor(np.float64(0.02)))
    y = q.quantize(xt, sp, 0, True)
    g = torch.randn(x.shape, generator=gen)
    y.backward(g)
    out["cgs_x"] = x.numpy()
    out["cgs_g"] = g.numpy()
    out["cgs_y"] = y.detach().numpy()
    out["cgs_dx"] = xt.grad.numpy()
    out["cgs_ds"] = sp.grad.numpy().reshape(1)
    out["cgs_vec"] = q.calib_grad_scale.numpy()
    out["cases"] = np.array(cases)
    save("uniform_learned", **out)


def gen_funlsq(gen):
    """Dead-code canonical LSQ (uniform.py:105-151) -- offered as mask_mode=1."""
    x = adversarial(0.05, 0, -8, 7, 4000, gen)
    x = x[np.isfinite(x) & (np.abs(x) < 1e30)]
    xt = torch.tensor(x, requires_grad=True)
    s = torch.tensor([0.05], requires_grad=True)
    gsc = 1.0 / np.sqrt(7 * x.size)
    y = FunLSQ.apply(xt, s, gsc, -8, 7)
    g = torch.randn(x.shape, generator=gen)
    y.backward(g)
    save("funlsq", x=x, g=g.numpy(), y=y.detach().numpy(), dx=xt.grad.numpy(),
         ds=s.grad.numpy().astype(np.float64), qp=np.array([0.05, 0, -8, 7, gsc]))


def gen_observer(gen):
    """MinMaxObserver sequences (observers/minmax.py) incl. the 0-initialised state and NaN handling."""
    out = {}
    seqs = {
        "pos_only": [torch.rand(257, generator=gen) + 0.5, torch.rand(33, generator=gen) * 3 + 0.1],
        "mixed": [torch.randn(1000, generator=gen), torch.randn(4, 3, 5, 5, generator=gen) * 4, torch.randn(7, generator=gen) * 0.1],
        "neg_only": [-torch.rand(100, generator=gen) - 1.0],
        "with_nan": [torch.randn(64, generator=gen), torch.tensor([1.0, float("nan"), -50.0, 70.0]), torch.randn(8, generator=gen) * 9],
        "with_inf": [torch.tensor([1.0, float("inf"), -2.0])],
        "zeros": [torch.zeros(10)],
    }
    names = []
    for name, seq in seqs.items():
        for sym in (True, False):
            for bits in (8, 4, 2):
                obs = MinMaxObserver(sym, bits)
                trace = []
                for t in seq:
                    s, z = obs.forward(t)
                    trace.append([obs.min_val, obs.max_val, s, float(z)])
                tag = f"{name}_{'sym' if sym else 'asym'}_b{bits}"
                out[f"{tag}_trace"] = np.array(trace, dtype=np.float64)
                names.append(tag)
        for i, t in enumerate(seq):
            out[f"{name}_in{i}"] = t.numpy()
        out[f"{name}_n"] = np.array(len(seq))
    out["cases"] = np.array(names)
    save("observer", **out)


def gen_manager(gen):
    """QuantizationManager calibration flow + LSQ init (quantization_manager.py:55-114)."""
    out = {}
    tags = []
    for bits, sym in ((8, True), (4, True), (4, False)):
        mgr = QuantizationManager("UniformQuantizer", "MinMaxObserver", bits, sym, is_learning_scale=True)
        mgr.is_learning_scale = False
        mgr.is_observer_qparam = True
        mgr.is_quantize = False
        tag = f"b{bits}_{'sym' if sym else 'asym'}"
        xs = [torch.randn(3, 8, 9, 9, generator=gen) * (1 + i) for i in range(3)]
        for i, x in enumerate(xs):
            y = mgr.quantize(x)
            assert y is x
            out[f"{tag}_in{i}"] = x.numpy()
        out[f"{tag}_mean_abs"] = np.array(mgr.mean_abs_x)
        out[f"{tag}_mean"] = np.array(mgr.mean_x)
        out[f"{tag}_std"] = np.array(mgr.std)
        out[f"{tag}_minmax_scale_zp"] = np.array([mgr.observer.min_val, mgr.observer.max_val, mgr.scale, mgr.zero_point], dtype=np.float64)
        out[f"{tag}_observer_bits"] = np.array(mgr.observer.num_bits)
        # quantize with the calibrated (Python float) qparams
        mgr.is_quantize = True
        mgr.is_observer_qparam = False
        yq = mgr.quantize(xs[0])
        out[f"{tag}_yq_fixed"] = yq.numpy()
        mgr.init_scaling_factor_for_learning()
        out[f"{tag}_lsq_init"] = np.array(mgr.scale, dtype=np.float64)
        mgr.is_learning_scale = True
        mgr.make_learn_qparameter()
        out[f"{tag}_param_dtype"] = np.array(str(mgr.scale.dtype))
        out[f"{tag}_zp_after_learn"] = np.array(float(mgr.zero_point))
        if sym:
            xt = xs[1].clone().requires_grad_(True)
            y = mgr.quantize(xt)
            g = torch.randn(y.shape, generator=gen)
            y.backward(g)
            out[f"{tag}_learn_g"] = g.numpy()
            out[f"{tag}_learn_y"] = y.detach().numpy()
            out[f"{tag}_learn_dx"] = xt.grad.numpy()
            out[f"{tag}_learn_ds"] = mgr.scale.grad.numpy().reshape(1)
        tags.append(tag)
    out["cases"] = np.array(tags)
    save("manager", **out)


def gen_lsq_per_channel(gen):
    """LSQFakeQuantize per channel (lsq_module.py:147-173, 254-274, 317-358); qparams live on dim 1."""
    from torch.ao.quantization.observer import MovingAveragePerChannelMinMaxObserver
    out = {}
    tags = []
    for tag, shape, config_act, qmin, qmax, qscheme in (
        ("act_affine", (3, 6, 7, 5), True, 0, 255, torch.per_channel_affine),
        ("w_sym4", (2, 5, 3, 3), False, -8, 7, torch.per_channel_symmetric),
        ("w_affine4", (1, 8, 4, 4), False, 0, 15, torch.per_channel_affine),
    ):
        fq = LSQFakeQuantize(learn_scale=True, config_act=config_act,
                             observer=MovingAveragePerChannelMinMaxObserver,
                             quant_min=qmin, quant_max=qmax, dtype=torch.qint8 if qmin < 0 else torch.quint8,
                             qscheme=qscheme, ch_axis=1)
        x0 = torch.randn(*shape, generator=gen) * torch.linspace(0.5, 3.0, shape[1]).view(1, -1, 1, 1) + 0.3
        fq(x0)  # observer pass: initialises scale_param / zero_point_param_float
        fq.disable_observer()
        # move the learned parameters off their initial (integer zp) values
        with torch.no_grad():
            fq.zero_point_param_float.add_(torch.linspace(-0.7, 0.9, shape[1]).view(1, -1, 1, 1))
            fq.scale_param.mul_(torch.linspace(0.8, 1.3, shape[1]).view(1, -1, 1, 1))
        x = (torch.randn(*shape, generator=gen) * 2.0 + 0.3).requires_grad_(True)
        y = fq(x)
        g = torch.randn(shape, generator=gen)
        y.backward(g)
        out[f"{tag}_x"] = x.detach().numpy()
        out[f"{tag}_g"] = g.numpy()
        out[f"{tag}_y"] = y.detach().numpy()
        out[f"{tag}_dx"] = x.grad.numpy()
        out[f"{tag}_scale"] = fq.scale_param.detach().numpy().reshape(-1)
        out[f"{tag}_zpf"] = fq.zero_point_param_float.detach().numpy().reshape(-1)
        out[f"{tag}_ds"] = fq.scale_param.grad.numpy().reshape(-1).astype(np.float64)
        out[f"{tag}_dz"] = fq.zero_point_param_float.grad.numpy().reshape(-1).astype(np.float64)
        out[f"{tag}_qp"] = np.array([qmin, qmax, int(config_act)], dtype=np.float64)
        tags.append(tag)
    out["cases"] = np.array(tags)
    save("lsq_per_channel", **out)


def gen_bn_fold(gen):
    """ConvBnReLU / LinearBnReLU construction-time fold (fused.py:92-108, 286-300)."""
    out = {}
    tags = []
    args = ("MinMaxObserver", "UniformQuantizer", "MinMaxObserver", "UniformQuantizer", True, True, True, 8, 8)
    for tag, cin, cout, k, bias in (("c3x3_nobias", 5, 12, 3, False), ("c1x1_bias", 16, 7, 1, True), ("c3x3_first", 3, 16, 3, False)):
        cv = torch.nn.Conv2d(cin, cout, k, 1, k // 2, bias=bias)
        bn = torch.nn.BatchNorm2d(cout, eps=0.001)
        with torch.no_grad():
            cv.weight.copy_(torch.randn(cv.weight.shape, generator=gen))
            if bias:
                cv.bias.copy_(torch.randn(cout, generator=gen))
            bn.weight.copy_(torch.randn(cout, generator=gen))
            bn.bias.copy_(torch.randn(cout, generator=gen))
            bn.running_mean.copy_(torch.randn(cout, generator=gen))
            bn.running_var.copy_(torch.rand(cout, generator=gen) * 2 + 1e-3)
        f = ConvBnReLU(cv, bn, torch.nn.ReLU(), *args)
        out[f"{tag}_W"] = cv.weight.detach().numpy()
        out[f"{tag}_b"] = cv.bias.detach().numpy() if bias else np.zeros(0, np.float32)
        out[f"{tag}_bn"] = np.stack([bn.weight.detach().numpy(), bn.bias.detach().numpy(),
                                     bn.running_mean.numpy(), bn.running_var.numpy()])
        out[f"{tag}_eps"] = np.array(bn.eps)
        out[f"{tag}_Wf"] = f.conv_fuse.weight.detach().numpy()
        out[f"{tag}_bf"] = f.conv_fuse.bias.detach().numpy()
        tags.append(tag)
    lin = torch.nn.Linear(10, 6, bias=True)
    bn = torch.nn.BatchNorm1d(6)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(6, 10, generator=gen))
        lin.bias.copy_(torch.randn(6, generator=gen))
        bn.weight.copy_(torch.randn(6, generator=gen))
        bn.bias.copy_(torch.randn(6, generator=gen))
        bn.running_mean.copy_(torch.randn(6, generator=gen))
        bn.running_var.copy_(torch.rand(6, generator=gen) + 0.1)
    f = LinearBnReLU(lin, bn, torch.nn.ReLU(), *args)
    tag = "linear"
    out[f"{tag}_W"] = lin.weight.detach().numpy()
    out[f"{tag}_b"] = lin.bias.detach().numpy()
    out[f"{tag}_bn"] = np.stack([bn.weight.detach().numpy(), bn.bias.detach().numpy(), bn.running_mean.numpy(), bn.running_var.numpy()])
    out[f"{tag}_eps"] = np.array(bn.eps)
    out[f"{tag}_Wf"] = f.linear_fuse.weight.detach().numpy()
    out[f"{tag}_bf"] = f.linear_fuse.bias.detach().numpy()
    tags.append(tag)
    out["cases"] = np.array(tags)
    save("bn_fold", **out)


class _Loader:
    """Mimics the (uint8 images, targets) batches the reference drivers feed (yolov8_qat.py:46-47)."""

    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        return iter((b, None) for b in self.batches)


def _data_calib(model, loader, device):
    model.eval()
    for imgs, _ in loader:
        model(imgs.float() / 255.0)


def gen_bn_reestimate(gen):
    """reestimate_BN_stats on a one-layer model with is_fuse_bn=False (estimate_bn.py:38-101)."""
    cv = torch.nn.Conv2d(3, 6, 3, 1, 1, bias=False)
    bn = torch.nn.BatchNorm2d(6, eps=0.001, momentum=0.03)
    with torch.no_grad():
        cv.weight.copy_(torch.randn(cv.weight.shape, generator=gen) * 0.3)
    layer = ConvBnReLU(cv, bn, torch.nn.ReLU(), "MinMaxObserver", "UniformQuantizer", "MinMaxObserver",
                       "UniformQuantizer", True, True, False, 8, 8)
    layer.weight_quantizer.is_quantize = False
    layer.activation_quantizer.is_quantize = False
    model = torch.nn.Sequential(layer)
    batches = [torch.randint(0, 256, (4, 3, 12, 10), generator=gen, dtype=torch.uint8) for _ in range(5)]
    conv_outs = [torch.nn.functional.conv2d(b.float() / 255.0, layer.conv_fuse.weight.detach(), None, 1, 1).numpy()
                 for b in batches]
    reestimate_BN_stats(model, _Loader(batches), num_batches=4)
    save("bn_reestimate", conv_out=np.stack(conv_outs), num_batches=np.array(4),
         running_mean=layer.bn.running_mean.numpy(), running_var=layer.bn.running_var.numpy(),
         momentum_after=np.array(layer.bn.momentum), training_after=np.array(layer.bn.training))


def gen_tiny_e2e(gen):
    """fuse -> calibrate -> activate_learning_qparam -> activate_quantizer -> fwd+bwd on TinyNet."""
    model = make_tiny(0)
    cfg = create_fuse_config_manager(default_config=FuseConfig(bits_w=8, bits_a=8))
    model = fuse_modules_unified(model, [["conv", "bn", "relu"]], is_trace=False, config_manager=cfg)
    calib = [torch.randint(0, 256, (2, 3, 32, 32), generator=gen, dtype=torch.uint8) for _ in range(2)]
    out = {"calib": np.stack([c.numpy() for c in calib])}
    fused_names = [n for n, m in model.named_modules() if hasattr(m, "weight_quantizer")]
    out["fused_names"] = np.array(fused_names)
    for n, m in model.named_modules():
        if hasattr(m, "weight_quantizer"):
            out[f"fold_{n}_W"] = m.conv_fuse.weight.detach().numpy()
            out[f"fold_{n}_b"] = m.conv_fuse.bias.detach().numpy()
    calibrate_qat_model(model, _Loader(calib), _data_calib, "cpu")
    for n, m in model.named_modules():
        if hasattr(m, "weight_quantizer"):
            for kind in ("weight_quantizer", "activation_quantizer"):
                q = getattr(m, kind)
                out[f"calib_{n}_{kind}"] = np.array([q.observer.min_val, q.observer.max_val, q.scale, q.zero_point], dtype=np.float64)
                out[f"calib_{n}_{kind}_mean_abs"] = np.array(q.mean_abs_x)
    activate_learning_qparam(model, use_init=True)
    activate_quantizer(model)
    for n, m in model.named_modules():
        if hasattr(m, "weight_quantizer"):
            for kind in ("weight_quantizer", "activation_quantizer"):
                out[f"init_{n}_{kind}"] = np.array(getattr(m, kind).scale.detach().numpy(), dtype=np.float64)
    model.train()
    x = torch.rand(2, 3, 32, 32, generator=gen)
    y = model(x)
    loss = (y ** 2).mean()
    loss.backward()
    out["x"] = x.numpy()
    out["y"] = y.detach().numpy()
    out["loss"] = np.array(loss.item())
    for n, p in model.named_parameters():
        out[f"grad_{n}"] = p.grad.numpy().astype(np.float64) if p.grad is not None else np.zeros(0)
    out["param_names"] = np.array([n for n, _ in model.named_parameters()])
    save("tiny_e2e", **out)


def gen_yolov8n_e2e(gen):
    """The tiny_e2e sequence on the reference's YOLOv8n (nets/yolov8.py, 57 fused layers) at 64x64: fuse -> calibrate ->
    activate_learning_qparam -> activate_quantizer -> fwd+bwd.  Weights come from tests/golden/yolo_fill.py (seeded), so
    only a few hundred scalars are stored: per-layer digests of the folded weights, calibration extrema / scales,
    LSQ-initialised scales and the loss."""
    import hashlib
    from nets.yolov8 import yolo_v8_n
    from yolo_fill import fill_
    model = fill_(yolo_v8_n(20), 0)
    cfg = create_fuse_config_manager(default_config=FuseConfig(bits_w=8, bits_a=8))
    model = fuse_modules_unified(model, [["conv", "bn", "relu"]], is_trace=False, config_manager=cfg)
    layers = [(n, m) for n, m in model.named_modules() if hasattr(m, "weight_quantizer")]
    calib = [torch.randint(0, 256, (2, 3, 64, 64), generator=gen, dtype=torch.uint8) for _ in range(2)]
    out = {"calib": np.stack([c.numpy() for c in calib]), "fused_names": np.array([n for n, _ in layers])}
    out["fold_sha256"] = np.array([hashlib.sha256(m.conv_fuse.weight.detach().numpy().tobytes()
                                                  + m.conv_fuse.bias.detach().numpy().tobytes()).hexdigest()
                                   for _, m in layers])
    calibrate_qat_model(model, _Loader(calib), _data_calib, "cpu")
    for kind in ("weight_quantizer", "activation_quantizer"):
        out[f"calib_{kind}"] = np.array([[getattr(m, kind).observer.min_val, getattr(m, kind).observer.max_val,
                                          getattr(m, kind).scale, getattr(m, kind).zero_point] for _, m in layers],
                                        dtype=np.float64)
    activate_learning_qparam(model, use_init=True)
    activate_quantizer(model)
    for kind in ("weight_quantizer", "activation_quantizer"):
        out[f"init_{kind}"] = np.array([float(getattr(m, kind).scale.detach()) for _, m in layers], dtype=np.float64)
    model.train()
    x = torch.rand(2, 3, 64, 64, generator=gen)
    outs = model(x)
    loss = sum((o ** 2).mean() for o in outs)
    loss.backward()
    out["x"] = x.numpy()
    out["loss"] = np.array(loss.item())
    out["out_abs_mean"] = np.array([float(o.detach().abs().mean()) for o in outs])
    out["scale_grads"] = np.array([[float(getattr(m, k).scale.grad) for k in ("weight_quantizer", "activation_quantizer")]
                                   for _, m in layers], dtype=np.float64)
    save("yolov8n_e2e", **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1234)
    print("writing golden fixtures from the reference at", ref_shim.REFERENCE_ROOT)
    gen_uniform_fixed(gen)
    gen_uniform_learned(gen)
    gen_funlsq(gen)
    gen_observer(gen)
    gen_manager(gen)
    gen_lsq_per_channel(gen)
    gen_bn_fold(gen)
    gen_bn_reestimate(gen)
    gen_tiny_e2e(gen)
    gen_compute_scale(torch.Generator().manual_seed(4321))
    gen_yolov8n_e2e(torch.Generator().manual_seed(777))


def gen_compute_scale(gen):
    """compute_scale (utils/estimate_bn.py:104-139): per-channel calib_grad_scale from BN affine params and weight moments."""
    from utils.estimate_bn import compute_scale
    cv = torch.nn.Conv2d(4, 6, 3, 1, 1, bias=False)
    bn = torch.nn.BatchNorm2d(6, eps=0.001)
    with torch.no_grad():
        cv.weight.copy_(torch.randn(cv.weight.shape, generator=gen) * 0.3)
        bn.weight.copy_(torch.rand(6, generator=gen) + 0.5)
        bn.bias.copy_(torch.randn(6, generator=gen) * 0.2)
    layer = ConvBnReLU(cv, bn, torch.nn.ReLU(), "MinMaxObserver", "UniformQuantizer", "MinMaxObserver",
                       "UniformQuantizer", True, True, False, 8, 8)
    model = torch.nn.Sequential(layer)
    compute_scale(model, None)
    save("compute_scale", W=cv.weight.detach().numpy(), gamma=bn.weight.detach().numpy(), beta=bn.bias.detach().numpy(),
         calib_grad_scale=layer.activation_quantizer.quantizer.calib_grad_scale.numpy())


if __name__ == "__main__":
    if os.environ.get("VSIQ_GOLDEN_ONLY") == "compute_scale":
        gen_compute_scale(torch.Generator().manual_seed(4321))
    elif os.environ.get("VSIQ_GOLDEN_ONLY") == "yolov8n_e2e":
        gen_yolov8n_e2e(torch.Generator().manual_seed(777))
    else:
        main()
