"""CPU-only checks: the C-ABI library loads and exports every symbol include/vsiq.h declares (no compute calls without a
GPU), the ctypes binding covers them, and the host-side logic (config manager, pattern matching, registry, layouts)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vsiq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsiq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from vsiquantization_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/vsiq.h but not exported by libvsiq.so"
    assert sorted(_lib.EXPORTED) == syms, "ctypes binding and header disagree"
    assert _lib.lib.vsiq_version() == 100
    assert b"invalid" in _lib.lib.vsiq_error_string(-1)


def test_no_cpu_fallback():
    from vsiquantization_b200 import ops
    from vsiquantization_b200.quantizers.uniform import UniformQuantizer
    with pytest.raises(RuntimeError):
        ops.fake_quant_forward(torch.zeros(8), 0.1, 0, ops.QSpec(-8, 7))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            UniformQuantizer(8, True).quantize(torch.zeros(8), 0.1, 0, False)
        from vsiquantization_b200._lib import lib
        assert lib.vsiq_device_info(None, None, None) == -4  # VSIQ_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vsiquantization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "vsiq_oracle" not in src, f


def test_layout_of():
    from vsiquantization_b200.ops import layout_of
    assert layout_of((4, 3, 5, 5), None) == (1, 1, 300)
    assert layout_of((16, 3, 3, 3), 0) == (1, 16, 27)
    assert layout_of((2, 8, 20, 20), 1) == (2, 8, 400)
    assert layout_of((7,), None) == (1, 1, 7) and layout_of((4, 6), -1) == (4, 6, 1)


def test_registry_and_quantizer_ranges():
    from vsiquantization_b200.utils.registry import CLASS_REGISTRY, register_class
    import vsiquantization_b200.quantizers.quantization_manager  # noqa: F401
    assert {"UniformQuantizer", "LSQQuantizer", "MinMaxObserver", "LSQObserver"} <= set(CLASS_REGISTRY)
    for bits in range(2, 9):
        q = CLASS_REGISTRY["UniformQuantizer"](bits, True)
        assert (q.qmin, q.qmax) == (-(2 ** (bits - 1)), 2 ** (bits - 1) - 1)
        q = CLASS_REGISTRY["UniformQuantizer"](bits, False)
        assert (q.qmin, q.qmax) == (0, 2 ** bits - 1)

    @register_class
    class Probe:
        pass
    assert CLASS_REGISTRY["Probe"] is Probe
    del CLASS_REGISTRY["Probe"]
    with pytest.raises(KeyError):
        CLASS_REGISTRY["LSQQuantiser"]


def test_fuse_config_manager(tmp_path):
    from vsiquantization_b200.modules.fuse_config import (FuseConfig, FuseConfigManager, create_fuse_config_manager,
                                                            load_fuse_config_from_yaml)
    d = FuseConfig()
    assert (d.observer_w_name, d.quantizer_w_name, d.w_symmetric, d.a_symmetric, d.is_fuse_bn, d.bits_w, d.bits_a) == \
        ("MinMaxObserver", "UniformQuantizer", True, True, True, 8, 8)
    with pytest.raises(TypeError):
        FuseConfig(no_such_field=1)
    m = FuseConfigManager()
    m.add_layer_config("backbone.*conv", FuseConfig(bits_w=4))
    m.add_layer_config(".*conv.*", FuseConfig(bits_w=2))
    m.add_layer_config("head[", FuseConfig(bits_w=6))  # invalid regex -> substring match
    assert m.get_config_for_layer("backbone.3.conv").bits_w == 4      # first match wins, insertion order
    assert m.get_config_for_layer("neck.conv1").bits_w == 2
    assert m.get_config_for_layer("xhead[0").bits_w == 6
    assert m.get_config_for_layer("fc").bits_w == 8
    assert m.get_all_patterns() == ["backbone.*conv", ".*conv.*", "head["]
    m.clear_layer_configs()
    assert m.get_config_for_layer("backbone.3.conv").bits_w == 8
    with pytest.raises(ValueError):
        create_fuse_config_manager(layer_configs={"a": 3})
    assert create_fuse_config_manager(layer_configs={"a": {"bits_a": 4}}).get_config_for_layer("a").bits_a == 4
    with pytest.raises(FileNotFoundError):
        load_fuse_config_from_yaml(str(tmp_path / "missing.yaml"))
    bad = tmp_path / "bad.yaml"
    bad.write_text("default: [unclosed")
    with pytest.raises(ValueError):
        load_fuse_config_from_yaml(str(bad))
    ok = tmp_path / "ok.yaml"
    ok.write_text("default:\n  bits_w: 2\n  bits_a: 4\nlayers:\n  \"backbone.*conv\":\n    w_symmetric: false\n    bits_w: 4\n")
    mgr = load_fuse_config_from_yaml(str(ok))
    assert (mgr.default_config.bits_w, mgr.default_config.bits_a) == (2, 4)
    assert mgr.get_config_for_layer("backbone.0.conv").w_symmetric is False


def test_reference_sample_yaml_loads():
    ref = "/root/reference/configs/fuse_config.yaml"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present")
    from vsiquantization_b200.modules.fuse_config import load_fuse_config_from_yaml
    mgr = load_fuse_config_from_yaml(ref)
    assert (mgr.default_config.bits_w, mgr.default_config.bits_a) == (2, 4)  # configs/fuse_config.yaml:13-14
    assert len(mgr.get_all_patterns()) == 5


def test_pattern_matching_without_tensors():
    from tiny_model import make_tiny
    from vsiquantization_b200.modules.fuse import PATTERN_TO_FUSED, find_fusable_sequences, get_module_type_str
    model = make_tiny(0)
    hits = find_fusable_sequences(model, ["conv", "bn", "relu"])
    assert [(p, n) for p, _, n in hits] == [("stem", ["conv", "norm", "relu"]), ("backbone.0", ["conv", "norm", "relu"]),
                                            ("backbone.1", ["conv", "norm", "relu"])]  # SiLU counts as 'relu'
    assert ("", ["head"]) in [(p, n) for p, _, n in find_fusable_sequences(model, ["conv"])]
    assert find_fusable_sequences(model, ["linear", "bn"]) == []
    assert len(PATTERN_TO_FUSED) == 8
    assert get_module_type_str(torch.nn.SiLU()) is None and get_module_type_str(torch.nn.BatchNorm1d(3)) == "bn"


def test_dropin_maps_reference_module_paths():
    from vsiquantization_b200 import dropin
    reg = dropin.install_plugins(registry={})
    assert sorted(reg) == ["LSQObserver", "LSQQuantizer", "MinMaxObserver", "MovingAverageMinMaxObserver",
                           "MovingAveragePerChannelMinMaxObserver", "UniformQuantizer"]
    assert set(dropin._TIER2) >= {"modules.fuse", "modules.fuse_config", "utils.quantize_manager", "utils.estimate_bn",
                                  "quantizers.uniform", "observers.minmax"}


def test_dropin_install_modules_in_a_fresh_interpreter():
    """Tier 2: after install_modules() the reference's import lines resolve to this package."""
    import subprocess
    import sys
    code = ("import vsiquantization_b200.dropin as d; d.install_modules();"
            "from modules.fuse import fuse_modules_unified; from modules.fuse_config import load_fuse_config_from_yaml, FuseConfig;"
            "from utils.quantize_manager import calibrate_qat_model, activate_learning_qparam, activate_quantizer;"
            "from utils.estimate_bn import reestimate_BN_stats, compute_scale; from utils.registry import CLASS_REGISTRY;"
            "from quantizers.quantization_manager import QuantizationManager;"
            "from quantizers.lsq_module import LSQFakeQuantize;"
            "print(fuse_modules_unified.__module__, sorted(CLASS_REGISTRY))")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-500:]
    assert "vsiquantization_b200.modules.fuse" in out.stdout and "LSQQuantizer" in out.stdout


def test_checkpoint_keeps_calibration_and_learned_qparams():
    """state_dict round trip: the reference loses fixed scale / zero-point and observer extrema
    (plain attributes, yolov8_qat.py:299); here they travel in `_extra_state`, learned qparams keep the reference's
    key names, and a reference-style checkpoint (no extra state) still loads strictly."""
    import torch
    from vsiquantization_b200.quantizers.quantization_manager import QuantizationManager as M

    class Layer(torch.nn.Module):
        def __init__(self, sym=True):
            super().__init__()
            self.weight_quantizer = M("UniformQuantizer", "MinMaxObserver", 8, sym)
            self.activation_quantizer = M("LSQQuantizer", "LSQObserver", 4, sym)

    def calibrated(sym=True):
        layer = Layer(sym)
        for k, q in enumerate((layer.weight_quantizer, layer.activation_quantizer)):
            st = torch.zeros(1, 8, dtype=torch.float64)
            st[0, :5] = torch.tensor([-1.5 - k, 2.0 + k, (2.0 + k) / 127, 3.0 * (not sym), 2.0], dtype=torch.float64)
            st[0, 5] = 1.25
            q.observer.load_state(st)
            q._calibrated = True
            q._invalidate()
            q.is_learning_scale, q.is_quantize = False, True
        return layer

    a = calibrated()
    sd = a.state_dict()
    assert set(sd) == {"weight_quantizer._extra_state", "activation_quantizer._extra_state"}
    b = Layer()
    assert not b.load_state_dict(sd, strict=True).missing_keys
    for qa, qb in ((a.weight_quantizer, b.weight_quantizer), (a.activation_quantizer, b.activation_quantizer)):
        assert (qb.scale, qb.zero_point) == (qa.scale, qa.zero_point) and isinstance(qb.scale, float)
        assert (qb.observer.min_val, qb.observer.max_val) == (qa.observer.min_val, qa.observer.max_val)
        assert (qb.is_learning_scale, qb.is_quantize, qb.is_observer_qparam) == (False, True, True)
    # learned qparams: reference key names, re-created on a freshly fused model
    c = calibrated(sym=False)
    for q in (c.weight_quantizer, c.activation_quantizer):
        q.is_learning_scale = True
        q.init_scaling_factor_for_learning()
        q.make_learn_qparameter()
    with torch.no_grad():
        c.weight_quantizer.scale.fill_(0.0123)
    sd = c.state_dict()
    assert {"weight_quantizer.scale", "weight_quantizer.zero_point", "activation_quantizer.scale"} <= set(sd)
    assert sd["weight_quantizer.scale"].dtype == torch.float64 and sd["weight_quantizer.scale"].shape == ()
    d = Layer(sym=False)
    res = d.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert isinstance(d.weight_quantizer.scale, torch.nn.Parameter) and d.weight_quantizer.scale.item() == 0.0123
    assert torch.equal(d.weight_quantizer.zero_point.detach(), c.weight_quantizer.zero_point.detach())
    assert d.weight_quantizer.is_learning_scale and d.activation_quantizer.is_learning_scale
    # a checkpoint written by the reference: learned scales only, no extra state
    ref_sd = {k: v for k, v in sd.items() if not k.endswith("_extra_state")}
    e = Layer(sym=False)
    res = e.load_state_dict(ref_sd, strict=True)
    assert not res.missing_keys and e.weight_quantizer.scale.item() == 0.0123


def test_state_dict_holds_tensors_only_and_survives_an_ema_loop():
    """Training engines walk state_dict() blindly: ultralytics' ModelEMA (the reference's qat_yolov11_ultralytics.py
    driver builds one over the fused model) tests ``v.dtype.is_floating_point`` on every entry, and the reference's
    load_partial_checkpoint reads ``.shape`` (utils/util.py:27-29, 401-405).  The extra state is one packed tensor."""
    import copy
    import torch
    from vsiquantization_b200.quantizers.quantization_manager import QuantizationManager as M

    class Layer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 4, 3)
            self.weight_quantizer = M("UniformQuantizer", "MinMaxObserver", 8, True)
            self.activation_quantizer = M("LSQQuantizer", "LSQObserver", 4, False)

    a = Layer()
    st = torch.zeros(1, 8, dtype=torch.float64)
    st[0, :6] = torch.tensor([-1.5, 2.0, 2.0 / 127, 0.0, 2.0, 1.25], dtype=torch.float64)
    a.weight_quantizer.observer.load_state(st)
    a.weight_quantizer._calibrated = True
    a.weight_quantizer._invalidate()
    a.activation_quantizer.scale = 0.125          # host code set explicit values
    a.activation_quantizer.zero_point = 3
    a.activation_quantizer.quantizer.calib_grad_scale = torch.tensor([0.5, 2.0, 1.0])
    sd = a.state_dict()
    assert all(isinstance(v, torch.Tensor) for v in sd.values()), {k: type(v) for k, v in sd.items()}
    # ultralytics ModelEMA.update
    ema = copy.deepcopy(a).eval()
    d = 0.9
    msd = a.state_dict()
    for k, v in ema.state_dict().items():
        if v.dtype.is_floating_point:
            v *= d
            v += (1 - d) * msd[k].detach()
    # the reference's load_partial_checkpoint
    model_dict = Layer().state_dict()
    filtered = {k: v for k, v in sd.items() if k in model_dict and v.shape == model_dict[k].shape}
    assert "conv.weight" in filtered
    # round trip keeps the Python types the reference's code expects
    b = Layer()
    assert not b.load_state_dict(sd, strict=True).missing_keys
    assert isinstance(b.weight_quantizer.scale, float) and b.weight_quantizer.scale == a.weight_quantizer.scale
    assert b.activation_quantizer.scale == 0.125 and isinstance(b.activation_quantizer.zero_point, int)
    assert b.activation_quantizer.zero_point == 3
    assert torch.equal(b.activation_quantizer.quantizer.calib_grad_scale, torch.tensor([0.5, 2.0, 1.0]))
    # a checkpoint written with the former dict layout still loads
    legacy = dict(sd)
    legacy["weight_quantizer._extra_state"] = {"version": 1, "flags": {"is_quantize": False}, "calibrated": True,
                                                "observer_state": st, "scale": None, "zero_point": None,
                                                "calib_grad_scale": 1}
    c = Layer()
    c.load_state_dict(legacy, strict=True)
    assert c.weight_quantizer.is_quantize is False and c.weight_quantizer.scale == a.weight_quantizer.scale


def test_multi_tensor_plan_geometry_on_the_host():
    """vsiq_mt_plan is pure host code: tile prefix, per-tensor vs per-channel tiling, argument validation."""
    from vsiquantization_b200 import _lib
    shapes = [(16, 27, 1), (64, 288, 64), (1, 5, 1), (256, 2304, 256), (0, 10, 1)]  # (rows, inner, qp_channels)
    tab = (_lib.MtEntry * len(shapes))()
    out = q = 0
    for e, (rows, inner, qpc) in zip(tab, shapes):
        e.x = 0x1000
        e.rows, e.inner, e.out_offset, e.qp_offset, e.qp_channels = rows, inner, out, q, (qpc if rows else 1)
        e.qp.qmin, e.qp.qmax, e.qp.scale_host = -8, 7, 0.1
        out += (rows * inner + 7) // 8 * 8
        q += e.qp_channels
    total = ctypes.c_uint32(0)
    assert _lib.lib.vsiq_mt_plan(tab, len(shapes), ctypes.byref(total)) == 0
    first = 0
    for e, (rows, inner, qpc) in zip(tab, shapes):
        assert e.first_tile == first
        if rows == 0:
            assert e.n_tiles == 0
        elif qpc > 1:   # per channel: every row is cut into `chunks` tiles of <= 1024 elements
            assert e.n_tiles == rows * e.chunks and e.chunks == -(-inner // e.tile) and 0 < e.tile <= 1024
        else:           # per tensor: the whole tensor is one row
            assert e.chunks == e.n_tiles == -(-(rows * inner) // e.tile) and 0 < e.tile <= 1024
        assert (e.tlo, e.thi) == (-8.5, 7.5 - 2 ** -21)  # qmin even: tie rounds to it; qmax odd: next float below 7.5
        first += e.n_tiles
    assert total.value == first and _lib.lib.vsiq_mt_workspace_bytes(first) == 256 + 16 * first
    tab[1].out_offset += 4                                   # not a multiple of 8 elements
    assert _lib.lib.vsiq_mt_plan(tab, len(shapes), ctypes.byref(total)) == -1
    tab[1].out_offset -= 4
    tab[1].qp_channels = 3                                   # neither 1 nor rows
    assert _lib.lib.vsiq_mt_plan(tab, len(shapes), ctypes.byref(total)) == -1
    tab[1].qp_channels = 64
    tab[0].qp.pre_op = 1                                     # fused ReLU is an activation feature
    assert _lib.lib.vsiq_mt_plan(tab, len(shapes), ctypes.byref(total)) == -3


@pytest.mark.parametrize("symmetric,bits", [(True, 8), (False, 8), (True, 4), (False, 4)])
@pytest.mark.parametrize("per_channel", [False, True])
def test_moving_average_observer_math_matches_torch_observers(symmetric, bits, per_channel):
    """observers/moving_average.py::ema_update_ (the [C]-sized running update + qparams that follow the CUDA observer
    pass) against torch.ao's MovingAverage(PerChannel)MinMaxObserver -- the observer LSQFakeQuantize uses
    (quantizers/lsq_module.py:1-2,91,113-115) -- bit for bit over several batches, NaN-free data, both schemes."""
    from torch.ao.quantization.observer import MovingAverageMinMaxObserver, MovingAveragePerChannelMinMaxObserver
    from vsiquantization_b200.observers.moving_average import ema_update_
    qmin, qmax = (-(2 ** (bits - 1)), 2 ** (bits - 1) - 1) if symmetric else (0, 2 ** bits - 1)
    kw = dict(averaging_constant=0.01, quant_min=qmin, quant_max=qmax, dtype=torch.qint8 if symmetric else torch.quint8)
    if per_channel:
        ref = MovingAveragePerChannelMinMaxObserver(
            ch_axis=1, qscheme=torch.per_channel_symmetric if symmetric else torch.per_channel_affine, **kw)
    else:
        ref = MovingAverageMinMaxObserver(qscheme=torch.per_tensor_symmetric if symmetric else torch.per_tensor_affine, **kw)
    C = 6 if per_channel else 1
    state = torch.zeros(C, 8, dtype=torch.float64)
    state[:, 2] = 1.0
    g = torch.Generator().manual_seed(3)
    for i in range(6):
        x = torch.randn(3, 6, 5, 4, generator=g) * (1 + i) + 0.3 * i
        if i == 2:
            x = x.abs() + 0.5          # all-positive batch: min(m, 0) matters
        ref(x)
        s_ref, z_ref = ref.calculate_qparams()
        if per_channel:
            bmin, bmax = torch.aminmax(x.permute(1, 0, 2, 3).flatten(1), dim=1)
        else:
            bmin, bmax = (t.reshape(1) for t in torch.aminmax(x))
        ema_update_(state, bmin, bmax, 0.01, qmin, qmax, symmetric)
        assert torch.equal(state[:, 0].float(), ref.min_val.reshape(-1)) and torch.equal(state[:, 1].float(), ref.max_val.reshape(-1))
        assert torch.equal(state[:, 2].float(), s_ref.reshape(-1).float()), (i, state[:, 2], s_ref)
        assert torch.equal(state[:, 3].long(), z_ref.reshape(-1).long())
        assert float(state[0, 4]) == i + 1


def test_tier1_plugins_inside_the_unmodified_reference():
    """Tier 1 of INTEGRATION.md against the REAL reference tree (build container only; skipped where /root/reference
    is absent): dropin.install_plugins() registers the native classes in the reference's own CLASS_REGISTRY, and the
    reference's QuantizationManager / ConvBnReLU then build them by name with its positional conventions
    (quantization_manager.py:41-42).  Construction only -- no kernel runs on this CPU-only host."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import ref_shim; ref_shim.install()\n"
        "import torch\n"
        "from utils.registry import CLASS_REGISTRY\n"
        "import quantizers.quantization_manager as rqm\n"
        "assert rqm.__file__.startswith(ref_shim.REFERENCE_ROOT)\n"
        "before = sorted(CLASS_REGISTRY)\n"
        "import vsiquantization_b200.dropin as d; d.install_plugins()\n"
        "m = rqm.QuantizationManager('LSQQuantizer', 'LSQObserver', 4, True)\n"
        "assert type(m.quantizer).__module__ == 'vsiquantization_b200.quantizers.uniform', type(m.quantizer)\n"
        "assert type(m.observer).__module__ == 'vsiquantization_b200.observers.minmax'\n"
        "assert (m.quantizer.qmin, m.quantizer.qmax, m.quantizer.calib_grad_scale, m.observer.num_bits) == (-8, 7, 1, 8)\n"
        "import modules.fused as rf\n"
        "assert rf.__file__.startswith(ref_shim.REFERENCE_ROOT)\n"
        "layer = rf.ConvBnReLU(torch.nn.Conv2d(3, 4, 3, bias=False), torch.nn.BatchNorm2d(4), torch.nn.ReLU(),\n"
        "                      'MinMaxObserver', 'UniformQuantizer', 'MinMaxObserver', 'UniformQuantizer', True, True, True, 8, 8)\n"
        "assert type(layer.weight_quantizer.quantizer).__module__.startswith('vsiquantization_b200')\n"
        "assert type(layer.activation_quantizer.observer).__module__.startswith('vsiquantization_b200')\n"
        "print('before', before, 'after', sorted(CLASS_REGISTRY))\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stderr[-800:]
    assert "before ['MinMaxObserver', 'UniformQuantizer'] after" in out.stdout and "LSQQuantizer" in out.stdout


def test_yolov8_fixture_is_the_reference_network():
    """vsiquantization_b200/nets/yolov8.py (the model behind every YOLOv8 number in DESIGN.md) against the reference's
    nets/yolov8.py (build container only): same parameter / buffer names, shapes and order for n, s, m, l, and the same
    float forward, bit for bit, with the same weights (YOLOv8n, 64x64 input, training-mode heads)."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import ref_shim; ref_shim.install()\n"
        "import torch\n"
        "import nets.yolov8 as ry\n"
        "from vsiquantization_b200.nets import yolov8 as oy\n"
        "assert ry.__file__.startswith(ref_shim.REFERENCE_ROOT)\n"
        "for name in 'nsml':\n"
        "    a = getattr(ry, 'yolo_v8_' + name)(20); b = getattr(oy, 'yolo_v8_' + name)(num_classes=20)\n"
        "    assert [(n, tuple(p.shape)) for n, p in a.named_parameters()] == [(n, tuple(p.shape)) for n, p in b.named_parameters()], name\n"
        "    assert [(n, tuple(p.shape)) for n, p in a.named_buffers()] == [(n, tuple(p.shape)) for n, p in b.named_buffers()], name\n"
        "torch.manual_seed(0)\n"
        "a = ry.yolo_v8_n(20); b = oy.yolo_v8_n(num_classes=20)\n"
        "b.load_state_dict(a.state_dict(), strict=True)\n"
        "x = torch.rand(2, 3, 64, 64)\n"
        "a.train(); b.train()\n"
        "ya, yb = a(x), b(x)\n"
        "assert len(ya) == len(yb) and all(torch.equal(u, v) for u, v in zip(ya, yb))\n"
        "print('same network', len(ya))\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-800:]
    assert "same network" in out.stdout


def test_lsq_fake_quantize_module_against_the_live_reference():
    """quantizers/lsq_module.py::LSQFakeQuantize (the reference's torch FakeQuantize subclass, lsq_module.py:73-173):
    same constructor, buffers, parameters, state_dict keys, phases, outputs and gradients as the live reference in
    eight configurations (tests/live_reference_facade_check.py; kernels replaced by oracle-backed stand-ins on this
    GPU-less host).  Skipped where /root/reference is absent."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "live_reference_facade_check.py")], cwd="/tmp",
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    assert "facade matches the live reference in 8 configurations" in out.stdout


def test_lsq_fake_quantize_module_has_no_cpu_fallback():
    from vsiquantization_b200.quantizers.lsq_module import LSQFakeQuantize
    fq = LSQFakeQuantize(learn_scale=True, quant_min=-128, quant_max=127, dtype=torch.qint8,
                         qscheme=torch.per_tensor_symmetric)
    assert set(fq.state_dict()) >= {"scale", "zero_point", "observer_enabled", "fake_quant_enabled",
                                    "activation_post_process.min_val", "activation_post_process.max_val"}
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            fq(torch.zeros(2, 3))
    fq.disable_observer()
    assert fq._obs_on is False and int(fq.observer_enabled[0]) == 0
