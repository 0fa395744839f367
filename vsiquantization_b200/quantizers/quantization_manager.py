"""QuantizationManager: per-tensor state machine between a fused layer and its observer / quantizer plugins.

Reference: quantizers/quantization_manager.py:10-114.  Same constructor, attributes (quantizer, observer, bits_width,
scale, zero_point, is_observer_qparam, is_learning_scale, is_quantize, is_symmetric, mean_abs_x, mean_x, std) and
methods (collect_qparameter, quantize, make_learn_qparameter, init_scaling_factor_for_learning); plugins are built
from the registry by name exactly like :41-42 (so the observer is always 8-bit unless ``observer_bits`` is given --
SURVEY.md 0.4).  What differs is WHERE the numbers live:

  * calibration (collect_qparameter) is one kernel launch and NO host synchronisation per call: min/max, the three
    statistics, the running state, scale and zero-point all stay on the device (the reference does 5 reductions and
    5 ``.item()`` syncs, :66-69).  ``scale`` / ``zero_point`` / ``mean_abs_x`` / ``mean_x`` / ``std`` are read back
    lazily, only when host code asks for them;
  * with fixed qparams the kernels read scale / zero-point straight from the observer state on the device;
  * init_scaling_factor_for_learning + make_learn_qparameter build the learnable Parameter on the device.

Checkpoints: the reference's calibrated scale / zero-point are plain Python attributes and its
observer extrema live outside any nn.Module, so ``state_dict()`` loses them and only the learned ``scale`` Parameters
survive (yolov8_qat.py:299,309).  Here the same keys are kept (``...weight_quantizer.scale`` / ``.zero_point`` for
learned qparams) and everything else -- observer state, fixed qparams, the mode flags -- travels in the module's
``_extra_state`` entry; loading a checkpoint that has learned qparams into a freshly fused model re-creates the
Parameters, and a reference checkpoint without ``_extra_state`` still loads.

Deliberate fixes where the reference cannot run: the constructor's ``is_symmetric`` is stored
(the reference hard-codes True, :50, so an asymmetric learnable zero-point is unreachable and crashes, uniform.py:50-52);
asymmetric managers therefore get the learnable zero-point the code at :100-101 intends.  Symmetric flows -- every
flow the reference can execute -- are unchanged.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from ..observers.minmax import MinMaxObserver  # noqa: F401  (registers the plugin classes)
from ..observers.moving_average import MovingAverageMinMaxObserver  # noqa: F401
from ..quantizers.uniform import UniformQuantizer  # noqa: F401
from ..utils.registry import CLASS_REGISTRY

_LAZY = ("scale", "zero_point")


class QuantizationManager(nn.Module):
    def __init__(self, quantizer_name: str, observer_name: str, bits_width: int, is_symmetric: bool,
                 is_learning_scale: bool = True, observer_bits: Optional[int] = None, ch_axis: Optional[int] = None,
                 param_dtype: torch.dtype = torch.float64) -> None:
        super().__init__()
        self.quantizer = CLASS_REGISTRY[quantizer_name](bits_width, is_symmetric)
        if observer_bits is None:
            self.observer = CLASS_REGISTRY[observer_name](is_symmetric)  # num_bits stays 8, like the reference
        else:
            self.observer = CLASS_REGISTRY[observer_name](is_symmetric, observer_bits)
        if ch_axis is not None:
            self.quantizer.ch_axis = ch_axis
            self.observer.ch_axis = ch_axis
        self.ch_axis = ch_axis
        self.bits_width = bits_width
        self.param_dtype = param_dtype  # the reference's learned scale is a 0-dim float64 Parameter
        self.scale = 1
        self.zero_point = 0
        self.is_observer_qparam = True
        self.is_learning_scale = is_learning_scale
        self.is_quantize = True
        self.is_symmetric = is_symmetric
        self._call_stats: List = []   # (device stats [C,5], elements per channel) per calibration call
        self._calibrated = False      # the observer state on the device is the source of scale / zero_point

    # ---- lazily materialised host views ---------------------------------------------------------------
    def __getattr__(self, name):
        # reached only when normal lookup fails: scale / zero_point were invalidated by a calibration call
        if name in _LAZY and "_calibrated" in self.__dict__ and self.__dict__["_calibrated"]:
            s, z = self.observer.get_scale_zero_point()  # one D2H copy
            self.__dict__["scale"], self.__dict__["zero_point"] = s, z
            return self.__dict__[name]
        return super().__getattr__(name)

    def _invalidate(self) -> None:
        self.__dict__.pop("scale", None)
        self.__dict__.pop("zero_point", None)

    def _stats_host(self):
        if not self._call_stats:
            return []
        st = torch.stack([s[0] for s, _ in self._call_stats]).cpu()  # per-tensor view: channel 0
        return [(st[i], n) for i, (_, n) in enumerate(self._call_stats)]

    @property
    def mean_abs_x(self):
        """Per-call mean|x| (reference: a list of floats, :66)."""
        return [float(s[2]) / n for s, n in self._stats_host()]

    @property
    def mean_x(self):
        return [float(s[3]) / n for s, n in self._stats_host()]

    @property
    def std(self):
        out = []
        for s, n in self._stats_host():
            mean = float(s[3]) / n
            var = (float(s[4]) - n * mean * mean) / (n - 1) if n > 1 else float("nan")
            out.append(max(var, 0.0) ** 0.5)
        return out

    # ---- checkpointing --------------------------------------------------------------------------------------
    def get_extra_state(self):
        """Everything state_dict() would otherwise lose: mode flags, the observer's running state, fixed qparams."""
        def host(v):
            if isinstance(v, nn.Parameter):
                return None  # learned qparams are ordinary state_dict entries
            if isinstance(v, torch.Tensor):
                return v.detach().cpu()
            return v
        st = getattr(self.observer, "state", None)
        return {"version": 1,
                "flags": {"is_observer_qparam": bool(self.is_observer_qparam), "is_learning_scale": bool(self.is_learning_scale),
                          "is_quantize": bool(self.is_quantize)},
                "calibrated": bool(self._calibrated),
                "observer_state": st.detach().cpu() if isinstance(st, torch.Tensor) else None,
                "scale": host(self.__dict__.get("scale")), "zero_point": host(self.__dict__.get("zero_point")),
                "calib_grad_scale": host(getattr(self.quantizer, "calib_grad_scale", 1))}

    def set_extra_state(self, state) -> None:
        if not state:
            return
        for k, v in state.get("flags", {}).items():
            setattr(self, k, v)
        st = state.get("observer_state")
        if st is not None and hasattr(self.observer, "load_state"):
            self.observer.load_state(st)
        self._calibrated = bool(state.get("calibrated", False)) and st is not None
        self._invalidate()
        for name in _LAZY:
            v = state.get(name)
            if v is not None and name not in self._parameters:
                if isinstance(v, torch.Tensor) and st is not None:
                    v = v.to(self.observer.state.device)
                self.__dict__[name] = v
        if self._calibrated and (state.get("scale") is not None):
            self._calibrated = "scale" in self._parameters  # explicit host values win over the observer state
        cgs = state.get("calib_grad_scale")
        if cgs is not None and hasattr(self.quantizer, "calib_grad_scale"):
            self.quantizer.calib_grad_scale = cgs

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # a checkpoint with learned qparams loaded into a freshly fused model: create the Parameters first
        for name in _LAZY:
            t = state_dict.get(prefix + name)
            if isinstance(t, torch.Tensor) and name not in self._parameters:
                self.__dict__.pop(name, None)
                st = getattr(self.observer, "state", None)
                dev = st.device if isinstance(st, torch.Tensor) else t.device
                self.register_parameter(name, nn.Parameter(torch.empty_like(t, device=dev), requires_grad=True))
                self._calibrated = False
        n_missing = len(missing_keys)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        extra = prefix + "_extra_state"
        if extra in missing_keys[n_missing:]:  # checkpoints written by the reference have no extra state
            missing_keys.remove(extra)

    # ---- reference interface -----------------------------------------------------------------------------
    def collect_qparameter(self, x):
        """Observe x while calibrating (:55-71): one launch, no synchronisation."""
        if not self.is_learning_scale and self.is_observer_qparam:
            self.observer.observe(x)
            self._call_stats.append((self.observer.last_stats, self.observer.last_count))
            self._calibrated = True
            self._invalidate()

    def can_fuse_relu(self) -> bool:
        """True when quantize(x, pre_relu=True) will run relu + fake-quant as ONE kernel pass."""
        collecting = (not self.is_learning_scale) and self.is_observer_qparam
        return bool(self.is_quantize and not collecting and getattr(self.quantizer, "supports_pre_relu", False))

    def quantize(self, x, pre_relu: bool = False, bias=None):
        """collect (if calibrating) then fake-quantise (if enabled) -- :73-90.  ``pre_relu``: x is the pre-activation
        and relu is applied here -- fused into the quantiser kernels when quantising, as a plain F.relu otherwise.
        ``bias`` (with pre_relu, channels_last x): the conv bias is added in the same pass and its gradient comes out of
        the backward kernel."""
        banked = self.__dict__.get("_banked")
        if banked is not None:  # bank.WeightBank already fake-quantised this weight in its multi-tensor launch
            self.__dict__["_banked"] = None
            if banked[0] is x and not pre_relu:
                return banked[1]
        if pre_relu:
            if not self.can_fuse_relu():
                if bias is not None:
                    x = x + bias.view(1, -1, 1, 1)
                return self.quantize(torch.nn.functional.relu(x))
            kw = {"pre_relu": True}
            if bias is not None:
                kw["bias"] = bias
            if "scale" in self._parameters or "zero_point" in self._parameters or not self._calibrated \
                    or "scale" in self.__dict__:
                return self.quantizer.quantize(x, self.scale, self.zero_point, self.is_learning_scale, **kw)
            s, z = self.observer.device_qparams()
            return self.quantizer.quantize(x, s, z, self.is_learning_scale, **kw)
        self.collect_qparameter(x)
        if not self.is_quantize:
            return x
        if "scale" in self._parameters or "zero_point" in self._parameters or not self._calibrated \
                or "scale" in self.__dict__:
            # learned Parameters, never-calibrated defaults, or values host code has set / read back
            return self.quantizer.quantize(x, self.scale, self.zero_point, self.is_learning_scale)
        # calibrated, fixed qparams: feed the kernels from the observer state on the device (no sync)
        s, z = self.observer.device_qparams()
        return self.quantizer.quantize(x, s, z, self.is_learning_scale)

    def init_scaling_factor_for_learning(self):
        """scale = 2*mean(mean|x|)/sqrt(2^(b-1)-1) (:105-114), computed on the device; becomes a 0-dim (per tensor)
        or [C] tensor of ``param_dtype``.  Without calibration data the current scale is kept."""
        st = self.observer.state
        if st is None or not self._call_stats:
            return
        out = torch.empty(st.shape[0], dtype=self.param_dtype, device=st.device)
        self.observer.lsq_init_scale(self.bits_width, out)
        self._calibrated = False
        self.__dict__.pop("zero_point", None)
        self.__dict__["scale"] = out.reshape(()) if out.numel() == 1 else out
        self.__dict__["zero_point"] = self._current_zero_point()

    def _current_zero_point(self):
        st = self.observer.state
        if st is None:
            return 0
        z = st[:, 3]
        return z.reshape(()) if z.numel() == 1 else z.clone()

    def make_learn_qparameter(self):
        """scale -> nn.Parameter; asymmetric: zero_point -> float nn.Parameter initialised at zp + 1e-9 (:92-103);
        symmetric: zero_point = 0."""
        scale = self.scale
        zp_now = self.zero_point
        self._calibrated = False  # from here on the Parameters (not the observer state) are the source of truth
        if isinstance(scale, nn.Parameter):
            pass
        elif isinstance(scale, torch.Tensor):
            self.__dict__.pop("scale", None)
            self.scale = nn.Parameter(scale.detach().clone().to(self.param_dtype), requires_grad=True)
        else:
            dev = self.observer.state.device if self.observer.state is not None else None
            self.__dict__.pop("scale", None)
            # torch.tensor(python float) is float32, torch.tensor(np.float64) is float64 -- the reference gets fp64
            # after init_scaling_factor_for_learning and fp32 otherwise (:99); param_dtype decides here.
            self.scale = nn.Parameter(torch.tensor(float(scale), dtype=self.param_dtype, device=dev), requires_grad=True)
        if not self.is_symmetric:
            zp = zp_now
            if not isinstance(zp, nn.Parameter):
                self.__dict__.pop("zero_point", None)
                if isinstance(zp, torch.Tensor):
                    z0 = zp.detach().to(torch.float32) + 1e-9
                else:
                    z0 = torch.tensor(float(zp) + 1e-9, dtype=torch.float32, device=self.scale.device)
                self.zero_point = nn.Parameter(z0.to(self.scale.device).reshape(self.scale.shape), requires_grad=True)
        else:
            self.__dict__.pop("zero_point", None)
            if "zero_point" not in self._parameters:
                self.zero_point = 0
        self._calibrated = False
