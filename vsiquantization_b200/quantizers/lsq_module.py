"""LSQFakeQuantize: the reference's learnable per-tensor / per-channel fake-quantise module on the CUDA kernels.

Reference: quantizers/lsq_module.py:73-173 -- a ``torch.quantization.FakeQuantize`` subclass that is NOT wired into the
registry / FuseConfig path but holds the only per-channel code of the reference.  Same constructor
``(learn_scale=False, config_act=False, observer=MovingAverageMinMaxObserver, quant_min=None, quant_max=None,
**observer_kwargs)``, same buffers (``scale``, ``zero_point``, ``observer_enabled``, ``fake_quant_enabled``, the torch
observer under ``activation_post_process`` with its ``min_val`` / ``max_val``), same parameters (``scale_param``,
``zero_point_param_float`` shaped ``[1, C, 1, ...]`` per channel, the dormant ``theta`` / ``gamma`` of the adaptive-
rounding experiment), hence the same ``state_dict`` keys, and the same two phases:

  observer phase (``observer_enabled``; lsq_module.py:113-144)
      running extrema <- moving average of the batch extrema, scale / zero-point <- ``calculate_qparams`` arithmetic,
      learnable parameters (re)initialised from them.  Here the pass over X is ONE read by the observer kernel (torch's
      observer permutes, flattens -- a copy -- and runs ``aminmax``); the [C]-sized update is observers/moving_average.py
      (bit-identical to torch's observer on CPU).
  fake-quant phase (``fake_quant_enabled``; lsq_module.py:146-173)
      learn_scale and observer off: y = s_c * (clamp(round_ste(x / s_c + z_c)) - z_c) with z_c = clamp(round_ste(zf_c)),
      both qparams behind ScaleGradient with g = (quant_max * numel [/ C]) ** -0.5 (x 5000 for activations,
      lsq_module.py:151-152, 317-340) -- one forward kernel, one backward kernel (dx + dscale[C] + dzp[C]);
      otherwise: the buffers ``scale`` / ``zero_point`` as constants, straight-through gradient to x.

Not reproduced: ``flag_adaptive`` rounding (theta / gamma, lsq_module.py:282-300), power-of-two quantisation, the
``scale_grad_*`` experiments -- dead code in the reference (never enabled by any call site)."""
from __future__ import annotations

import torch
from torch.ao.quantization import FakeQuantize, MovingAverageMinMaxObserver

from .. import ops
from ..observers.moving_average import torch_qparams
from .uniform import LSQQuantizer, _as_cuda

_SYMMETRIC = (torch.per_tensor_symmetric, torch.per_channel_symmetric)
_AFFINE = (torch.per_tensor_affine, torch.per_channel_affine)


class LSQFakeQuantize(FakeQuantize):
    def __init__(self, learn_scale=False, config_act=False, observer=MovingAverageMinMaxObserver, quant_min=None,
                 quant_max=None, **observer_kwargs):
        super().__init__(observer, quant_min, quant_max, **observer_kwargs)
        self.learn_scale = learn_scale
        self.flag_param_quant = False
        self.flag_adaptive = False
        self.config_act = config_act
        if self.qscheme not in _SYMMETRIC + _AFFINE:
            raise NotImplementedError(f"LSQFakeQuantize supports symmetric and affine integer schemes, got {self.qscheme}")
        if self.qscheme in _SYMMETRIC and self.dtype not in (torch.qint8, torch.int8):
            raise NotImplementedError("symmetric schemes are supported on the signed range (dtype=torch.qint8)")
        # the plugin that owns the kernels' view of this quantiser (range, channel axis, activation boost)
        self._quantizer = LSQQuantizer(8, True, ch_axis=(1 if self.is_per_channel else None),
                                       grad_boost=5000.0 if config_act else 1.0)
        self._quantizer.qmin, self._quantizer.qmax = int(self.quant_min), int(self.quant_max)
        self._quantizer.symmetric = False  # a tensor zero-point is always rounded + clamped here (lsq_module.py:354-358)
        self._eps = float(self.activation_post_process.eps)  # read once on the host: no per-call synchronisation
        # host mirrors of the two enable buffers: the reference tests ``buffer[0] == 1`` on every forward, which is a
        # device -> host synchronisation once the module lives on the GPU (and cannot be captured in a CUDA graph)
        self._obs_on, self._fq_on = True, True

    def enable_observer(self, enabled: bool = True) -> None:
        super().enable_observer(enabled)
        self._obs_on = bool(enabled)

    def enable_fake_quant(self, enabled: bool = True) -> None:
        super().enable_fake_quant(enabled)
        self._fq_on = bool(enabled)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._obs_on = bool(int(self.observer_enabled[0]))   # one read per load, not per forward
        self._fq_on = bool(int(self.fake_quant_enabled[0]))

    # ---- observer phase ---------------------------------------------------------------------------------
    def _batch_extrema(self, X: torch.Tensor):
        """[C] fp32 batch minima / maxima from one pass of the observer kernel (per tensor: C = 1)."""
        stats = ops.observe(_as_cuda(X.detach()), 1 if self.is_per_channel else None, want_stats=True)
        return stats[:, 0].to(torch.float32), stats[:, 1].to(torch.float32)

    def _observe(self, X: torch.Tensor) -> None:
        app = self.activation_post_process
        bmin, bmax = self._batch_extrema(X)
        bmin, bmax = bmin.to(app.min_val.device), bmax.to(app.max_val.device)
        c = app.averaging_constant
        if self.is_per_channel:
            if app.min_val.numel() == 0 or app.max_val.numel() == 0:        # shape test: no synchronisation
                new_min, new_max = bmin, bmax
            else:
                new_min = app.min_val + c * (bmin - app.min_val)
                new_max = app.max_val + c * (bmax - app.max_val)
            app.min_val.resize_(new_min.shape)
            app.max_val.resize_(new_max.shape)
        else:
            lo, hi = app.min_val.reshape(1), app.max_val.reshape(1)
            fresh = torch.isinf(lo) & (lo > 0) & torch.isinf(hi) & (hi < 0)  # torch's "== inf and == -inf" first call
            new_min = torch.where(fresh, bmin, lo + c * (bmin - lo)).reshape(app.min_val.shape)
            new_max = torch.where(fresh, bmax, hi + c * (bmax - hi)).reshape(app.max_val.shape)
        app.min_val.copy_(new_min)
        app.max_val.copy_(new_max)
        _scale, _zero_point = torch_qparams(app.min_val.reshape(-1), app.max_val.reshape(-1), int(self.quant_min),
                                            int(self.quant_max), self.qscheme in _SYMMETRIC, self._eps)
        _scale, _zero_point = _scale.to(self.scale.device), _zero_point.to(self.zero_point.device)
        if self.scale.shape != _scale.shape:
            self.scale.resize_(_scale.shape)
            self.zero_point.resize_(_zero_point.shape)
        self.scale.copy_(_scale)
        self.zero_point.copy_(_zero_point)
        if self.learn_scale:
            scale_init, zero_point_init = _scale, _zero_point.float()
            if self.is_per_channel:
                view = [1, -1] + [1] * (X.dim() - 2)
                scale_init, zero_point_init = scale_init.view(view), zero_point_init.view(view)
            if not self.flag_param_quant:
                self.register_parameter("scale_param", torch.nn.Parameter(scale_init.clone()))
                self.register_parameter("zero_point_param_float", torch.nn.Parameter(zero_point_init.clone()))
                if not self.config_act:  # dormant adaptive-rounding parameters, kept for state_dict compatibility
                    self.register_parameter("theta", torch.nn.Parameter(torch.ones_like(X), requires_grad=False))
                    self.register_parameter("gamma", torch.nn.Parameter(torch.zeros_like(X), requires_grad=False))
                self.flag_param_quant = True
            else:
                self.scale_param.data.copy_(scale_init)
                self.zero_point_param_float.data.copy_(zero_point_init)

    # ---- forward ------------------------------------------------------------------------------------------
    def forward(self, X):
        if self._obs_on:
            self._observe(X)
        if self._fq_on:
            if self.flag_adaptive:
                raise NotImplementedError("adaptive rounding (flag_adaptive) is dead code in the reference and not reproduced")
            if self.learn_scale and not self._obs_on:
                X = self._quantizer.quantize(X, self.scale_param, self.zero_point_param_float, True)
            else:
                scale, zero_point = self.scale, self.zero_point.to(torch.float32)
                if self.is_per_channel:
                    view = [1, -1] + [1] * (X.dim() - 2)
                    scale, zero_point = scale.view(view), zero_point.view(view)
                else:
                    scale, zero_point = scale.reshape(()), zero_point.reshape(())
                X = self._quantizer.quantize(X, scale, zero_point, False)
        return X

    def activate_grad_theta(self):
        """lsq_module.py:211-214 (the parameters stay unused: adaptive rounding is not reproduced)."""
        self.theta.requires_grad = True
        self.gamma.requires_grad = True

    def calculate_grad_scale(self, quant_tensor):
        """(quant_max * numel [/ C]) ** -0.5  (lsq_module.py:317-340; its discarded temporaries are not computed)."""
        n = quant_tensor.numel()
        if self.is_per_channel:
            n /= quant_tensor.shape[1]
        return (self.quant_max * n) ** -0.5
