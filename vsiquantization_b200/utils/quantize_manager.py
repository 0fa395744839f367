"""QAT control API (reference: utils/quantize_manager.py:4-118): mode switches over every module that owns a
``weight_quantizer`` / ``activation_quantizer``.  Same names, arguments and semantics."""
from __future__ import annotations


def _managers(model, layer_names=None):
    for name, module in model.named_modules():
        for attr in ("weight_quantizer", "activation_quantizer"):
            if hasattr(module, attr) and (layer_names is None or name in layer_names):
                yield name, getattr(module, attr)


def calibrate_qat_model(model, dataloader, data_calib, device=None):
    """Observer mode on, learning and quantisation off, model.eval(), then ``data_calib(model, dataloader, device)``
    runs the forward passes (quantize_manager.py:4-31).  Each observed tensor costs one kernel launch and no host
    synchronisation; scales / zero-points are read back lazily afterwards."""
    for _, q in _managers(model):
        q.is_observer_qparam = True
        q.is_learning_scale = False
        q.is_quantize = False
    model.eval()
    data_calib(model, dataloader, device)


def activate_learning_qparam(model, layer_names=None, use_init=True, active=True):
    """Switch (selected) layers to learnable qparams: optional LSQ initialisation from the calibration statistics, then
    scale (and, for asymmetric quantisers, zero-point) become nn.Parameters (quantize_manager.py:34-66)."""
    for _, q in _managers(model, layer_names):
        q.is_learning_scale = active
        if use_init:
            q.init_scaling_factor_for_learning()
        if active:
            q.make_learn_qparameter()


def deactivate_learning_qparam(model, layer_names=None):
    activate_learning_qparam(model, layer_names=layer_names, active=False)


def activate_quantizer(model, layer_names=None, active=True):
    """Enable / disable fake quantisation in the forward pass (quantize_manager.py:83-103)."""
    for _, q in _managers(model, layer_names):
        q.is_quantize = active


def deactivate_quantizer(model, layer_names=None):
    activate_quantizer(model, layer_names=layer_names, active=False)
