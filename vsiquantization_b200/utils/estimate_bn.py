"""BN statistics re-estimation and gradient-scale calibration (reference: utils/estimate_bn.py:8-161).

reestimate_BN_stats: for every ConvBnReLU that kept its BatchNorm (is_fuse_bn=False) the running mean / variance are
replaced by the average over ``num_batches`` batches of the per-batch mean and UNBIASED variance of the conv output
(the reference gets these by running the BN layer in training mode with momentum 1, :60-65, :82-87, :96-97).  Here the
per-batch moments come from ONE pass of the observer kernel over the conv output (min/max ride along for free), the
sums are accumulated and finalised by small kernels, and the layer's output during re-estimation is normalised with
the batch statistics exactly as training-mode BN would.  With torch.distributed initialised (``sync=True``) the
per-channel sums are all-reduced first (SyncBN-style), so the result equals single-process re-estimation over the
global batch."""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F

from .. import ops
from ..modules.fused import ConvBnReLU
from ..quantizers.fake_quantize import FakeQuantize


class ReestimateBNStats:
    """Callable wrapper for training engines (estimate_bn.py:8-35)."""

    def __init__(self, model, data_loader, num_batches=50):
        self.model = model
        self.data_loader = data_loader
        self.num_batches = num_batches

    def __call__(self, engine):
        print("-- Reestimate current BN statistics --")
        reestimate_BN_stats(self.model, self.data_loader, self.num_batches)


def _dist_on(sync: bool) -> bool:
    d = torch.distributed
    return bool(sync and d.is_available() and d.is_initialized() and d.get_world_size() > 1)


def _make_hook(ctx: dict):
    """ctx carries the per-iteration facts the hook needs: ``sync`` (all-reduce the sums), ``weight`` (0.0 when this rank
    only replays a batch to keep the collectives aligned), ``local_images`` / ``global_images`` (this rank's and all
    ranks' image count of the iteration -- the global per-channel element count follows without a per-layer sync)."""
    def hook(module: ConvBnReLU, x: torch.Tensor, act=None, collect=None) -> torch.Tensor:
        """Moments of this batch -> running sums; returns training-mode BN of x, with ``act`` ("relu" / "silu" / None)
        applied (the layer hands its activation over so that normalise + ReLU are one pass on channels_last tensors).
        ``collect``: the layer's output QuantizationManager while it is only observing -- normalise + activation + its
        observer then run as ONE pass (vsiq_ci_epilogue_observe) and the returned tensor is the layer's final output."""
        bn = module.bn
        stats = ops.observe(x, ch_axis=1)                      # [C,5]: .., sum x, sum x^2  -- one read of x
        count = float(x.numel() // x.shape[1])
        if ctx["sync"] and ctx.get("peer") is not None and x.shape[1] <= ctx["peer_max_channels"]:
            # one kernel per rank over NVLink peer memory: publish this shard's sums, wait for the peers', add them in
            # rank order, finish the moments (csrc/peer_exchange.cu) -- no NCCL call, no separate moments launch
            count = count / ctx["local_images"] * ctx["global_images"]
            mean, var_b, _ = ctx["peer"].bn_moments(stats, ctx["weight"], count, module.running_mean_sum,
                                                    module.running_var_sum)
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
            if collect is not None:
                y = collect.collect_epilogue(x, act, bn=(mean, var_b, bn.weight, bn.bias, bn.eps))
                return y if y is not None else collect.quantize(_normalise(x, mean, var_b, bn, act))
            return _normalise(x, mean, var_b, bn, act)
        if ctx["sync"]:
            if ctx["weight"] != 1.0:
                stats.mul_(ctx["weight"])
            # the whole [C,5] block in place, one collective and no staging copies (the min / max columns come back as
            # sums and are not used by the moments)
            torch.distributed.all_reduce(stats, op=torch.distributed.ReduceOp.SUM)
            count = count / ctx["local_images"] * ctx["global_images"]  # ranks may hold different batch sizes
        mean, var_b, _ = ops.bn_moments_finalize(stats, count, module.running_mean_sum, module.running_var_sum)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        # training-mode BN normalises with the batch mean and the BIASED batch variance
        if collect is not None:
            y = collect.collect_epilogue(x, act, bn=(mean, var_b, bn.weight, bn.bias, bn.eps))
            return y if y is not None else collect.quantize(_normalise(x, mean, var_b, bn, act))
        return _normalise(x, mean, var_b, bn, act)
    return hook


def _normalise(x, mean, var_b, bn, act):
    if act in (None, "relu") and ops.ci_supported(x):
        return ops.ci_bn_normalize(x, mean, var_b, bn.weight, bn.bias, bn.eps, relu=act == "relu")
    y = F.batch_norm(x, mean, var_b, bn.weight, bn.bias, False, 0.0, bn.eps)
    if act == "relu":
        return F.relu(y)
    return F.silu(y) if act == "silu" else y


def reestimate_BN_stats(model, data_loader, num_batches=50, store_ema_stats=False, sync=True):
    """Same signature as estimate_bn.py:38 (+ ``sync``).  Requires layers built with is_fuse_bn=False (the reference
    needs ``module.bn`` too, :60).

    With torch.distributed initialised and ``sync=True`` the per-channel sums of every layer are all-reduced (SyncBN-style):
    iteration k re-estimates from the union of every rank's k-th batch.  Ranks may hold different numbers of batches and
    different batch sizes: one tiny all-reduce per iteration tells every rank how many images take part, a rank that has
    run out of data replays its last batch with weight zero so the per-layer collectives stay aligned, and the loop ends
    when no rank has data left."""
    model.eval()
    layers = [(n, m) for n, m in model.named_modules() if isinstance(m, ConvBnReLU) and hasattr(m, "bn")]
    ctx = {"sync": _dist_on(sync), "weight": 1.0, "local_images": 1, "global_images": 1, "peer": None}
    if ctx["sync"]:
        from .. import _lib
        from ..parallel import peer_exchange_for
        ctx["peer"] = peer_exchange_for(None)
        ctx["peer_max_channels"] = _lib.lib.vsiq_peer_max_channels()
    hook = _make_hook(ctx)
    for _, m in layers:
        m.running_mean_sum = torch.zeros_like(m.bn.running_mean)
        m.running_var_sum = torch.zeros_like(m.bn.running_var)
        if store_ema_stats:
            if not hasattr(m, "running_mean_ema"):
                m.register_buffer("running_mean_ema", copy.deepcopy(m.bn.running_mean))
                m.register_buffer("running_var_ema", copy.deepcopy(m.bn.running_var))
            else:
                m.running_mean_ema = copy.deepcopy(m.bn.running_mean)
                m.running_var_ema = copy.deepcopy(m.bn.running_var)
        m._bn_reestimate = hook
    device = next(model.parameters()).device
    batch_count = 0
    it = iter(data_loader)
    last = None
    taken = 0
    try:
        with torch.no_grad():
            while True:
                item = next(it, None) if taken < num_batches else None
                if item is not None:
                    taken += 1
                    last = item[0]
                if ctx["sync"]:
                    n_local = float(item[0].shape[0]) if item is not None else 0.0
                    t = torch.tensor([1.0 if item is not None else 0.0, n_local], dtype=torch.float64, device=device)
                    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
                    n_have, n_global = t.tolist()  # the one host synchronisation of the iteration
                    if n_have == 0:
                        break
                    if last is None:
                        raise RuntimeError("reestimate_BN_stats(sync=True): this rank has no batch at all")
                    ctx["weight"] = 1.0 if item is not None else 0.0
                    ctx["local_images"], ctx["global_images"] = float(last.shape[0]), n_global
                elif item is None:
                    break
                imgs = last.to(device, non_blocking=True).float() / 255.0
                model(imgs)
                batch_count += 1
    finally:
        for _, m in layers:
            m._bn_reestimate = None
    if ctx["peer"] is not None:
        done, failed = ctx["peer"].status()
        bad = torch.tensor([1.0 if failed else 0.0], device=device)
        torch.distributed.all_reduce(bad, op=torch.distributed.ReduceOp.MAX)  # every rank raises, or none does
        if float(bad.item()):
            from ..parallel import drop_peer_exchange
            drop_peer_exchange(None)  # later calls use the NCCL collective
            raise RuntimeError(f"reestimate_BN_stats: a peer-memory exchange timed out (this rank: exchange {failed or '-'} "
                               f"after {done} completed); the running statistics were not updated. Later calls use the "
                               "NCCL collective; VSIQ_PEER_EXCHANGE=0 selects it from the start")
    if batch_count:
        for _, m in layers:
            ops.bn_reestimate_finish(m.running_mean_sum, m.running_var_sum, batch_count, m.bn.running_mean,
                                     m.bn.running_var)
    model.eval()


def pdf_cdf(x, mu, sigma):
    """Normal pdf / cdf (estimate_bn.py:145-162)."""
    dist = torch.distributions.Normal(loc=mu, scale=sigma)
    return torch.exp(dist.log_prob(x)), dist.cdf(x)


def compute_scale(model, data_loader=None, num_batches=50, store_ema_stats=False):
    """Per-channel calib_grad_scale of the activation quantiser from the BN affine parameters and the weight moments
    (estimate_bn.py:104-139).  [C]-sized host-of-the-hot-path math: plain torch ops on tiny vectors."""
    model.eval()
    for _, m in model.named_modules():
        if isinstance(m, ConvBnReLU) and hasattr(m, "bn"):
            w = m.conv_fuse.weight.data.detach()
            mu_w, sigma_w = w.mean(), w.std()
            mu_a, sigma_a = m.bn.bias.data.clone(), m.bn.weight.data.clone()
            pdf, cdf = pdf_cdf(mu_a / sigma_a, 0, 1)
            m.activation_quantizer.quantizer.calib_grad_scale = 1 / (
                (mu_w ** 2 + sigma_w ** 2) / ((mu_a ** 2 + sigma_a ** 2) * cdf + mu_a * sigma_a * pdf))
    model.eval()


__all__ = ["ReestimateBNStats", "reestimate_BN_stats", "compute_scale", "pdf_cdf", "FakeQuantize"]
