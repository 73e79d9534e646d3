"""Loss and evaluation statistics on the device.

``SDFWeightedMSELoss`` mirrors ``DDPM_DANRA_conditional/training_DANRA_conditional.py:33-56`` (same constructor, same
``forward(input, target, sdf)``); without autograd (validation loss on sampled fields) it is ONE fused reduction kernel,
with autograd it is the reference's torch expression.  ``daily_errors`` / ``pixel_errors`` / ``histogram`` / ``bias`` are the
reductions ``evaluation_DANRA_conditional.py:121-170`` computes with ``nanmean`` / ``plt.hist`` after a host round trip: here
they run on the gathered ensemble where it already lives."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native as N


def _cuda_f32(t, name):
    if not t.is_cuda:
        raise N.NativeError(f"{name} must be a CUDA tensor (the statistics kernels have no CPU path)")
    return t.detach().to(torch.float32).contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class SDFWeightedMSELoss(nn.Module):
    def __init__(self, max_land_weight=1.0, min_sea_weight=0.5):
        super().__init__()
        self.max_land_weight = max_land_weight
        self.min_sea_weight = min_sea_weight

    def forward(self, input, target, sdf):
        if torch.is_grad_enabled() and (input.requires_grad or target.requires_grad):
            weights = torch.sigmoid(sdf) * (self.max_land_weight - self.min_sea_weight) + self.min_sea_weight
            return (weights * (input - target) ** 2).mean()
        return weighted_mse(input, target, sdf, self.max_land_weight, self.min_sea_weight)


def weighted_mse(input, target, sdf=None, max_land_weight=1.0, min_sea_weight=0.5):
    """mean(w (input - target)^2) with w = sigmoid(sdf) (max - min) + min; ``sdf=None`` is the plain MSE.  0-d CUDA tensor."""
    a, b = _cuda_f32(input, "input"), _cuda_f32(target, "target")
    if a.shape != b.shape:
        raise ValueError("input and target must have the same shape")
    s = None
    if sdf is not None:
        s = _cuda_f32(sdf.expand_as(input) if sdf.shape != input.shape else sdf, "sdf")
    out = torch.empty((), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        N.check(N.lib().b2d_op_weighted_mse(a.data_ptr(), b.data_ptr(), N.ptr(s), float(max_land_weight), float(min_sea_weight),
                                            out.data_ptr(), a.numel(), _stream()))
    return out


def daily_errors(gen, eval_):
    """Per-sample nan-aware (MAE, RMSE) over the trailing spatial dimensions: ``|g-e|.nanmean((1,2))`` and
    ``sqrt(((g-e)^2).nanmean((1,2)))`` of evaluation_DANRA_conditional.py:121-122.  gen/eval: [N, ...]."""
    g, e = _cuda_f32(gen, "gen"), _cuda_f32(eval_, "eval")
    if g.shape != e.shape:
        raise ValueError("gen and eval must have the same shape")
    n, hw = g.shape[0], g[0].numel()
    mae = torch.empty(n, device=g.device, dtype=torch.float32)
    rmse = torch.empty_like(mae)
    with torch.cuda.device(g.device):
        N.check(N.lib().b2d_op_eval_daily(g.data_ptr(), e.data_ptr(), mae.data_ptr(), rmse.data_ptr(), n, hw, _stream()))
    return mae, rmse


def pixel_errors(gen, eval_):
    """Per-pixel nan-aware (MAE, RMSE, bias) over the leading sample dimension; outputs have the per-sample shape."""
    g, e = _cuda_f32(gen, "gen"), _cuda_f32(eval_, "eval")
    if g.shape != e.shape:
        raise ValueError("gen and eval must have the same shape")
    n, hw = g.shape[0], g[0].numel()
    outs = [torch.empty(g.shape[1:], device=g.device, dtype=torch.float32) for _ in range(3)]
    with torch.cuda.device(g.device):
        N.check(N.lib().b2d_op_eval_pixel(g.data_ptr(), e.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), n,
                                          hw, _stream()))
    return tuple(outs)


def histogram(x, bins: int, range_):
    """``numpy.histogram(x, bins, range)`` counts on the device (NaN and out-of-range values dropped, right edge closed)."""
    v = _cuda_f32(x, "x").reshape(-1)
    lo, hi = float(range_[0]), float(range_[1])
    counts = torch.empty(bins, device=v.device, dtype=torch.int64)
    with torch.cuda.device(v.device):
        N.check(N.lib().b2d_op_histogram(v.data_ptr(), v.numel(), lo, hi, bins, counts.data_ptr(), _stream()))
    return counts


def bias(gen, eval_):
    """``nanmean(gen) - nanmean(eval)`` (the figure title of evaluation_DANRA_conditional.py:165)."""
    return torch.nanmean(_cuda_f32(gen, "gen")) - torch.nanmean(_cuda_f32(eval_, "eval"))
