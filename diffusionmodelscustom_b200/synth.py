"""Deterministic synthetic weights and DANRA-shaped inputs.

There is no network for checkpoints or data, and the reference ships no weights
(SURVEY.md §2 row 16), so parity and throughput runs use seeded synthetic state_dicts
whose keys/shapes are exactly the reference's (checked with ``load_state_dict(strict=True)``
in ``tests/golden/make_golden.py``) and whose distributions follow the reference's init
recipe (xavier_uniform on every Conv2d/ConvTranspose2d weight with bias 0.01,
``training_DANRA_conditional.py:739-753``; PyTorch defaults elsewhere).

Everything here is generated from a CPU ``torch.Generator`` so the same bits come out
in this container and on the GPU box (same torch build).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

ENC_CH = [64, 64, 128, 256, 512]          # modules_DANRA_conditional.py:170
DEC_IO = [(512, 256), (256, 128), (128, 64), (64, 64)]   # Decoder.make_layers :539-569


def _u(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def _xavier_conv(g, cout, cin, kh, kw, transposed=False):
    # torch's fan computation uses dim0/dim1 of the stored weight regardless of transposition
    rf = kh * kw
    if transposed:
        shape = (cin, cout, kh, kw)
        fan_in, fan_out = cout * rf, cin * rf
    else:
        shape = (cout, cin, kh, kw)
        fan_in, fan_out = cin * rf, cout * rf
    return _u(g, shape, math.sqrt(6.0 / (fan_in + fan_out)))


def _linear(g, out_f, in_f):
    b = 1.0 / math.sqrt(in_f)
    return _u(g, (out_f, in_f), b), _u(g, (out_f,), b)


def _bn(sd, g, prefix, c, randomize):
    if randomize:
        sd[prefix + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
    else:
        sd[prefix + ".weight"] = torch.ones(c)
        sd[prefix + ".bias"] = torch.zeros(c)
        sd[prefix + ".running_mean"] = torch.zeros(c)
        sd[prefix + ".running_var"] = torch.ones(c)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)


def _attn(sd, g, prefix, c, ff=False, ln_name="layernorm", mha_name="attention", randomize=True, ff_name="ff_self"):
    # LayerNorm affine is randomised a little so that gamma/beta handling is exercised
    sd[f"{prefix}.{ln_name}.weight"] = 1.0 + (0.1 * torch.randn(c, generator=g) if randomize else 0)
    sd[f"{prefix}.{ln_name}.bias"] = (0.05 * torch.randn(c, generator=g)) if randomize else torch.zeros(c)
    sd[f"{prefix}.{mha_name}.in_proj_weight"] = _u(g, (3 * c, c), math.sqrt(6.0 / (4 * c)))
    sd[f"{prefix}.{mha_name}.in_proj_bias"] = 0.02 * torch.randn(3 * c, generator=g)
    w, _ = _linear(g, c, c)
    sd[f"{prefix}.{mha_name}.out_proj.weight"] = w
    sd[f"{prefix}.{mha_name}.out_proj.bias"] = 0.02 * torch.randn(c, generator=g)
    if ff:
        sd[f"{prefix}.{ff_name}.0.weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[f"{prefix}.{ff_name}.0.bias"] = 0.05 * torch.randn(c, generator=g)
        sd[f"{prefix}.{ff_name}.1.weight"], sd[f"{prefix}.{ff_name}.1.bias"] = _linear(g, c, c)
        sd[f"{prefix}.{ff_name}.3.weight"], sd[f"{prefix}.{ff_name}.3.bias"] = _linear(g, c, c)


def synth_state_dict_r(c_in_total: int, c_out: int = 1, num_classes=None, img_hw=None,
                       has_lsm=False, has_topo=False, time_embedding: int = 256,
                       seed: int = 42, randomize_bn: bool = False, clean: bool = False):
    """Family R (``DiffusionNet(Encoder, Decoder)``) state_dict with reference keys.

    c_in_total counts x channels + lsm + topo + cond image channels (what ``conv1`` sees,
    modules_DANRA_conditional.py:157-164,178).
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    E = "encoder."
    # clean=True: the DDPM_clean_application/src/unet.py generation — no lsm/elevation buffers, attention = mha/layernorm/ff
    akw = dict(ff=True, mha_name="mha", ff_name="ff") if clean else {}
    if has_lsm and not clean:
        sd[E + "lsm"] = torch.zeros(1, *img_hw)
    if has_topo and not clean:
        sd[E + "elevation"] = torch.zeros(1, *img_hw)
    sd[E + "conv1.weight"] = _xavier_conv(g, 64, c_in_total, 8, 8)
    _bn(sd, g, E + "bn1", 64, randomize_bn)
    cin = 64
    for li, cout in enumerate(ENC_CH[1:], start=1):
        for bi in range(2):
            p = f"{E}layer{li}.{bi}."
            stride2 = (li > 1 and bi == 0)
            sd[p + "conv1.weight"] = _xavier_conv(g, cout, cin if bi == 0 else cout, 3, 3)
            _bn(sd, g, p + "bn1", cout, randomize_bn)
            sd[p + "conv2.weight"] = _xavier_conv(g, cout, cout, 3, 3)
            _bn(sd, g, p + "bn2", cout, randomize_bn)
            if stride2:
                sd[p + "downsample.0.weight"] = _xavier_conv(g, cout, cin, 1, 1)
                _bn(sd, g, p + "downsample.1", cout, randomize_bn)
        cin = cout
    for i, ch in enumerate(ENC_CH):
        w, b = _linear(g, ch, time_embedding)
        sd[f"{E}time_projection_layers.{i}.1.weight"] = w
        sd[f"{E}time_projection_layers.{i}.1.bias"] = b
    for i, ch in enumerate(ENC_CH):
        _attn(sd, g, f"{E}attention_layers.{i}", ch, **akw)
    sd[E + "conv2.weight"] = _xavier_conv(g, 64, 64, 8, 8)
    if num_classes is not None:
        sd[E + "label_emb.weight"] = torch.randn(num_classes, time_embedding, generator=g)
    D = "decoder."
    for i, (ci, co) in enumerate(DEC_IO):
        p = f"{D}residual_layers.{i}."
        _attn(sd, g, p + "attention", co, **akw)
        w, b = _linear(g, co, time_embedding)
        sd[p + "time_projection_layer.1.weight"] = w
        sd[p + "time_projection_layer.1.bias"] = b
        sd[p + "transpose.weight"] = _xavier_conv(g, ci, ci, 2, 2, transposed=True)
        sd[p + "transpose.bias"] = torch.full((ci,), 0.01)
        sd[p + "conv.weight"] = _xavier_conv(g, co, ci, 3, 3)
        sd[p + "conv.bias"] = torch.full((co,), 0.01)
    p = D + "final_layer."
    w, b = _linear(g, c_out, time_embedding)
    sd[p + "time_projection_layer.1.weight"] = w
    sd[p + "time_projection_layer.1.bias"] = b
    sd[p + "transpose.weight"] = _xavier_conv(g, 64, 64, 2, 2, transposed=True)
    sd[p + "transpose.bias"] = torch.full((64,), 0.01)
    sd[p + "conv.weight"] = _xavier_conv(g, c_out, 64, 3, 3)
    sd[p + "conv.bias"] = torch.full((c_out,), 0.01)
    return sd


def _double_conv(sd, g, prefix, cin, cout, mid=None):
    mid = mid or cout
    sd[prefix + ".double_conv.0.weight"] = _xavier_conv(g, mid, cin, 3, 3)
    sd[prefix + ".double_conv.1.weight"] = 1.0 + 0.1 * torch.randn(mid, generator=g)
    sd[prefix + ".double_conv.1.bias"] = 0.05 * torch.randn(mid, generator=g)
    sd[prefix + ".double_conv.3.weight"] = _xavier_conv(g, cout, mid, 3, 3)
    sd[prefix + ".double_conv.4.weight"] = 1.0 + 0.1 * torch.randn(cout, generator=g)
    sd[prefix + ".double_conv.4.bias"] = 0.05 * torch.randn(cout, generator=g)


def synth_state_dict_d(c_in: int = 2, c_out: int = 1, time_dim: int = 256, seed: int = 42):
    """Family D (``UNet_downscale``, DDPM_clean_application/src/unet_ms.py:103-136) state_dict."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    _double_conv(sd, g, "inc", c_in, 64)
    for name, ci, co in (("down1", 64, 128), ("down2", 128, 256), ("down3", 256, 256)):
        _double_conv(sd, g, f"{name}.maxpool_conv.1", ci, ci)
        _double_conv(sd, g, f"{name}.maxpool_conv.2", ci, co)
        sd[f"{name}.emb_layer.1.weight"], sd[f"{name}.emb_layer.1.bias"] = _linear(g, co, time_dim)
    for name, c in (("sa1", 128), ("sa2", 256), ("sa3", 256)):
        _attn(sd, g, name, c, ff=True, ln_name="ln", mha_name="mha")
    _double_conv(sd, g, "bot1", 256, 256)
    _double_conv(sd, g, "bot3", 256, 256)
    for name, ci, co in (("up1", 512, 128), ("up2", 256, 64), ("up3", 128, 64)):
        _double_conv(sd, g, f"{name}.conv.0", ci, ci)
        _double_conv(sd, g, f"{name}.conv.1", ci, co, ci // 2)
        sd[f"{name}.emb_layer.1.weight"], sd[f"{name}.emb_layer.1.bias"] = _linear(g, co, time_dim)
    for name, c in (("sa4", 128), ("sa5", 64), ("sa6", 64)):
        _attn(sd, g, name, c, ff=True, ln_name="ln", mha_name="mha")
    sd["outc.weight"] = _xavier_conv(g, c_out, 64, 1, 1)
    sd["outc.bias"] = torch.full((c_out,), 0.01)
    # reorder to the module registration order of the reference (cosmetic; load_state_dict is by key)
    return sd


def _smooth(g, b, h, w, coarse):
    import torch.nn.functional as F
    z = torch.randn(b, 1, max(h // coarse, 2), max(w // coarse, 2), generator=g)
    return F.interpolate(z, size=(h, w), mode="bilinear", align_corners=False)


def synth_inputs(batch: int, hw: int, seed: int = 42, has_lsm=True, has_topo=True, has_cond=True,
                 num_classes=None, c_hr: int = 1, lowres=None):
    """Synthetic DANRA-shaped fields: x_T~N(0,1); lsm in {0,1} (land≈0.45);
    topo = smooth non-negative metres × lsm; cond image = ERA5-like °C N(8.8, 6.3²), block-upsampled;
    y = uniform season class."""
    g = torch.Generator().manual_seed(seed)
    out = {"x": torch.randn(batch, c_hr, hw, hw, generator=g)}
    lsm = (_smooth(g, batch, hw, hw, 8) > 0.12).float()
    topo = (_smooth(g, batch, hw, hw, 8).abs() * 110.0) * lsm
    cond = 8.8 + 6.3 * _smooth(g, batch, hw, hw, 8)
    out["lsm"] = lsm if has_lsm else None
    out["topo"] = topo if has_topo else None
    out["cond"] = cond if has_cond else None
    out["y"] = (torch.randint(0, num_classes, (batch,), generator=g) if num_classes else None)
    if lowres is not None:
        out["y_lowres"] = 8.8 + 6.3 * torch.randn(batch, 1, lowres, lowres, generator=g)
    return out


def step_noise(batch: int, c_hr: int, hw: int, n_timesteps: int, seed: int = 1):
    """Host-generated z_i for i = T-1 … 1 (z at i == 1 is unused: diffusion_DANRA_conditional.py:149-152).
    Returned as [T, B, C, H, W] indexed by i (row 0 unused) so ref and ours consume identical noise."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n_timesteps, batch, c_hr, hw, hw, generator=g)
    z[0].zero_()
    z[1].zero_()
    return z
