"""CPU oracle for the DDPM reverse-diffusion sampling path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file; the product path (``diffusionmodelscustom_b200``) never does.

This is a *functional FP32 restatement* (plain ``torch`` tensor algebra on CPU, driven directly by a
reference-keyed ``state_dict``) of the reference's algorithm; every function cites the reference
lines it follows.  The reference is Python, so the restatement is Python too; the arithmetic that
lives in third-party torchvision (``BasicBlock.forward``, v0.16.1 pinned / v0.26 here,
``torchvision/models/resnet.py:59-105``) is restated from its published structure.

Pinning: the reference holds no golden vectors or tests for this path, so this oracle is
pinned against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every fixture.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- schedules
def beta_schedule(n_timesteps, beta_min, beta_max, scheduler="linear", version=1):
    """diffusion_DANRA_conditional.py:53-77 (v1) / DDPM_clean_application/src/diffusion_modules.py:50-69 (v2)."""
    if scheduler == "linear":
        return torch.linspace(beta_min, beta_max, n_timesteps)
    if version == 1:
        T = n_timesteps - 1
        betas = [beta_min + 0.5 * (beta_max - beta_min) * (1 + math.cos((i / T) * math.pi))
                 for i in reversed(range(n_timesteps))]
        return torch.tensor(betas, dtype=torch.float64).float()
    t = torch.linspace(0, n_timesteps, n_timesteps + 1)
    ft = torch.cos(((t / n_timesteps + 0.008) / 1.008) * math.pi / 2) ** 2
    alphat = ft / ft[0]
    return torch.clip(1 - alphat[1:] / alphat[:-1], 0.0001, 0.9999)


def schedule_tables(n_timesteps, beta_min, beta_max, scheduler="linear", version=1):
    """diffusion_DANRA_conditional.py:47-51."""
    betas = beta_schedule(n_timesteps, beta_min, beta_max, scheduler, version)
    alphas = 1 - betas
    alpha_hat = torch.cumprod(alphas, dim=0)
    return betas, alphas, alpha_hat


def posterior_update(x, eps, z, i, betas, alphas, alpha_hat):
    """diffusion_DANRA_conditional.py:135-157 — coefficients indexed by i (not i-1); sigma = sqrt(beta)."""
    alpha, beta, ahat = alphas[i], betas[i], alpha_hat[i]
    x = (1 / torch.sqrt(alpha)) * (x - ((1 - alpha) / torch.sqrt(1 - ahat)) * eps)
    return x + torch.sqrt(beta) * z


# --------------------------------------------------------------------------- shared pieces
def image_self_attention(sd, p, x, n_heads, ln="layernorm", mha="attention", ff=False, ffp="ff_self"):
    """modules_DANRA_conditional.py:91-110 (ImageSelfAttention) / unet_ms.py:21-27 (SelfAttention, ff=True);
    formulas verified against nn.MultiheadAttention in SURVEY.md Appendix A."""
    N, C, H, W = x.shape
    tok = x.reshape(N, C, H * W).permute(0, 2, 1)
    xn = F.layer_norm(tok, (C,), sd[f"{p}.{ln}.weight"], sd[f"{p}.{ln}.bias"], 1e-5)
    qkv = xn @ sd[f"{p}.{mha}.in_proj_weight"].t() + sd[f"{p}.{mha}.in_proj_bias"]
    q, k, v = qkv.split(C, dim=-1)
    d = C // n_heads
    L = H * W

    def heads(t):
        return t.reshape(N, L, n_heads, d).permute(0, 2, 1, 3)

    q, k, v = heads(q), heads(k), heads(v)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    o = torch.softmax(s, dim=-1) @ v
    o = o.permute(0, 2, 1, 3).reshape(N, L, C)
    o = o @ sd[f"{p}.{mha}.out_proj.weight"].t() + sd[f"{p}.{mha}.out_proj.bias"] + tok
    if ff:
        h = F.layer_norm(o, (C,), sd[f"{p}.{ffp}.0.weight"], sd[f"{p}.{ffp}.0.bias"], 1e-5)
        h = F.gelu(h @ sd[f"{p}.{ffp}.1.weight"].t() + sd[f"{p}.{ffp}.1.bias"])
        o = h @ sd[f"{p}.{ffp}.3.weight"].t() + sd[f"{p}.{ffp}.3.bias"] + o
    return o.permute(0, 2, 1).reshape(N, C, H, W)


def _bn_eval(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _basic_block(sd, p, x, stride):
    """torchvision BasicBlock.forward (resnet.py:59-105): conv3x3(s)-bn-relu-conv3x3-bn (+downsample(x)) -relu."""
    out = F.relu(_bn_eval(sd, p + "bn1", F.conv2d(x, sd[p + "conv1.weight"], None, stride, 1)))
    out = _bn_eval(sd, p + "bn2", F.conv2d(out, sd[p + "conv2.weight"], None, 1, 1))
    if (p + "downsample.0.weight") in sd:
        x = _bn_eval(sd, p + "downsample.1", F.conv2d(x, sd[p + "downsample.0.weight"], None, stride, 0))
    return F.relu(out + x)


def enc_time_embedding(t, channels=256):
    """Encoder.pos_encoding, modules_DANRA_conditional.py:203-211: base 1000, layout [sin | cos]."""
    t = t.unsqueeze(-1).float()
    inv_freq = 1.0 / (1000 ** (torch.arange(0, channels, 2).float() / channels))
    return torch.cat([torch.sin(t.repeat(1, channels // 2) * inv_freq),
                      torch.cos(t.repeat(1, channels // 2) * inv_freq)], dim=-1)


def dec_time_embedding(t, dim=256, n=10000):
    """SinusoidalEmbedding.forward, modules_DANRA_conditional.py:42-63: interleaved sin,cos, base 10000.
    The reference divides an int64 0-dim tensor by a Python float => FP32 result."""
    div = torch.tensor([n ** (2 * i / dim) for i in range(dim // 2)], dtype=torch.float64)
    # int64 tensor / python float -> float32 division by the float32-rounded... torch promotes the
    # python scalar to the default dtype: emb = float32(t) / float32(div)
    emb = t.float().unsqueeze(-1) / div.float()
    out = torch.zeros(t.shape[0], dim)
    out[:, 0::2] = torch.sin(emb)
    out[:, 1::2] = torch.cos(emb)
    return out


def _tproj(sd, p, temb):
    return F.silu(temb) @ sd[p + ".1.weight"].t() + sd[p + ".1.bias"]


# --------------------------------------------------------------------------- Family R
def family_r_forward(sd, x, t, y=None, cond_img=None, lsm_cond=None, topo_cond=None, n_heads=4,
                     has_lsm=None, has_topo=None, taps=None, downscaling=False):
    """DiffusionNet.forward (modules_DANRA_conditional.py:597-616) = Decoder(*Encoder(...), t).

    has_lsm/has_topo mirror ``hasattr(self,'lsm')`` / ``hasattr(self,'elevation')`` (:228-233) and
    default to the presence of the registered buffers in the state_dict.  A state_dict of the newer generation
    (DDPM_clean_application/src/unet.py: attention keys ``mha``/``ff``, :91-119; lsm/topo concatenated whenever they are
    passed, :232-241) is recognised by its keys."""
    E = "encoder."
    clean = (E + "attention_layers.0.mha.in_proj_weight") in sd
    akw = dict(mha="mha", ff=True, ffp="ff") if clean else {}
    if clean:
        has_lsm = lsm_cond is not None if has_lsm is None else has_lsm
        has_topo = topo_cond is not None if has_topo is None else has_topo
    has_lsm = (E + "lsm") in sd if has_lsm is None else has_lsm
    has_topo = (E + "elevation") in sd if has_topo is None else has_topo
    # Encoder.forward :228-238 — concat order [x, lsm, topo, cond_img]
    if has_lsm:
        x = torch.cat([x, lsm_cond], dim=1)
    if has_topo:
        x = torch.cat([x, topo_cond], dim=1)
    if cond_img is not None:
        x = torch.cat((x, cond_img), dim=1)
    # :243-244; the Downscaling generation embeds t with SinusoidalEmbedding instead (modules_DANRA_downscaling.py:190)
    temb = dec_time_embedding(t) if downscaling else enc_time_embedding(t)
    if y is not None:
        temb = temb + sd[E + "label_emb.weight"][y]                 # :256
    fm = []
    # :260-266 (no norm/activation between conv1 and attention)
    f = F.conv2d(x, sd[E + "conv1.weight"], None, 2, 3)
    f = f + _tproj(sd, E + "time_projection_layers.0", temb)[:, :, None, None]
    f = image_self_attention(sd, E + "attention_layers.0", f, n_heads, **akw)
    fm.append(f)
    # :269-282
    h = F.relu(_bn_eval(sd, E + "bn1", F.conv2d(f, sd[E + "conv2.weight"], None, 2, 3)))
    for li in range(1, 5):                                          # :276-309
        for bi in range(2):
            h = _basic_block(sd, f"{E}layer{li}.{bi}.", h, 2 if (li > 1 and bi == 0) else 1)
        h = h + _tproj(sd, f"{E}time_projection_layers.{li}", temb)[:, :, None, None]
        h = image_self_attention(sd, f"{E}attention_layers.{li}", h, n_heads, **akw)
        fm.append(h)
        if taps is not None:
            taps[f"fmap{li + 1}"] = h
    if taps is not None:
        taps["fmap1"] = fm[0]
    # Decoder.forward :512-536
    D = "decoder."
    dtemb = dec_time_embedding(t)
    out = fm[4]
    for i in range(4):
        p = f"{D}residual_layers.{i}."
        out = _decoder_block(sd, p, out, fm[3 - i], dtemb, n_heads, akw)
        if taps is not None:
            taps[f"dec{i}"] = out
    # final_layer: no skip, no t, no attention, IN2 = Identity, act = Identity (:503-509, :535)
    p = D + "final_layer."
    out = F.conv_transpose2d(out, sd[p + "transpose.weight"], sd[p + "transpose.bias"], stride=2)
    out = F.instance_norm(out, eps=1e-5)
    return F.conv2d(out, sd[p + "conv.weight"], sd[p + "conv.bias"], 1, 1)


def _decoder_block(sd, p, fmap, prev, dtemb, n_heads, akw=None):
    """DecoderBlock.forward, modules_DANRA_conditional.py:425-460."""
    out = F.conv_transpose2d(fmap, sd[p + "transpose.weight"], sd[p + "transpose.bias"], stride=2)
    out = F.instance_norm(out, eps=1e-5)                            # affine=False, batch stats always
    out = F.conv2d(out, sd[p + "conv.weight"], sd[p + "conv.bias"], 1, 1)
    out = F.instance_norm(out, eps=1e-5)
    out = out + prev
    out = out + _tproj(sd, p + "time_projection_layer", dtemb)[:, :, None, None]
    out = image_self_attention(sd, p + "attention", out, n_heads, **(akw or {}))
    return F.relu(out)


# --------------------------------------------------------------------------- Family D
def _double_conv(sd, p, x, residual=False):
    """DoubleConv.forward, unet_ms.py:30-49."""
    h = F.conv2d(x, sd[p + ".double_conv.0.weight"], None, 1, 1)
    h = F.gelu(F.group_norm(h, 1, sd[p + ".double_conv.1.weight"], sd[p + ".double_conv.1.bias"], 1e-5))
    h = F.conv2d(h, sd[p + ".double_conv.3.weight"], None, 1, 1)
    h = F.group_norm(h, 1, sd[p + ".double_conv.4.weight"], sd[p + ".double_conv.4.bias"], 1e-5)
    return F.gelu(x + h) if residual else h


def family_d_time_embedding(t, channels=256):
    """UNet_downscale.pos_encoding, unet_ms.py:138-146: base 10000, layout [sin | cos]."""
    t = t.unsqueeze(-1).float()
    inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
    return torch.cat([torch.sin(t.repeat(1, channels // 2) * inv_freq),
                      torch.cos(t.repeat(1, channels // 2) * inv_freq)], dim=-1)


def family_d_forward(sd, x, t, y_lowres, interp_mode="bicubic", time_dim=256, taps=None):
    """UNet_downscale.forward, unet_ms.py:148-179."""
    temb = family_d_time_embedding(t, time_dim)
    if y_lowres is not None:
        yy = F.interpolate(y_lowres.float(), size=[x.shape[-1], x.shape[-2]], mode=interp_mode)
    else:
        yy = torch.zeros_like(x)
    x = torch.cat([x, yy], dim=1)

    def down(p, h):                                                  # Down.forward :70-73
        h = F.max_pool2d(h, 2)
        h = _double_conv(sd, p + ".maxpool_conv.1", h, residual=True)
        h = _double_conv(sd, p + ".maxpool_conv.2", h)
        return h + _tproj(sd, p + ".emb_layer", temb)[:, :, None, None]

    def up(p, h, skip):                                              # Up.forward :95-100
        h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=True)
        h = torch.cat([skip, h], dim=1)
        h = _double_conv(sd, p + ".conv.0", h, residual=True)
        h = _double_conv(sd, p + ".conv.1", h)
        return h + _tproj(sd, p + ".emb_layer", temb)[:, :, None, None]

    def sa(p, h):
        return image_self_attention(sd, p, h, 4, ln="ln", mha="mha", ff=True)

    x1 = _double_conv(sd, "inc", x)
    x2 = sa("sa1", down("down1", x1))
    x3 = sa("sa2", down("down2", x2))
    x4 = sa("sa3", down("down3", x3))
    x4 = _double_conv(sd, "bot1", x4)
    x4 = _double_conv(sd, "bot3", x4)
    u1 = sa("sa4", up("up1", x4, x3))
    u2 = sa("sa5", up("up2", u1, x2))
    u3 = sa("sa6", up("up3", u2, x1))
    if taps is not None:
        taps.update(x1=x1, x2=x2, x3=x3, bot=x4, u1=u1, u2=u2, u3=u3)
    return F.conv2d(u3, sd["outc.weight"], sd["outc.bias"])


# --------------------------------------------------------------------------- sampling loop
@torch.no_grad()
def sample(model_fn, x, n_timesteps, beta_min, beta_max, noise=None, scheduler="linear", version=1,
           generator=None, record_at=None):
    """DiffusionUtils.sample, diffusion_DANRA_conditional.py:105-159: i = T-1 … 1 (T-1 evaluations),
    z ~ N(0,1) for i > 1 and 0 at i == 1.  ``model_fn(x, t_long[B])`` returns eps_hat.
    ``noise`` ([T,B,C,H,W], indexed by i) injects host-generated z so two implementations consume the
    same draws; otherwise ``torch.randn`` with ``generator``."""
    betas, alphas, alpha_hat = schedule_tables(n_timesteps, beta_min, beta_max, scheduler, version)
    rec = {}
    for i in reversed(range(1, n_timesteps)):
        t = (torch.ones(x.shape[0]) * i).long()
        if record_at is not None and i in record_at:
            rec[i] = {"x": x.clone()}
        eps = model_fn(x, t)
        if record_at is not None and i in record_at:
            rec[i]["eps"] = eps.clone()
        if i > 1:
            z = noise[i] if noise is not None else torch.randn(x.shape, generator=generator)
        else:
            z = torch.zeros_like(x)
        x = posterior_update(x, eps, z, i, betas, alphas, alpha_hat)
    return (x, rec) if record_at is not None else x
