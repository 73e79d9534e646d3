"""Drop-in classes for the Downscaling generation of Family R — ``DDPM_DANRA_Downscaling/modules_DANRA_downscaling.py`` and
``diffusion_DANRA_downscaling.py``: the UNCONDITIONAL ResNet-UNet (``Encoder(input_channels, time_embedding, block, block_layers,
n_heads)``, ``forward(x, t)``; no lsm / topo / cond_img / label inputs) whose encoder embeds ``t`` with the interleaved
base-10000 ``SinusoidalEmbedding`` (:190) instead of ``pos_encoding``, driven by ``DiffusionUtils.sample(x, model)`` (:98-142).
Same native program as ``modules.py`` with ``stem_embedding = 1``."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import modules as M
from .diffusion import DiffusionUtils as _DiffusionUtils

SinusoidalEmbedding = M.SinusoidalEmbedding
ImageSelfAttention = M.ImageSelfAttention
DecoderBlock = M.DecoderBlock
Decoder = M.Decoder


class Encoder(M.Encoder):
    def __init__(self, input_channels: int, time_embedding: int, block=None, block_layers: list = [2, 2, 2, 2], n_heads: int = 4):
        super().__init__(input_channels, time_embedding, block, block_layers, n_heads, None, None, None, False, None)


class DiffusionNet(M.DiffusionNet):
    """``DiffusionNet(encoder, decoder)`` of modules_DANRA_downscaling.py:489-520; ``forward(x, t)``."""

    def __init__(self, encoder: Encoder, decoder: Decoder):
        super().__init__(encoder, decoder)

    def _config(self, img_size, max_batch):
        cfg = super()._config(img_size, max_batch)
        cfg.stem_embedding = 1
        return cfg

    @torch.no_grad()
    def forward(self, x: torch.Tensor, t: torch.Tensor, *unused):
        return super().forward(x, t)


class DiffusionUtils(_DiffusionUtils):
    """``sample(x, model)`` of diffusion_DANRA_downscaling.py:98-142 (same posterior arithmetic as the conditional generation)."""

    def sample(self, x: torch.Tensor, model: nn.Module, *, noise: torch.Tensor = None, seed: int = None, sample_offset: int = 0):
        out, _ = self._run(x, model, None, None, None, None, noise, seed, sample_offset)
        return out
