"""Drop-in classes for the newest Family R generation, ``DDPM_clean_application/src/unet.py``:
``ImageSelfAttention`` has the LN-Linear-GELU-Linear tail (:91-119, keys ``mha.*``, ``layernorm.*``, ``ff.{0,1,3}.*``) and the
encoder is told ``cond_on_lsm`` / ``cond_on_topo`` instead of being handed buffers (:130-141); ``forward`` concatenates lsm /
topography whenever they are passed (:232-241).  Everything else — constructor arguments, ``state_dict`` keys, the positional
``model(x, t, y, cond_img, lsm_cond, topo_cond)`` contract — is as in ``modules.py``; the arithmetic is the same native program
with the attention FF enabled (``attn_ff=1``)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native as N
from . import modules as M


class ImageSelfAttention(nn.Module):
    def __init__(self, input_channels: int, n_heads: int):
        super().__init__()
        self.input_channels = input_channels
        self.n_heads = n_heads
        self.mha = nn.MultiheadAttention(self.input_channels, self.n_heads, batch_first=True)
        self.layernorm = nn.LayerNorm([self.input_channels])
        self.ff = nn.Sequential(nn.LayerNorm([self.input_channels]), nn.Linear(self.input_channels, self.input_channels),
                                nn.GELU(), nn.Linear(self.input_channels, self.input_channels))


class Encoder(M.Encoder):
    def __init__(self, input_channels: int, time_embedding: int, block=None, block_layers: list = [2, 2, 2, 2],
                 n_heads: int = 4, num_classes: int = None, cond_on_lsm=True, cond_on_topo=True, cond_on_img=False,
                 cond_img_dim=None):
        super().__init__(input_channels, time_embedding, block, block_layers, n_heads, num_classes, None, None, cond_on_img,
                         cond_img_dim)
        self.cond_on_lsm, self.cond_on_topo = bool(cond_on_lsm), bool(cond_on_topo)
        self.input_channels += int(self.cond_on_lsm) + int(self.cond_on_topo)
        self.conv1 = nn.Conv2d(self.input_channels, 64, kernel_size=(8, 8), stride=(2, 2), padding=(3, 3), bias=False)
        self.attention_layers = nn.ModuleList([ImageSelfAttention(ch, n_heads) for ch in M.FMAP_CHANNELS])


class DecoderBlock(M.DecoderBlock):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        if self.compute_attn:
            self.attention = ImageSelfAttention(self.output_channels, self.n_heads)


class Decoder(M.Decoder):
    def make_layers(self, n: int = 4):
        layers = []
        for i in range(n):
            in_ch = self.last_fmap_channels if i == 0 else layers[i - 1].output_channels
            out_ch = in_ch // 2 if i != (n - 1) else self.first_fmap_channels
            layers.append(DecoderBlock(in_ch, out_ch, time_embedding=self.time_embedding, compute_attn=True,
                                       n_heads=self.n_heads))
        return nn.ModuleList(layers)


class DiffusionNet(M.DiffusionNet):
    """``DiffusionNet(encoder, decoder)`` of src/unet.py:571-616."""

    def __init__(self, encoder: Encoder, decoder: Decoder):
        super().__init__(encoder, decoder)

    def _config(self, img_size, max_batch):
        cfg = super()._config(img_size, max_batch)
        cfg.attn_ff = 1
        return cfg
