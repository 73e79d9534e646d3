"""B200-native DDPM reverse-diffusion sampling for the DANRA conditional UNets of TheaQG/DiffusionModelsCustom.

Public surface = the reference's own:
    from diffusionmodelscustom_b200 import Encoder, Decoder, DiffusionNet, DiffusionUtils
plus the north_star aliases ``UNet`` and ``Diffusion``.  Everything executes in ``libb200ddpm.so`` (hand-written
sm_100a CUDA behind the C ABI of ``include/b200ddpm.h``); there is no eager/CPU fallback.
"""
from .modules import Decoder, DecoderBlock, DiffusionNet, Encoder, ImageSelfAttention, SinusoidalEmbedding, UNet  # noqa: F401
from .diffusion import Diffusion, DiffusionUtils, DiffusionUtilsV2  # noqa: F401
from .unet_ms import UNet_downscale  # noqa: F401
from .generation import generate_ensemble, load_checkpoint, save_bundle  # noqa: F401
from . import evaluation  # noqa: F401
from .evaluation import SDFWeightedMSELoss  # noqa: F401
from . import unet  # noqa: F401  (DDPM_clean_application/src/unet.py generation: unet.Encoder / unet.Decoder / unet.DiffusionNet)

__all__ = ["Encoder", "Decoder", "DecoderBlock", "DiffusionNet", "ImageSelfAttention", "SinusoidalEmbedding", "UNet",
           "DiffusionUtils", "DiffusionUtilsV2", "Diffusion", "UNet_downscale", "generate_ensemble", "load_checkpoint",
           "save_bundle", "SDFWeightedMSELoss", "evaluation"]
