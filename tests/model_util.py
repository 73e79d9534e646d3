"""Model/input builders: re-exported from the package (diffusionmodelscustom_b200/configs.py)."""
from diffusionmodelscustom_b200.configs import build_ours_d, build_ours_r, inputs_d, inputs_r  # noqa: F401
