"""Parity cases: re-exported from the package (diffusionmodelscustom_b200/configs.py)."""
from diffusionmodelscustom_b200.configs import D_CASES, R_CASES, SAMPLE_CASES  # noqa: F401
