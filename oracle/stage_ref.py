"""Stage the UNMODIFIED reference sources of the sampling path into oracle/_ref/ (git-ignored, travels to the GPU box).

TEST / BASELINE INFRASTRUCTURE ONLY — nothing under diffusionmodelscustom_b200/ imports this or oracle/_ref.

    python oracle/stage_ref.py [--check]

The reference is pure Python with no build system and no setup.py, so "installing" it is a byte-for-byte copy of the few
files the path needs (SURVEY.md §8(c) staging list), from where they lie under /root/reference, by this committed recipe.
The copies stay OUT of the repository history (``oracle/_ref/`` is in .gitignore): they exist so that ``bench.py --impl
reference`` and the ``cpu_baseline`` leg can time the reference's own code on the GPU box's host cores
(``cpu_baseline.kind == "reference"``), where /root/reference does not exist.  ``__graft_entry__.build()`` runs this whenever
/root/reference is present; a MANIFEST with sha256 of every staged file is written next to them and ``--check`` verifies it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("B2D_REFERENCE_ROOT", "/root/reference")

FILES = [
    "DDPM_DANRA_conditional/modules_DANRA_conditional.py",
    "DDPM_DANRA_conditional/modules_DANRA_flexible.py",
    "DDPM_DANRA_conditional/diffusion_DANRA_conditional.py",
    "DDPM_clean_application/__init__.py",
    "DDPM_clean_application/src/__init__.py",
    "DDPM_clean_application/src/unet.py",
    "DDPM_clean_application/src/unet_ms.py",
    "DDPM_clean_application/src/diffusion_modules.py",
    "DDPM_DANRA_Downscaling/modules_DANRA_downscaling.py",
    "DDPM_DANRA_Downscaling/diffusion_DANRA_downscaling.py",
]


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage() -> str:
    if not os.path.isdir(SRC):
        raise SystemExit(f"{SRC} not present (the GPU box only uses the already staged files)")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return DST


def available() -> bool:
    return os.path.exists(os.path.join(DST, "MANIFEST.json"))


def check() -> bool:
    if not available():
        return False
    m = json.load(open(os.path.join(DST, "MANIFEST.json")))["files"]
    return all(os.path.exists(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h for rel, h in m.items())


if __name__ == "__main__":
    if "--check" in sys.argv:
        print("staged and intact" if check() else "not staged")
    else:
        print(stage())
