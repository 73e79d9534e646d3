"""Builders shared by GPU model tests, smoke() and bench.py: our drop-in model with the seeded synthetic weights."""
import torch

import diffusionmodelscustom_b200 as P
from diffusionmodelscustom_b200 import synth


def build_ours_r(case, device="cuda"):
    H = case["hw"]
    if case.get("clean"):
        from diffusionmodelscustom_b200 import unet as U
        enc = U.Encoder(1, 256, cond_on_lsm=case["has_lsm"], cond_on_topo=case["has_topo"], cond_on_img=case["has_cond"],
                        cond_img_dim=(1, H, H) if case["has_cond"] else None, num_classes=case["num_classes"],
                        n_heads=case.get("n_heads", 4))
        dec = U.Decoder(512, 1, 256, 64, n_heads=case.get("n_heads", 4))
        net = U.DiffusionNet(enc, dec)
        sd = synth.synth_state_dict_r(case["c_in"], 1, case["num_classes"], (H, H), case["has_lsm"], case["has_topo"],
                                      seed=case["wseed"], randomize_bn=case["randomize_bn"], clean=True)
        net.load_state_dict(sd, strict=True)
        net.eval()
        return net.to(device), sd
    z = torch.zeros(1, H, H)
    enc = P.Encoder(1, 256, lsm_tensor=z if case["has_lsm"] else None, topo_tensor=z.clone() if case["has_topo"] else None,
                    cond_on_img=case["has_cond"], cond_img_dim=(1, H, H) if case["has_cond"] else None,
                    num_classes=case["num_classes"], n_heads=case.get("n_heads", 4))
    dec = P.Decoder(512, 1, 256, 64, n_heads=case.get("n_heads", 4))
    net = P.DiffusionNet(enc, dec)
    sd = synth.synth_state_dict_r(case["c_in"], 1, case["num_classes"], (H, H), case["has_lsm"], case["has_topo"],
                                  seed=case["wseed"], randomize_bn=case["randomize_bn"])
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net.to(device), sd


def inputs_r(case, batch=None, device="cuda"):
    B = batch or case["batch"]
    inp = synth.synth_inputs(B, case["hw"], seed=case["iseed"], has_lsm=case["has_lsm"], has_topo=case["has_topo"],
                             has_cond=case["has_cond"], num_classes=case["num_classes"])
    dev = {k: (v.to(device) if v is not None else None) for k, v in inp.items()}
    return inp, dev


def build_ours_d(case, device="cuda"):
    net = P.UNet_downscale(c_in=case["c_in"], c_out=1, time_dim=256, interp_mode="bicubic", img_size=case["hw"], device=device)
    sd = synth.synth_state_dict_d(case["c_in"], 1, seed=case["wseed"])
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net.to(device), sd


def inputs_d(case, batch=None, device="cuda"):
    B = batch or case["batch"]
    inp = synth.synth_inputs(B, case["hw"], seed=case["iseed"], lowres=case["lowres"])
    dev = {k: (v.to(device) if v is not None else None) for k, v in inp.items()}
    return inp, dev
