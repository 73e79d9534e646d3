"""Generate golden fixtures by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py [--long]

For every case the reference model is built with the reference's own constructors, our seeded
synthetic state_dict is loaded with ``strict=True`` (which also pins key names and shapes), the
reference forward / ``DiffusionUtils.sample`` is run on seeded synthetic inputs, and only the
OUTPUTS are stored (inputs and weights regenerate bit-identically from their seeds via
``diffusionmodelscustom_b200.synth``).  Noise for ``sample`` is injected by patching
``torch.randn_like`` so that the reference consumes host-generated z_i.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "DDPM_DANRA_conditional"))
sys.path.insert(0, REF)

from diffusionmodelscustom_b200 import synth  # noqa: E402
from tests.cases import R_CASES, D_CASES, SAMPLE_CASES  # noqa: E402

torch.set_num_threads(8)


def build_ref_r(case, module_name="modules_DANRA_conditional"):
    H = case["hw"]
    if case.get("clean"):
        from DDPM_clean_application.src import unet as mod
        enc = mod.Encoder(1, 256, cond_on_lsm=case["has_lsm"], cond_on_topo=case["has_topo"], cond_on_img=case["has_cond"],
                          cond_img_dim=(1, H, H) if case["has_cond"] else None, num_classes=case["num_classes"],
                          n_heads=case.get("n_heads", 4))
        dec = mod.Decoder(512, 1, 256, 64, n_heads=case.get("n_heads", 4))
        net = mod.DiffusionNet(enc, dec)
        sd = synth.synth_state_dict_r(case["c_in"], 1, case["num_classes"], (H, H), case["has_lsm"], case["has_topo"],
                                      seed=case["wseed"], randomize_bn=case["randomize_bn"], clean=True)
        net.load_state_dict(sd, strict=True)
        net.eval()
        return net
    mod = __import__(module_name)
    lsm = torch.zeros(1, H, H) if case["has_lsm"] else None
    topo = torch.zeros(1, H, H) if case["has_topo"] else None
    enc = mod.Encoder(1, 256, lsm_tensor=lsm, topo_tensor=topo, cond_on_img=case["has_cond"],
                      cond_img_dim=(1, H, H) if case["has_cond"] else None, num_classes=case["num_classes"],
                      n_heads=case.get("n_heads", 4))
    dec = mod.Decoder(512, 1, 256, 64, n_heads=case.get("n_heads", 4))
    net = mod.DiffusionNet(enc, dec)
    sd = synth.synth_state_dict_r(case["c_in"], 1, case["num_classes"], (H, H), case["has_lsm"], case["has_topo"],
                                  seed=case["wseed"], randomize_bn=case["randomize_bn"])
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net


def run_r(case, module_name="modules_DANRA_conditional"):
    net = build_ref_r(case, module_name)
    inp = synth.synth_inputs(case["batch"], case["hw"], seed=case["iseed"], has_lsm=case["has_lsm"],
                             has_topo=case["has_topo"], has_cond=case["has_cond"], num_classes=case["num_classes"])
    out = {}
    with torch.no_grad():
        for t in case["ts"]:
            tt = torch.full((case["batch"],), t, dtype=torch.long)
            x = inp["x"] * case.get("x_scale", 1.0)
            out[f"eps_t{t}"] = net(x, tt, inp["y"], inp["cond"], inp["lsm"], inp["topo"]).numpy()
    return out, net, inp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", action="store_true", help="also run the T=1000 free-running case (minutes)")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()

    for name, case in R_CASES.items():
        if args.only and args.only != name:
            continue
        out, _, _ = run_r(case, case.get("module", "modules_DANRA_conditional"))
        np.savez_compressed(os.path.join(HERE, f"r_{name}.npz"), **out)
        print("R", name, {k: (v.shape, float(np.abs(v).mean())) for k, v in out.items()})

    # Family D
    from DDPM_clean_application.src import unet_ms
    for name, case in D_CASES.items():
        if args.only and args.only != name:
            continue
        net = unet_ms.UNet_downscale(c_in=case["c_in"], c_out=1, time_dim=256, interp_mode="bicubic",
                                     img_size=case["hw"], device="cpu")
        net.load_state_dict(synth.synth_state_dict_d(case["c_in"], 1, seed=case["wseed"]), strict=True)
        net.eval()
        inp = synth.synth_inputs(case["batch"], case["hw"], seed=case["iseed"], lowres=case["lowres"])
        out = {}
        with torch.no_grad():
            for t in case["ts"]:
                tt = torch.full((case["batch"],), t, dtype=torch.long)
                out[f"eps_t{t}"] = net(inp["x"], tt, inp["y_lowres"]).numpy()
        np.savez_compressed(os.path.join(HERE, f"d_{name}.npz"), **out)
        print("D", name, {k: (v.shape, float(np.abs(v).mean())) for k, v in out.items()})

    # sampling loop (v1 DiffusionUtils.sample) with injected noise
    import diffusion_DANRA_conditional as dref
    dref.tqdm.tqdm = lambda it, *a, **k: it   # silence the progress bar
    for name, sc in SAMPLE_CASES.items():
        if args.only and args.only != name:
            continue
        if sc.get("long") and not args.long:
            continue
        case = R_CASES[sc["model"]]
        net = build_ref_r(case)
        B, H, T = sc["batch"], case["hw"], sc["T"]
        inp = synth.synth_inputs(B, H, seed=case["iseed"], has_lsm=case["has_lsm"], has_topo=case["has_topo"],
                                 has_cond=case["has_cond"], num_classes=case["num_classes"])
        z = synth.step_noise(B, 1, H, T, seed=sc["zseed"])
        du = dref.DiffusionUtils(T, 1e-4, 0.02, "cpu", "linear")
        counter = {"i": T - 1}
        real = torch.randn_like

        def fake_randn_like(x, *a, **k):
            i = counter["i"]
            counter["i"] -= 1
            return z[i].clone()

        torch.randn_like = fake_randn_like
        try:
            x0 = du.sample(inp["x"].clone(), net, inp["y"], inp["cond"], inp["lsm"], inp["topo"])
        finally:
            torch.randn_like = real
        assert counter["i"] == 1, counter
        np.savez_compressed(os.path.join(HERE, f"sample_{name}.npz"), x0=x0.numpy(),
                            betas=du.betas.numpy(), alpha_hat=du.alpha_hat.numpy())
        print("S", name, x0.shape, float(x0.std()), float(x0.abs().max()))


if __name__ == "__main__":
    main()
