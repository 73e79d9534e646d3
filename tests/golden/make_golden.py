"""Generate golden fixtures by running the UNMODIFIED reference (imported from /root/reference through oracle/ref_runner.py).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py [--long] [--only NAME]

For every case the reference model is built with the reference's own constructors, our seeded
synthetic state_dict is loaded with ``strict=True`` (which also pins key names and shapes), the
reference forward / ``DiffusionUtils.sample`` is run on seeded synthetic inputs, and only the
OUTPUTS are stored (inputs and weights regenerate bit-identically from their seeds via
``diffusionmodelscustom_b200.synth``).  Noise for ``sample`` is injected by patching
``torch.randn_like`` (and ``torch.randn`` for the v2 sampler's own x_T draw) so that the reference consumes host-generated
x_T / z_i.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from diffusionmodelscustom_b200 import synth  # noqa: E402
from diffusionmodelscustom_b200.configs import D_CASES, R_CASES, SAMPLE_CASES  # noqa: E402
from oracle import ref_runner as RR  # noqa: E402

os.environ.setdefault("B2D_GOLDEN_FROM", "/root/reference")
RR.STAGED = "/nonexistent"          # goldens always come from /root/reference itself, never from a staged copy
torch.set_num_threads(8)


def run_r(case):
    net = RR.build_ref_r(case, synth)
    inp = synth.synth_inputs(case["batch"], case["hw"], seed=case["iseed"], has_lsm=case["has_lsm"],
                             has_topo=case["has_topo"], has_cond=case["has_cond"], num_classes=case["num_classes"])
    out = {}
    with torch.no_grad():
        for t in case["ts"]:
            tt = torch.full((case["batch"],), t, dtype=torch.long)
            x = inp["x"] * case.get("x_scale", 1.0)
            if case.get("downscaling"):
                out[f"eps_t{t}"] = net(x, tt).numpy()
            else:
                out[f"eps_t{t}"] = net(x, tt, inp["y"], inp["cond"], inp["lsm"], inp["topo"]).numpy()
    return out


def run_sample(name, sc):
    T, B = sc["T"], sc["batch"]
    sched = sc.get("scheduler", "linear")
    if sc.get("family") == "D":
        case = D_CASES[sc["model"]]
        net = RR._PositionalAdapter(RR.build_ref_d(case, synth))
        inp = synth.synth_inputs(B, case["hw"], seed=case["iseed"], lowres=case["lowres"])
        args = (inp["y_lowres"],)
    else:
        case = R_CASES[sc["model"]]
        net = RR.build_ref_r(case, synth)
        inp = synth.synth_inputs(B, case["hw"], seed=case["iseed"], has_lsm=case["has_lsm"], has_topo=case["has_topo"],
                                 has_cond=case["has_cond"], num_classes=case["num_classes"])
        args = (inp["y"], inp["cond"], inp["lsm"], inp["topo"])
    H = case["hw"]
    z = synth.step_noise(B, 1, H, T, seed=sc["zseed"])
    counter = {"i": T - 1}
    real_like, real_randn = torch.randn_like, torch.randn

    def fake_randn_like(x, *a, **k):
        i = counter["i"]
        counter["i"] -= 1
        return z[i].clone()

    torch.randn_like = fake_randn_like
    try:
        if sc.get("v2"):
            dmod = RR.import_ref("DDPM_clean_application.src.diffusion_modules")
            dmod.tqdm.tqdm = lambda it, *a, **k: it
            du = dmod.DiffusionUtils(T, 1e-4, 0.02, "cpu", sched, img_size=H, data_scaled=sc.get("data_scaled", False))
            torch.randn = lambda *a, **k: inp["x"].clone()          # the sampler's own x_T draw
            try:
                x0 = du.sample(B, net, 1, *args)
            finally:
                torch.randn = real_randn
        else:
            dref = RR.import_ref("diffusion_DANRA_conditional")
            dref.tqdm.tqdm = lambda it, *a, **k: it
            du = dref.DiffusionUtils(T, 1e-4, 0.02, "cpu", sched)
            x0 = du.sample(inp["x"].clone(), net, *args)
    finally:
        torch.randn_like = real_like
    assert counter["i"] == 1, counter
    np.savez_compressed(os.path.join(HERE, f"sample_{name}.npz"), x0=x0.numpy(), betas=du.betas.numpy(),
                        alpha_hat=du.alpha_hat.numpy())
    print("S", name, tuple(x0.shape), float(x0.std()), float(x0.abs().max()), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", action="store_true", help="also run the long free-running cases (minutes each)")
    ap.add_argument("--only", default=None)
    ap.add_argument("--samples-only", action="store_true")
    args = ap.parse_args()

    if not args.samples_only:
        for name, case in R_CASES.items():
            if args.only and args.only != name:
                continue
            out = run_r(case)
            np.savez_compressed(os.path.join(HERE, f"r_{name}.npz"), **out)
            print("R", name, {k: (v.shape, float(np.abs(v).mean())) for k, v in out.items()})
        for name, case in D_CASES.items():
            if args.only and args.only != name:
                continue
            net = RR.build_ref_d(case, synth)
            inp = synth.synth_inputs(case["batch"], case["hw"], seed=case["iseed"], lowres=case["lowres"])
            out = {}
            with torch.no_grad():
                for t in case["ts"]:
                    tt = torch.full((case["batch"],), t, dtype=torch.long)
                    out[f"eps_t{t}"] = net(inp["x"], tt, inp["y_lowres"]).numpy()
            np.savez_compressed(os.path.join(HERE, f"d_{name}.npz"), **out)
            print("D", name, {k: (v.shape, float(np.abs(v).mean())) for k, v in out.items()})

    for name, sc in SAMPLE_CASES.items():
        if args.only and args.only != name:
            continue
        if sc.get("long") and not args.long:
            continue
        run_sample(name, sc)


if __name__ == "__main__":
    main()
