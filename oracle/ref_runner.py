"""Run the UNMODIFIED reference (staged under oracle/_ref by oracle/stage_ref.py, or /root/reference in the build container)
on the host cores.  TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py's ``cpu_baseline`` / ``--impl reference`` legs,
tests/golden/make_golden.py and the tests — never by the product package.

The reference model is built with the reference's own constructors, the seeded synthetic ``state_dict`` is strict-loaded, and
the reference's own ``DiffusionUtils.sample`` drives it (stock code path: its python-loop ``SinusoidalEmbedding``, its
``nn.MultiheadAttention`` with materialised weights, its per-step host tensor for ``t``)."""
from __future__ import annotations

import importlib
import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")


def ref_root():
    if os.path.exists(os.path.join(STAGED, "MANIFEST.json")):
        return STAGED
    if os.path.isdir("/root/reference"):
        return "/root/reference"
    return None


def _ensure_paths(root):
    for p in (os.path.join(root, "DDPM_DANRA_conditional"), os.path.join(root, "DDPM_DANRA_Downscaling"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "matplotlib" not in sys.modules:      # modules_DANRA_downscaling imports pyplot at the top; absent in this image
        try:
            import matplotlib  # noqa: F401
        except Exception:
            m, mp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            m.pyplot = mp
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, mp


def import_ref(name):
    root = ref_root()
    if root is None:
        raise RuntimeError("reference sources not available: run `python oracle/stage_ref.py` in the build container")
    _ensure_paths(root)
    return importlib.import_module(name)


def build_ref_r(case, synth, module_name="modules_DANRA_conditional"):
    """Family R through the reference's constructors (ddpm_DANRA_conditional_wValid__128x128.py:332-345 for cfg3)."""
    H = case["hw"]
    if case.get("clean"):
        mod = import_ref("DDPM_clean_application.src.unet")
        enc = mod.Encoder(1, 256, cond_on_lsm=case["has_lsm"], cond_on_topo=case["has_topo"], cond_on_img=case["has_cond"],
                          cond_img_dim=(1, H, H) if case["has_cond"] else None, num_classes=case["num_classes"],
                          n_heads=case.get("n_heads", 4))
        dec = mod.Decoder(512, 1, 256, 64, n_heads=case.get("n_heads", 4))
        net = mod.DiffusionNet(enc, dec)
        sd = synth.synth_state_dict_r(case["c_in"], 1, case["num_classes"], (H, H), case["has_lsm"], case["has_topo"],
                                      seed=case["wseed"], randomize_bn=case["randomize_bn"], clean=True)
    elif case.get("downscaling"):
        mod = import_ref("modules_DANRA_downscaling")
        net = mod.DiffusionNet(mod.Encoder(1, 256, n_heads=case.get("n_heads", 4)), mod.Decoder(512, 1, 256, 64, n_heads=case.get("n_heads", 4)))
        sd = synth.synth_state_dict_r(1, 1, None, (H, H), False, False, seed=case["wseed"], randomize_bn=case["randomize_bn"])
    else:
        mod = import_ref(case.get("module", module_name))
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


def build_ref_d(case, synth):
    mod = import_ref("DDPM_clean_application.src.unet_ms")
    net = mod.UNet_downscale(c_in=case["c_in"], c_out=1, time_dim=256, interp_mode=case.get("interp_mode", "bicubic"),
                             img_size=case["hw"], device="cpu")
    net.load_state_dict(synth.synth_state_dict_d(case["c_in"], 1, seed=case["wseed"]), strict=True)
    net.eval()
    return net


def time_reference_steps(case_name, batch, rev_steps, repeats, warmup, threads, T=1000):
    """Times `repeats` calls (after `warmup` untimed ones) of the reference's own DiffusionUtils.sample, each bounded to
    `rev_steps` reverse steps: the tables are the full T-step ones, only ``n_timesteps`` (the loop bound, diffusion_DANRA_
    conditional.py:127) is lowered, so every timed step executes exactly the code a full T-step job would.  Returns
    (samples/s extrapolated to T-1 steps, seconds per call list)."""
    from diffusionmodelscustom_b200 import synth
    from diffusionmodelscustom_b200.configs import D_CASES, R_CASES
    torch.set_num_threads(threads)
    dref = import_ref("diffusion_DANRA_conditional")
    dref.tqdm.tqdm = lambda it, *a, **k: it          # progress bar off
    du = dref.DiffusionUtils(T, 1e-4, 0.02, "cpu", "linear")
    du.n_timesteps = rev_steps + 1
    if case_name in D_CASES:
        case = D_CASES[case_name]
        net = build_ref_d(case, synth)
        inp = synth.synth_inputs(batch, case["hw"], seed=case["iseed"], lowres=case["lowres"])
        # UNet_downscale.forward(x, t, y): the low-res field rides in DiffusionUtils.sample's `y` slot; the three extra
        # positional None's the v1 sampler passes are swallowed by a thin callable (the model itself is untouched)
        model = _PositionalAdapter(net)
        call = lambda: du.sample(inp["x"].clone(), model, inp["y_lowres"])
    else:
        case = R_CASES[case_name]
        net = build_ref_r(case, synth)
        inp = synth.synth_inputs(batch, case["hw"], seed=case["iseed"], has_lsm=case["has_lsm"], has_topo=case["has_topo"],
                                 has_cond=case["has_cond"], num_classes=case["num_classes"])
        call = lambda: du.sample(inp["x"].clone(), net, inp["y"], inp["cond"], inp["lsm"], inp["topo"])
    secs = []
    for k in range(warmup + repeats):
        t0 = time.perf_counter()
        out = call()
        dt = time.perf_counter() - t0
        if k >= warmup:
            secs.append(dt)
    assert torch.isfinite(out).all()
    t_rev = (sum(secs) / len(secs)) / rev_steps
    return batch / (t_rev * (T - 1)), secs


class _PositionalAdapter:
    def __init__(self, net):
        self.net = net

    def eval(self):
        self.net.eval()
        return self

    def __call__(self, x, t, y=None, *ignored):
        return self.net(x, t, y)
