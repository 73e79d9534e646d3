"""BASELINE.json configurations as concrete synthetic cases, and builders of the drop-in models with seeded synthetic weights
and DANRA-shaped inputs.  Shared by bench.py, __graft_entry__.smoke(), the profiling tools, the golden
generator and the tests (tests/cases.py and tests/model_util.py re-export from here).  The names map to BASELINE.json
configs (cfgN) where applicable."""
import torch

from . import synth

def _r(hw, batch, has_lsm, has_topo, has_cond, num_classes, ts=(999, 500, 1), wseed=42, iseed=7,
       randomize_bn=False, **kw):
    c_in = 1 + int(has_lsm) + int(has_topo) + int(has_cond)
    d = dict(hw=hw, batch=batch, has_lsm=has_lsm, has_topo=has_topo, has_cond=has_cond, num_classes=num_classes,
             c_in=c_in, ts=list(ts), wseed=wseed, iseed=iseed, randomize_bn=randomize_bn)
    d.update(kw)
    return d


R_CASES = {
    # cfg 1: unconditional, 64x64, c_in = 1
    "cfg1_uncond_64": _r(64, 2, False, False, False, None),
    # cfg 2: LSM + topography conditioning, 64x64, c_in = 3
    "cfg2_lsmtopo_64": _r(64, 2, True, True, False, None),
    # cfg 3: full conditioning + season classes at 128x128, c_in = 4
    "cfg3_full_128": _r(128, 1, True, True, True, 4, ts=(999, 1)),
    # cfg 5: same network through the modules_DANRA_flexible import path
    "cfg5_flexible_128": _r(128, 1, True, True, True, 4, ts=(700,), module="modules_DANRA_flexible"),
    # BN folding exercised with randomised running stats / affine; x scaled like late-trajectory states
    "full_64_randbn": _r(64, 3, True, True, True, 4, ts=(999, 250), randomize_bn=True, wseed=43, iseed=8),
    "full_64_bigx": _r(64, 2, True, True, True, 4, ts=(40,), x_scale=300.0, wseed=44, iseed=9),
    # smallest legal field (fmap5 is 1x1, attention over a single token), other head count
    "full_32_heads8": _r(32, 2, True, True, True, 4, ts=(321,), n_heads=8, wseed=45, iseed=10),
    # Downscaling generation (DDPM_DANRA_Downscaling): unconditional, encoder embeds t with the interleaved base-10000 embedding
    "downscaling_uncond_64": _r(64, 2, False, False, False, None, ts=(999, 321, 1), wseed=48, iseed=13, downscaling=True,
                                module="modules_DANRA_downscaling"),
    # launcher default of the clean application (test/launch.py:62): ONE attention head (head_dim = C = 64 ... 512)
    "cfg2_heads1_64": _r(64, 2, True, True, False, None, ts=(700,), n_heads=1, wseed=49, iseed=14),
    # newest generation (DDPM_clean_application/src/unet.py): attention with FF tail, cond_on_lsm/topo flags, 8 heads
    "clean_ff_64_heads8": _r(64, 2, True, True, True, 4, ts=(999, 77), n_heads=8, wseed=47, iseed=12, clean=True),
}

D_CASES = {
    # cfg 4: UNet_downscale 64x64, HR + bicubic-upsampled low-res field, c_in = 2
    "cfg4_downscale_64": dict(hw=64, batch=2, c_in=2, lowres=8, ts=[999, 300, 1], wseed=42, iseed=7),
    "downscale_32": dict(hw=32, batch=3, c_in=2, lowres=4, ts=[555], wseed=46, iseed=11),
    # the other F.interpolate modes of UNet_downscale(interp_mode=...) (unet_ms.py:105,156); non-integer scale factor 32/5
    "downscale_32_bilinear": dict(hw=32, batch=2, c_in=2, lowres=5, ts=[400], wseed=46, iseed=15, interp_mode="bilinear"),
    "downscale_32_nearest": dict(hw=32, batch=2, c_in=2, lowres=5, ts=[400], wseed=46, iseed=16, interp_mode="nearest"),
}

SAMPLE_CASES = {
    "cfg2_T50": dict(model="cfg2_lsmtopo_64", batch=2, T=50, zseed=1),
    "cfg1_T1000": dict(model="cfg1_uncond_64", batch=2, T=1000, zseed=1, long=True),
    # free-running at 128x128 (the headline resolution), full T=1000, and Family D over 199 reverse steps
    "cfg3_T1000": dict(model="cfg3_full_128", batch=1, T=1000, zseed=2, long=True),
    "cfg4_T200": dict(model="cfg4_downscale_64", family="D", batch=2, T=200, zseed=3, long=True),
    # schedule / sampler variants: v1 raised-cosine beta ramp (diffusion_DANRA_conditional.py:65-77); v2 sampler
    # (src/diffusion_modules.py:101-186) with the Nichol-Dhariwal cosine schedule and data_scaled (x0.005) noise
    "v1_cosine_T30": dict(model="cfg2_lsmtopo_64", batch=2, T=30, zseed=5, scheduler="cosine"),
    "v2_scaled_T40": dict(model="clean_ff_64_heads8", batch=2, T=40, zseed=6, v2=True, data_scaled=True),
    "v2_cosine_scaled_T40": dict(model="clean_ff_64_heads8", batch=2, T=40, zseed=4, v2=True, scheduler="cosine",
                                 data_scaled=True),
}


def _pkg():
    import diffusionmodelscustom_b200 as P
    return P


def build_ours_r(case, device="cuda"):
    H = case["hw"]
    if case.get("clean"):
        from . import unet as U
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
    if case.get("downscaling"):
        from . import downscaling as DS
        net = DS.DiffusionNet(DS.Encoder(1, 256, n_heads=case.get("n_heads", 4)), DS.Decoder(512, 1, 256, 64, n_heads=case.get("n_heads", 4)))
        sd = synth.synth_state_dict_r(1, 1, None, (H, H), False, False, seed=case["wseed"], randomize_bn=case["randomize_bn"])
        net.load_state_dict(sd, strict=True)
        net.eval()
        return net.to(device), sd
    z = torch.zeros(1, H, H)
    enc = _pkg().Encoder(1, 256, lsm_tensor=z if case["has_lsm"] else None, topo_tensor=z.clone() if case["has_topo"] else None,
                    cond_on_img=case["has_cond"], cond_img_dim=(1, H, H) if case["has_cond"] else None,
                    num_classes=case["num_classes"], n_heads=case.get("n_heads", 4))
    dec = _pkg().Decoder(512, 1, 256, 64, n_heads=case.get("n_heads", 4))
    net = _pkg().DiffusionNet(enc, dec)
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
    net = _pkg().UNet_downscale(c_in=case["c_in"], c_out=1, time_dim=256, interp_mode=case.get("interp_mode", "bicubic"),
                                img_size=case["hw"], device=device)
    sd = synth.synth_state_dict_d(case["c_in"], 1, seed=case["wseed"])
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net.to(device), sd


def inputs_d(case, batch=None, device="cuda"):
    B = batch or case["batch"]
    inp = synth.synth_inputs(B, case["hw"], seed=case["iseed"], lowres=case["lowres"])
    dev = {k: (v.to(device) if v is not None else None) for k, v in inp.items()}
    return inp, dev
