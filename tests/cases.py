"""Parity cases shared by the golden generator, the CPU oracle tests and the GPU parity tests.
The names map to BASELINE.json configs (cfgN) where applicable."""

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
    # newest generation (DDPM_clean_application/src/unet.py): attention with FF tail, cond_on_lsm/topo flags, 8 heads
    "clean_ff_64_heads8": _r(64, 2, True, True, True, 4, ts=(999, 77), n_heads=8, wseed=47, iseed=12, clean=True),
}

D_CASES = {
    # cfg 4: UNet_downscale 64x64, HR + bicubic-upsampled low-res field, c_in = 2
    "cfg4_downscale_64": dict(hw=64, batch=2, c_in=2, lowres=8, ts=[999, 300, 1], wseed=42, iseed=7),
    "downscale_32": dict(hw=32, batch=3, c_in=2, lowres=4, ts=[555], wseed=46, iseed=11),
}

SAMPLE_CASES = {
    "cfg2_T50": dict(model="cfg2_lsmtopo_64", batch=2, T=50, zseed=1),
    "cfg1_T1000": dict(model="cfg1_uncond_64", batch=2, T=1000, zseed=1, long=True),
}
