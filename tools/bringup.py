"""GPU bring-up report (not a test): runs each kernel class and the model with per-layer taps, printing errors instead of
asserting, so that one gpurun call localises every problem.  Usage on the GPU box: python tools/bringup.py"""
import math
import sys
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from diffusionmodelscustom_b200 import _native as N   # noqa: E402
from oracle import ddpm_oracle as O                   # noqa: E402
from tests import gpu_util as G                       # noqa: E402
from diffusionmodelscustom_b200.configs import R_CASES                       # noqa: E402
from diffusionmodelscustom_b200.configs import build_ours_r, inputs_r   # noqa: E402


def bf(x):
    return x.to(torch.float16).float()


def try_(name, fn):
    try:
        r = fn()
        print(f"[bringup] {name}: {r}", flush=True)
    except Exception as e:   # noqa
        print(f"[bringup] {name}: EXC {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def conv_case(B, H, Cin, Cout, R, stride, pad, impl):
    g = torch.Generator().manual_seed(1)
    x = bf(torch.randn(B, Cin, H, H, generator=g))
    w = bf(torch.randn(Cout, Cin, R, R, generator=g) / math.sqrt(Cin * R * R))
    bias = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, bias, stride, pad)
    out = G.conv2d(G.nhwc_f16(x), G.pack_conv_weight(w), bias.cuda(), None, None, B, H, H, Cin, Cout, R, stride, pad, impl=impl)
    return G.rel_l2(G.to_nchw_f32(out), ref)


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    for impl in (1, 0):
        tag = "simt" if impl else "tc"
        try_(f"conv {tag} 1x1 c64->64 16px", lambda: conv_case(2, 16, 64, 64, 1, 1, 0, impl))
        try_(f"conv {tag} 1x1 c128->128 16px", lambda: conv_case(2, 16, 128, 128, 1, 1, 0, impl))
        try_(f"conv {tag} 3x3 s1 c64 32px", lambda: conv_case(2, 32, 64, 64, 3, 1, 1, impl))
        try_(f"conv {tag} 3x3 s2 c64->128 16px", lambda: conv_case(2, 16, 64, 128, 3, 2, 1, impl))
        try_(f"conv {tag} 8x8 s2 c64 32px", lambda: conv_case(2, 32, 64, 64, 8, 2, 3, impl))
        try_(f"conv {tag} 3x3 s1 c512 2px", lambda: conv_case(5, 2, 512, 512, 3, 1, 1, impl))
    for name in ("cfg2_lsmtopo_64",):
        case = R_CASES[name]
        for simt in (True, False):
            def model():
                net, sd = build_ours_r(case)
                net.debug_simt_conv = simt
                inp, dev = inputs_r(case)
                t = torch.full((case["batch"],), 500, dtype=torch.long)
                taps = {}
                ref = O.family_r_forward(sd, inp["x"], t, inp["y"], inp["cond"], inp["lsm"], inp["topo"], taps=taps)
                eps = net(dev["x"], t.cuda(), dev["y"], dev["cond"], dev["lsm"], dev["topo"])
                torch.cuda.synchronize()
                rep = {}
                for k in ("fmap1", "fmap2", "fmap3", "fmap4", "fmap5", "dec0", "dec1", "dec2", "dec3"):
                    rep[k] = round(G.rel_l2(net.debug_read(k, case["batch"]), taps[k]), 5)
                rep["eps"] = round(G.rel_l2(eps, ref), 5)
                return rep
            try_(f"model {name} simt={simt}", model)


def family_d():
    from diffusionmodelscustom_b200.configs import D_CASES
    from diffusionmodelscustom_b200.configs import build_ours_d, inputs_d
    case = D_CASES["cfg4_downscale_64"]
    for simt in (True, False):
        def model():
            net, sd = build_ours_d(case)
            net.debug_simt_conv = simt
            inp, dev = inputs_d(case)
            t = torch.full((case["batch"],), 300, dtype=torch.long)
            taps = {}
            ref = O.family_d_forward(sd, inp["x"], t, inp["y_lowres"], taps=taps)
            eps = net(dev["x"], t.cuda(), dev["y_lowres"])
            torch.cuda.synchronize()
            rep = {k: round(G.rel_l2(net.debug_read(k, case["batch"]), taps[k]), 5) for k in ("x1", "x2", "x3", "bot", "u1", "u2", "u3")}
            rep["eps"] = round(G.rel_l2(eps, ref), 5)
            return rep
        try_(f"family D cfg4 simt={simt}", model)


if __name__ == "__main__":
    if "--d" in sys.argv:
        family_d()
        sys.exit(0)
    main()
