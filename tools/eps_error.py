"""Per-step eps_hat relative L2 error of the native path against the committed reference goldens, printed per case and timestep
(the parity tests only assert the gate): used to compare kernel variants (B2D_ATTN_V, B2D_ATTN_POLY, ...) at model level.
    python tools/eps_error.py [case ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tests import gpu_util as G                      # noqa: E402
from tests.cases import R_CASES                       # noqa: E402
from tests.model_util import build_ours_r, inputs_r   # noqa: E402

names = sys.argv[1:] or list(R_CASES)
gd = os.path.join("tests", "golden")
allv = []
for name in names:
    case = R_CASES[name]
    gold = np.load(os.path.join(gd, f"r_{name}.npz"))
    net, _ = build_ours_r(case)
    inp, dev = inputs_r(case)
    errs = []
    for t in case["ts"]:
        tt = torch.full((case["batch"],), t, dtype=torch.long, device="cuda")
        eps = net(dev["x"] * case.get("x_scale", 1.0), tt, dev["y"], dev["cond"], dev["lsm"], dev["topo"])
        errs.append(G.rel_l2(eps, gold[f"eps_t{t}"]))
    allv += errs
    print(f"{name:28s} " + " ".join(f"t={t}:{e:.2e}" for t, e in zip(case["ts"], errs)))
print(f"max {max(allv):.2e}  mean {sum(allv) / len(allv):.2e}")
