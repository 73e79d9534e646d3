"""Profiling driver: N plain eps evaluations of a bench workload (one reverse step each, ~85 launches), so that ncu can
capture exactly one step (`-s <launches> -c <launches>`), plus the live per-launch CUDA-event table (b2d_profile_step).
    python tools/profile_step.py [--workload cfg2] [--batch 64] [--iters 3] [--perop gpurun_out/perop.json]"""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from bench import WORKLOADS                            # noqa: E402
from diffusionmodelscustom_b200.configs import D_CASES, R_CASES, build_ours_d, build_ours_r, inputs_d, inputs_r  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--perop", default="")
ap.add_argument("--sample-steps", type=int, default=0, help="also run a short reverse loop (posterior update kernel)")
a = ap.parse_args()
case_name = WORKLOADS[a.workload]["case"]
batch = a.batch or WORKLOADS[a.workload]["global_batch"]
if case_name in D_CASES:
    case = D_CASES[case_name]
    net, _ = build_ours_d(case)
    _, d = inputs_d(case, batch)
    d.update(cond=d.pop("y_lowres"), lsm=None, topo=None, y=None)
else:
    case = R_CASES[case_name]
    net, _ = build_ours_r(case)
    _, d = inputs_r(case, batch)
t = torch.full((batch,), 500, dtype=torch.long, device="cuda")
for _ in range(a.iters):
    eps = net(d["x"], t, d["y"], d["cond"], d["lsm"], d["topo"])
if a.sample_steps:
    from diffusionmodelscustom_b200 import DiffusionUtils
    du = DiffusionUtils(a.sample_steps + 1, 1e-4, 0.02, "cuda")
    du.sample(d["x"], net, d["y"], d["cond"], d["lsm"], d["topo"], seed=1)
torch.cuda.synchronize()
print("launches per step:", net.launch_count(), "finite:", bool(torch.isfinite(eps).all()))
if a.perop:
    prof = net.profile_step(d["x"], t.cpu(), d["y"], d["cond"], d["lsm"], d["topo"], reps=10)
    json.dump(prof, open(a.perop, "w"), indent=0)
    tot = sum(p["ms"] for p in prof)
    for p in prof:
        tf = p["flops"] / p["ms"] / 1e9 if p["flops"] else 0
        print(f'{p["name"]:14s} {p["klass"]:15s} {p["ms"]*1e3:8.1f} us  {tf:8.1f} TF/s  {p["bytes"]/p["ms"]/1e6:8.1f} GB/s')
    print("total", tot, "ms")
