#!/usr/bin/env python
"""Headline benchmark: samples/s for full T=1000 DDPM sampling (999 UNet evaluations + posterior updates per sample).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is ONE pass of the hot path over one batch: DiffusionUtils.sample of B samples per GPU (T-1 graph replays).
Workload at N=1 = BASELINE.json configs[1] (cfg2: Family R, LSM+topography conditioning, 64x64, batch 64, T=1000);
N>1 shards independent samples (weak scaling: per-GPU batch fixed) with no data-path collective and one NCCL all_gather
of the final fields per step.

`value`  : device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
`e2e`    : same metric through the reference-facing API with HOST buffers (pinned): H2D of x_T + conditioning and D2H of
           the fields inside the timed region (DiffusionUtils.sample_host -> b2d_sample_host).
`roofline`: dominant kernel class of one reverse step, timed live with CUDA events between launches (b2d_profile_step).
`cpu_baseline`: the CPU oracle port (oracle/ddpm_oracle.py, FP32 torch on the host cores) on a bounded sample.
`--impl reference`: the reference's CPU implementation of the path = that same oracle port (the reference is Python and
           cannot travel to the GPU box; the port is pinned to it by tests/golden), all host threads, bounded steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 1000
WORKLOADS = {
    # name: (case in tests/cases.py, per-GPU batch, FLOPs per sample per reverse step [BASELINE.md §3])
    "cfg1": ("cfg1_uncond_64", 4, 1.381e9),
    "cfg2": ("cfg2_lsmtopo_64", 64, 1.398e9),
    "cfg3": ("cfg3_full_128", 32, 12.527e9),
    "cfg4": ("cfg4_downscale_64", 64, 17.75e9),      # Family D (UNet_downscale), low-res field 8x8 bicubic-upsampled
}


# DRAM bytes per launch (dram__bytes via ncu, cold-cache, averaged over the launches of the class) for the cfg2 step at batch 64:
# profiles/r1_ncu_step_cfg2_b64_final_sections.txt.  Reported as roofline.traffic for that workload only.
NCU_TRAFFIC_CFG2 = {"conv_tc": 2.91e6, "attn_tc": 15.76e6, "gemm_stream": 7.42e6, "norm_fused": 5.01e6, "attn_block": 1.28e6,
                    "tail_conv": 33.6e6, "stem_conv": 17.9e6, "plane_stats": 33.6e6, "temb_project": 1.65e6}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_throughput(case_name, batch, steps, warmup, threads):
    """Oracle port (FP32 torch on CPU) of one reverse step, timed; samples/s = batch / (t_step * 999)."""
    from diffusionmodelscustom_b200 import synth
    from oracle import ddpm_oracle as O
    from tests.cases import D_CASES, R_CASES
    torch.set_num_threads(threads)
    if case_name in D_CASES:
        case = D_CASES[case_name]
        sd = synth.synth_state_dict_d(case["c_in"], 1, seed=case["wseed"])
        inp = synth.synth_inputs(batch, case["hw"], seed=case["iseed"], lowres=case["lowres"])
        model_fn = lambda x, t: O.family_d_forward(sd, x, t, inp["y_lowres"])
    else:
        case = R_CASES[case_name]
        sd = synth.synth_state_dict_r(case["c_in"], 1, case["num_classes"], (case["hw"],) * 2, case["has_lsm"],
                                      case["has_topo"], seed=case["wseed"])
        inp = synth.synth_inputs(batch, case["hw"], seed=case["iseed"], has_lsm=case["has_lsm"], has_topo=case["has_topo"],
                                 has_cond=case["has_cond"], num_classes=case["num_classes"])
        model_fn = lambda x, t: O.family_r_forward(sd, x, t, inp["y"], inp["cond"], inp["lsm"], inp["topo"])
    betas, alphas, ahat = O.schedule_tables(T_STEPS, 1e-4, 0.02)
    x = inp["x"].clone()
    g = torch.Generator().manual_seed(1)
    times = []
    with torch.no_grad():
        for k in range(warmup + steps):
            i = T_STEPS - 1 - k
            t0 = time.perf_counter()
            t = torch.full((batch,), i, dtype=torch.long)
            eps = model_fn(x, t)
            x = O.posterior_update(x, eps, torch.randn(x.shape, generator=g), i, betas, alphas, ahat)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    t_step = sum(times) / len(times)
    return batch / (t_step * (T_STEPS - 1)), t_step


def run_reference(args, rank, world):
    if rank != 0:
        return
    case_name, _, _ = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    batch = 16          # bounded sample: large enough for the CPU GEMMs to thread well, small enough to finish in seconds
    t0 = time.perf_counter()
    sps, t_step = cpu_port_throughput(case_name, batch, max(args.steps, 1), max(args.warmup, 1), threads)
    sample = (f"{args.steps} timed reverse steps (after {max(args.warmup, 1)} warm-up) of the same network at batch {batch}, "
              f"extrapolated x999 steps; every step runs the identical graph")
    line = {"impl": "reference", "metric": "samples_per_s_T1000_ddpm_sampling", "value": sps, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3 * (T_STEPS - 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload_desc, "T": T_STEPS},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    case_name, batch, flops_per_sample_step = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from tests.cases import D_CASES, R_CASES
    is_d = case_name in D_CASES
    case = D_CASES[case_name] if is_d else R_CASES[case_name]
    if is_d:
        args.workload_desc = (f"{args.workload}: Family D UNet_downscale, {case['hw']}x{case['hw']}, c_in={case['c_in']} (x + "
                              f"{case['lowres']}x{case['lowres']} low-res field, bicubic), per-GPU batch {batch}, T={T_STEPS} linear beta")
    else:
        args.workload_desc = (f"{args.workload}: Family R DiffusionNet, {case['hw']}x{case['hw']}, c_in={case['c_in']} "
                              f"(lsm={case['has_lsm']}, topo={case['has_topo']}, cond={case['has_cond']}, "
                              f"classes={case['num_classes']}), per-GPU batch {batch}, T={T_STEPS} linear beta")
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    from diffusionmodelscustom_b200 import DiffusionUtils
    from tests.model_util import build_ours_d, build_ours_r, inputs_d, inputs_r

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if is_d:
        net, _ = build_ours_d(case, dev)
        host, d = inputs_d(dict(case, iseed=case["iseed"] + rank), batch, dev)
        for dd in (host, d):     # Family D: the low-res field travels in the cond_img slot of the sampler
            dd.update(cond=dd.pop("y_lowres"), lsm=None, topo=None, y=None)
    else:
        net, _ = build_ours_r(case, dev)
        host, d = inputs_r(dict(case, iseed=case["iseed"] + rank), batch, dev)
    du = DiffusionUtils(T_STEPS, 1e-4, 0.02, dev, "linear")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    gather = [torch.empty_like(d["x"]) for _ in range(world)] if world > 1 else None
    offset = rank * batch

    def step_device(k):
        x0 = du.sample(d["x"], net, d["y"], d["cond"], d["lsm"], d["topo"], seed=1234 + k, sample_offset=offset)
        if world > 1:
            dist.all_gather(gather, x0)
        return x0

    pinned = {k: (v.pin_memory() if v is not None else None) for k, v in host.items()}

    def step_host(k):
        return du.sample_host(pinned["x"], net, pinned["y"], pinned["cond"], pinned["lsm"], pinned["topo"], seed=1234 + k,
                              sample_offset=offset, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, clocks=None):
        for k in range(warmup):
            fn(k)
        barrier()
        if clocks:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        out = None
        for k in range(steps):
            flush.zero_()                    # L2 flush between timed iterations
            out = fn(warmup + k)
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        if world > 1:
            tmax = torch.tensor([ms], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax.item())
        return ms, wall, out

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, wall, x0 = timed(step_device, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if sampler else None
    launches = net.launch_count() * args.steps
    assert torch.isfinite(x0).all(), "non-finite samples"
    value = world * batch * args.steps / (ms / 1e3)
    # end-to-end through host buffers (max over ranks of wall-clock bracketed by barriers; includes H2D/D2H)
    ms_e2e, wall_e2e, x0h = timed(step_host, args.steps, 1)
    e2e_value = world * batch * args.steps / max(wall_e2e, ms_e2e / 1e3)
    if world > 1:
        t = torch.tensor([e2e_value], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        e2e_value = float(t.item())
    h2d = sum(v.numel() * v.element_size() for v in pinned.values() if v is not None)
    d2h = pinned["x"].numel() * 4

    if rank == 0:
        pk = peaks()
        # live per-kernel profile of one reverse step (same program, CUDA events between launches)
        tt = torch.full((batch,), 500, dtype=torch.long)
        prof = net.profile_step(d["x"], tt, d["y"], d["cond"], d["lsm"], d["topo"], reps=5)
        by = {}
        for p in prof:
            k = by.setdefault(p["klass"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            k["ms"] += p["ms"]; k["flops"] += p["flops"]; k["bytes"] += p["bytes"]; k["launches"] += 1
        step_ms = sum(k["ms"] for k in by.values())
        kernels = {}
        for name, k in sorted(by.items(), key=lambda kv: -kv[1]["ms"]):
            kernels[name] = {"share": round(k["ms"] / step_ms, 4), "ms_per_step": round(k["ms"], 5), "launches": k["launches"],
                             "tflops": round(k["flops"] / (k["ms"] * 1e-3) / 1e12, 3) if k["flops"] else None,
                             "gbs": round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1)}
        top = max(by.items(), key=lambda kv: kv[1]["ms"])
        tname, tk = top
        tensor_bound = tname in ("conv_tc", "flash_attn")
        if tensor_bound:
            ach = tk["flops"] / (tk["ms"] * 1e-3) / 1e12
            roof = {"kernel": tname, "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sustained"], "traffic": None, "peak_source": pk["src"] + " sustained bf16 (same rate as fp16)"}
        else:
            ach = tk["bytes"] / (tk["ms"] * 1e-3) / 1e9
            roof = {"kernel": tname, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": ach / pk["hbm"], "traffic": None, "peak_source": pk["src"]}
        roof["avg_launch_ms"] = tk["ms"] / tk["launches"]
        roof["launches_per_step"] = tk["launches"]
        roof["algorithmic_bytes_per_launch"] = tk["bytes"] / tk["launches"]
        if args.workload == "cfg2" and batch == 64:
            roof["traffic"] = NCU_TRAFFIC_CFG2.get(tname)
            roof["traffic_source"] = "profiles/r1_ncu_step_cfg2_b64_final_sections.txt (ncu dram bytes per launch, cold cache)"
        whole = value * flops_per_sample_step * (T_STEPS - 1) / 1e12 / world
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sps, t_step = cpu_port_throughput(case_name, 4, 3, 1, threads)
            cpu = {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                   "sample": "3 timed reverse steps (1 warm-up) of the same network at batch 4 on the host cores, x999"}
        line = {"metric": "samples_per_s_T1000_ddpm_sampling", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp16 operands, fp32 accumulate (fp32 state/update)",
                "data": "synthetic",
                "config": {"workload": args.workload_desc, "T": T_STEPS, "unet_evals_per_sample": T_STEPS - 1,
                           "global_batch": world * batch, "parallelism": f"sample-sharded x{world}, no per-step collective",
                           "l2": "256 MiB buffer written between timed steps", "rng": "in-kernel Philox4x32-10"},
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roof, "kernels": kernels,
                "whole_step": {"tflops_per_gpu": whole, "frac_of_sustained_peak": whole / pk["tf_sustained"],
                               "reverse_step_ms_graph": ms / args.steps / (T_STEPS - 1),
                               "reverse_step_ms_sum_of_kernels": step_ms},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
