"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one line per kernel class:
launches, mean duration, tensor-pipe %, SFU(XU) %, DRAM bytes and GB/s, L2 bytes, achieved occupancy, registers."""
import csv, re, subprocess, sys
from collections import OrderedDict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def col(r, name, default=0.0):
    i = ix.get(name)
    if i is None or r[i] == "": return default
    try: return float(r[i].replace(",", ""))
    except ValueError: return default
want = OrderedDict(dur="gpu__time_duration.sum", tensor="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                   xu="sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", dram_r="dram__bytes_read.sum", dram_w="dram__bytes_write.sum", dram_rate="dram__bytes.sum.per_second", l2hit="lts__t_sector_hit_rate.pct",
                   dram_pct="gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", l2="lts__t_bytes.sum",
                   occ="sm__warps_active.avg.pct_of_peak_sustained_active", regs="launch__registers_per_thread",
                   smthr="sm__throughput.avg.pct_of_peak_sustained_elapsed")
units = rows[1]
def to_us(r):
    v = col(r, want["dur"]); u = units[ix[want["dur"]]]
    return v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
def to_bytes(r, key):
    i = ix.get(want[key])
    if i is None: return 0.0
    v = col(r, want[key]); u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
agg = OrderedDict()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("b2d::", "").replace("void ", "")
    a = agg.setdefault(name, dict(n=0, us=0, tensor=0, xu=0, dr=0, dw=0, dpct=0, l2=0, occ=0, regs=0, smthr=0, grid=r[ix["Grid Size"]]))
    rate_i = ix.get(want["dram_rate"])
    if rate_i is not None and r[rate_i] != "":
        mult = {"byte/s": 1, "kbyte/s": 1e3, "mbyte/s": 1e6, "gbyte/s": 1e9, "tbyte/s": 1e12}.get(units[rate_i].lower(), 1)
        if want["dram_r"] not in ix:   # sections-only capture: bytes = rate x duration
            a["dr"] += col(r, want["dram_rate"]) * mult * to_us(r) * 1e-6
    a["l2hit"] = a.get("l2hit", 0) + col(r, want["l2hit"])
    a["n"] += 1; a["us"] += to_us(r); a["tensor"] += col(r, want["tensor"]); a["xu"] += col(r, want["xu"])
    a["dr"] += to_bytes(r, "dram_r"); a["dw"] += to_bytes(r, "dram_w"); a["dpct"] += col(r, want["dram_pct"]); a["l2"] += to_bytes(r, "l2")
    a["occ"] += col(r, want["occ"]); a["regs"] = col(r, want["regs"]); a["smthr"] += col(r, want["smthr"])
print(f"{'kernel':34s} {'n':>3s} {'us/launch':>9s} {'tensor%':>8s} {'xu%':>6s} {'sm_thr%':>8s} {'dram MB/l':>10s} {'dram GB/s':>10s} {'dram%':>6s} {'L2hit%':>8s} {'warps%':>7s} {'regs':>5s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = a["n"]; us = a["us"] / n; db = (a["dr"] + a["dw"]) / n
    print(f"{k[:34]:34s} {n:3d} {us:9.2f} {a['tensor']/n:8.2f} {a['xu']/n:6.2f} {a['smthr']/n:8.2f} {db/1e6:10.3f} {db/us/1e3:10.1f} {a['dpct']/n:6.2f} {a.get('l2hit',0)/n:8.2f} {a['occ']/n:7.2f} {a['regs']:5.0f}")
