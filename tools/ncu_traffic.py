"""Per kernel class: mean DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) and mean duration from an ncu
raw CSV (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file x.csv <cmd>` or
`ncu -i rep --page raw --csv`).  Writes/updates profiles/r2_ncu_traffic.json[workload] = {class: bytes per launch}, which bench.py
reports as roofline.traffic.
    python tools/ncu_traffic.py <csv> <workload> [--out profiles/r2_ncu_traffic.json]"""
import csv
import json
import os
import re
import sys

CLASS_OF = [("attn_tc5", "attn_tc32"), ("attn_tc", "attn_tc"), ("conv_tc_kernel", "conv_tc"), ("conv_tcp_kernel", "conv_tc"), ("gemm_stream_kernel", "gemm_stream"), ("norm_fused_kernel", "norm_fused"),
            ("attn_block_kernel", "attn_block"), ("flash_attn_kernel", "flash_attn"), ("attn_wide_kernel", "flash_attn"),
            ("tail_tc_kernel", "tail_conv"), ("tail_mma_kernel", "tail_conv"), ("tail_conv_kernel", "tail_conv"), ("stem_mma_kernel", "stem_conv"),
            ("stem_conv_kernel", "stem_conv"), ("plane_stats_kernel", "plane_stats"), ("temb_project_kernel", "temb_project"),
            ("groupnorm_apply_kernel", "groupnorm_apply"), ("sample_stats_kernel", "sample_stats"), ("maxpool2_kernel", "maxpool"),
            ("upsample_cat_kernel", "upsample_cat"), ("outc_kernel", "outc"), ("layernorm_rows_kernel", "layernorm"),
            ("posterior_update_kernel", "posterior_update"), ("instnorm_apply_kernel", "instnorm_apply")]


def main():
    path, workload = sys.argv[1], sys.argv[2]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "profiles/r2_ncu_traffic.json"
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    agg = {}
    if "Metric Name" in hdr:      # long format of --log-file
        iname, imet, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        iid = hdr.index("ID")
        per = {}
        for r in rows[1:]:
            d = per.setdefault(r[iid], {"name": r[iname]})
            v = float(r[ival].replace(",", ""))
            u = r[iunit].lower()
            v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)
            d[r[imet]] = v
        items = per.values()
    else:
        units = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        items = []
        for r in rows[2:]:
            d = {"name": r[ix["Kernel Name"]]}
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                if m in ix and r[ix[m]] != "":
                    v = float(r[ix[m]].replace(",", ""))
                    u = units[ix[m]].lower()
                    v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1)
                    d[m] = v
            items.append(d)
    for d in items:
        cls = next((c for pat, c in CLASS_OF if pat in d["name"]), None)
        if cls is None:
            continue
        a = agg.setdefault(cls, dict(n=0, bytes=0.0, ns=0.0))
        a["n"] += 1
        a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
    res = json.load(open(out)) if os.path.exists(out) else {}
    res[workload] = {c: a["bytes"] / a["n"] for c, a in agg.items()}
    res.setdefault("_launches", {})[workload] = {c: a["n"] for c, a in agg.items()}
    res.setdefault("_mean_us_under_ncu", {})[workload] = {c: a["ns"] / a["n"] / 1e3 for c, a in agg.items()}
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    for c, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"{c:18s} n={a['n']:3d}  {a['bytes']/a['n']/1e6:10.3f} MB/launch  {a['ns']/a['n']/1e3:9.2f} us/launch")


if __name__ == "__main__":
    main()
