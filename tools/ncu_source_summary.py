"""Reduce `ncu -i rep --page source --csv` to executed warp instructions per opcode (optionally per unit of work) and the
hottest SASS lines by stall samples:  python tools/ncu_source_summary.py src.csv [units]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if "Source" in r)
h = rows[hi]
rows = rows[hi:]
src = h.index("Source")
ex = [i for i, n in enumerate(h) if n.startswith("# Warp Instructions Executed") or n == "Instructions Executed"][0]
ss = [i for i, n in enumerate(h) if n.startswith("# Samples") or "Warp Stall Sampling (All" in n][0]
agg, samp, lines = collections.Counter(), collections.Counter(), []
for r in rows[1:]:
    try:
        e, s = int(r[ex]), int(r[ss])
    except (ValueError, IndexError):
        continue
    toks = r[src].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
    op = op.split(".")[0]
    agg[op] += e
    samp[op] += s
    lines.append((s, e, r[src]))
tot, ts = sum(agg.values()), sum(samp.values())
print(f"executed warp instructions {tot}, stall samples {ts}; per unit = / {units:g}")
print(f"{'opcode':10s} {'executed':>12s} {'per unit':>9s} {'samples':>8s} {'share':>6s}")
for op, e in agg.most_common(30):
    print(f"{op:10s} {e:12d} {e / units:9.1f} {samp[op]:8d} {100 * samp[op] / max(ts, 1):5.1f}%")
print("\ntop SASS lines by stall samples:")
for s, e, t in sorted(lines, reverse=True)[:30]:
    print(f"{s:6d} {e:10d}  {t[:110]}")
