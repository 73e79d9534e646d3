"""Static SASS opcode counts per kernel of libb200ddpm.so -> profiles/r2_sass_opcodes.txt
    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess

LIB = "diffusionmodelscustom_b200/libb200ddpm.so"
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "BAR", "HMMA", "MUFU", "HFMA2", "VIMNMX3", "ELECT", "R2UR"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
counts, cur, order = {}, None, []
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", next(it)).replace("void ", "").replace("b2d::", "")
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
print("SASS opcode counts per kernel of libb200ddpm.so (cuobjdump -sass; static instruction counts; tools/sass_opcodes.py).")
print("UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,")
print("BAR = named / CTA barriers, HMMA = legacy mma.sync, MUFU = SFU, HFMA2 = packed-half FMA (polynomial exp2), ELECT = elect.sync.\n")
print(f"{'kernel':64s}" + "".join(f"{c:>9s}" for c in COLS))
for k in order:
    c = counts[k]
    print(f"{k[:64]:64s}" + "".join(f"{c[x]:9d}" for x in COLS))
