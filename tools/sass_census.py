"""Per-kernel census of the tensor-core / TMA / TMEM instructions in libishara_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA load / store), HMMA (legacy mma.sync),
LDGSTS (cp.async). usage: python tools/sass_census.py [path/to/lib.so] > profiles/r02_sass_census.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ishara_b200", "lib",
                                                        "libishara_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "UCGABAR"]
per = OrderedDict()
cur = None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it, m.group(1))
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"^void ", "", cur)
        cur = re.sub(r">\(.*$", ">", cur) if ">(" in cur else re.sub(r"\(.*$", "", cur)   # drop the parameter list, keep template arguments
        cur = cur.replace("(int)", "").replace("(bool)", "")[:110]
        per.setdefault(cur, Counter())
        continue
    if cur is None:
        continue
    mm = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not mm:
        continue
    op = mm.group(1)
    for k in KEYS:
        if op == k or op.startswith(k + "."):
            per[cur][k] += 1
    if op.startswith("UTCHMMA") and ".2CTA" in op:
        per[cur]["UTCHMMA.2CTA"] += 1
print(f"# SASS census of {os.path.basename(lib)} ({os.path.getsize(lib)} bytes): instruction counts per kernel (static, cuobjdump -sass)")
print(f"# {'kernel':110s} " + " ".join(f"{k:>12s}" for k in KEYS))
tot = Counter()
for k, c in per.items():
    if sum(c.values()) == 0:
        continue
    tot.update(c)
    print(f"{k:112s} " + " ".join(f"{c[x]:12d}" for x in KEYS))
print(f"{'TOTAL':112s} " + " ".join(f"{tot[x]:12d}" for x in KEYS))
