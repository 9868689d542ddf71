"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for ONE step
(the launches between two consecutive input casts). usage: summarise_launches.py launches.csv [step_index]"""
import collections
import csv
import re
import sys


def main(path, step=3):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    # a step starts at the input cast; with sub-batch lanes there is one cast per lane, back to back
    starts = [i for i, n in enumerate(names) if "cast_pad" in n and (i == 0 or "cast_pad" not in names[i - 1])]
    s, e = starts[step], starts[step + 1]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[s:e]:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("ishara::<unnamed>::", "").replace("void ", "")
        d = float(r["Metric Value"]) / 1e3
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += d
        tot += d
    print(f"# one step = launches [{s},{e}) of {path}; gpu__time_duration.sum per launch (cold-cache, serialised)")
    print("kernel,launches,total_us,share")
    for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"\"{n}\",{c},{d:.1f},{d / tot:.4f}")
    print(f"\"TOTAL\",{e - s},{tot:.1f},1.0")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
