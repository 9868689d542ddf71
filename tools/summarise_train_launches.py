"""Aggregates an ncu launch list (gpu__time_duration.sum CSV) of tools/train_bench.py --steps 1 --warmup 1 by kernel:
takes the launches of the LAST step. python tools/summarise_train_launches.py gpurun_out/train_launches.csv"""
import collections
import csv
import re
import sys


def main(path, top=32):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            rows.append((row["Kernel Name"], v / 1000.0 if row["Metric Unit"] == "ns" else v))
    step = rows[len(rows) - len(rows) // 2:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in step:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"void |ishara::|\(anonymous namespace\)::|unnamed>::", "", k)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"launches {len(step)}  total {tot:.1f} us")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:3d}  avg {v[1] / v[0]:7.1f}  {k[:90]}")


if __name__ == "__main__":
    main(sys.argv[1])
