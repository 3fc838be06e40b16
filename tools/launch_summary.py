"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("gps::<unnamed>::", "")[:80]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"{'total us':>10} {'share':>6} {'n':>4} {'avg us':>8}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v:10.1f} {100 * v / T:5.1f}% {cnt[k]:4d} {v / cnt[k]:8.1f}  {k}")
    print(f"{T:10.1f} 100.0% {sum(cnt.values()):4d}           (all launches in the window)")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
