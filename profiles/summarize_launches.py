"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import csv, re, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict(); tot = 0.0; n = 0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]); name = re.sub(r"sb::|cub::CUB_\w+::|detail::|radix::|void ", "", name)[:64]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; n += 1
print(f"total {tot:.3f} ms over {n} launches")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v:8.3f} ms {100 * v / tot:5.1f}%  x{c:3d}  {k}")
