"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and share per kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += v * scale; cnt[name] += 1
total = sum(tot.values())
print(f"total {total:.3f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:9.3f} ms {100*v/total:5.1f}%  x{cnt[k]:5d}  {k[:110]}")
