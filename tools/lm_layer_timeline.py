"""Per-position kernel times of one LM train step from an ncu launch list (tools/profile_lm_step.sh):
the forward and backward passes repeat one launch sequence per layer; prints the median time of every position in it.

    python tools/lm_layer_timeline.py gpurun_out/lm_launches.csv <launches_per_step> [n_layers]"""
import csv, re, statistics, sys

path, per_step = sys.argv[1], int(sys.argv[2])
L = int(sys.argv[3]) if len(sys.argv) > 3 else 24
lines = [l for l in open(path) if not l.startswith("==")]
by_id = {}
for row in csv.DictReader(lines):
    i = int(row["ID"])
    d = by_id.setdefault(i, {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", ""),
                             "grid": row.get("Grid Size", "")})
    v = float(row["Metric Value"].replace(",", ""))
    if row["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row.get("Metric Unit", "ns"), 1e-3)
    elif row["Metric Name"].startswith("dram__bytes"):
        d["mb"] = d.get("mb", 0.0) + v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(row.get("Metric Unit", "byte"), 1e-6)
allk = [by_id[i] for i in sorted(by_id)]
start = max(i for i, k in enumerate(allk) if k["name"].startswith("embed_fwd"))   # the last step starts at its embedding gather
seq = allk[start:]
names = [s["name"] for s in seq]
i_ce = max(i for i, n in enumerate(names) if n.startswith("ce_"))
fwd, bwd = seq[:i_ce - 2], seq[i_ce + 1:]          # forward ends with final rmsnorm + lm_head GEMM before the CE kernels
print(f"step: {sum(s['us'] for s in seq) / 1e3:.3f} ms over {len(seq)} launches (forward {sum(s['us'] for s in seq[:i_ce + 1]) / 1e3:.3f} ms)")


def periodic(part, label, skip_head):
    body = part[skip_head:]
    per = len(body) // L
    body = body[:per * L]
    print(f"--- {label}: {per} launches per layer, {sum(s['us'] for s in body) / 1e3:.3f} ms in {L} layers")
    for k in range(per):
        col = [body[l * per + k] for l in range(L)]
        assert len({c["name"] for c in col}) == 1, (k, {c["name"] for c in col})
        us = statistics.median(c["us"] for c in col)
        mb = statistics.median(c.get("mb", 0.0) for c in col)
        print(f"  {k:2d} {us:8.1f} us {mb:8.1f} MB  x{L}  {col[0]['name'][:70]}  grid {col[0]['grid']}")
    return per


periodic(fwd, "forward", 1)                        # launch 0 = embedding gather
# backward: lm_head dgrad / wgrad(+reduce) / final rmsnorm bwd (+ colsum reduce) precede the layers; find the period
for head in range(2, 12):
    rest = bwd[head:]
    tail = len(rest) % L
    per = len(rest) // L
    if per and all(len({rest[l * per + k]["name"] for l in range(L)}) == 1 for k in range(per)):
        print("backward head:", [(s["name"][:30], round(s["us"], 1)) for s in bwd[:head]])
        periodic(bwd, "backward", head)
        print("backward tail:", [(s["name"][:30], round(s["us"], 1)) for s in rest[per * L:]])
        break
