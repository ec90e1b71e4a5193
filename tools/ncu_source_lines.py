"""Per-source-line warp-stall summary of an .ncu-rep captured with --import-source on (first kernel in the report).

    python tools/ncu_source_lines.py gpurun_out/prof.ncu-rep [min_pct]

Prints, for every CUDA source line with >= min_pct of the stall samples: samples, share, instructions executed and the
dominant stall reasons (columns of `ncu --page source --print-source cuda,sass --csv`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
i_s, i_ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines = [r for r in rows if len(r) == len(hdr) and r[0] not in ("", "Line No")]
tot = sum(int(r[i_s] or 0) for r in lines) or 1
tot_inst = sum(int(r[i_ie] or 0) for r in lines)
print(f"{rep}: {tot} samples, {tot_inst} warp instructions over {len(lines)} source lines")
for r in lines:
    s = int(r[i_s] or 0)
    if 100.0 * s / tot < min_pct:
        continue
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
    sts = " ".join(f"{h}:{100 * v // max(s, 1)}%" for v, h in st if v)
    print(f"{r[0]:>5s} {s:7d} {100.0 * s / tot:5.1f}% inst {int(r[i_ie] or 0):9d} | {sts:42s} | {r[1].strip()[:110]}")
