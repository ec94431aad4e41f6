"""Top stall sites of one kernel from an .ncu-rep source page (SASS view): python tools/ncu_hot.py file.ncu-rep [kernel-substring] [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": line.split(",", 1)[1].strip('"'), "lines": []}
        blocks.append(cur)
    elif cur is not None:
        cur["lines"].append(line)
for b in blocks:
    if sub not in b["name"]:
        continue
    rows = list(csv.reader(io.StringIO("\n".join(b["lines"]))))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    data = rows[1:]
    tot = sum(int(r[col["# Samples"]] or 0) for r in data)
    print("====", b["name"][:120], "samples", tot, "instructions", len(data))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in stall_cols}
    print("  by reason:", {k[6:]: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 50 > tot})
    ranked = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
    for i in sorted(ranked):
        r = data[i]
        s = int(r[col["# Samples"]] or 0)
        why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"  {i:5d} {100 * s / max(tot, 1):5.1f}%  {r[col['Source']].strip()[:70]:70s} {why}")
