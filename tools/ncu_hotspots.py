"""Per-instruction hot spots of an ncu report's source page (needs -lineinfo / --import-source on):
    python tools/ncu_hotspots.py rep.ncu-rep [top_n]
prints the instructions with the most stall samples, their dominant stall reason, and the sample share of SASS regions
delimited by control-flow / barrier instructions."""
import csv, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        n = int(r[ci["# Samples"]] or 0)
    except ValueError:
        continue
    st = {h: int(r[ci[h]] or 0) for h in stalls}
    data.append((r[ci["Address"]], r[ci["Source"]].strip(), n, int(r[ci["Instructions Executed"]] or 0), st))
tot = sum(d[2] for d in data) or 1
print(f"total samples {tot}, instructions {len(data)}")
agg = {}
for d in data:
    for k, v in d[4].items():
        agg[k] = agg.get(k, 0) + v
print("stall totals: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
print("---- top instructions")
for idx, d in sorted(enumerate(data), key=lambda x: -x[1][2])[:top]:
    dom = max(d[4].items(), key=lambda x: x[1])
    print(f"{idx:5d} {100 * d[2] / tot:5.2f}% exec={d[3]:9d} {dom[0][6:]:>16s}  {d[1][:90]}")
print("---- cumulative by 64-instruction window")
for i in range(0, len(data), 64):
    s = sum(d[2] for d in data[i:i + 64])
    ex = sum(d[3] for d in data[i:i + 64])
    if s:
        print(f"{i:5d}-{i + 63:5d} {100 * s / tot:5.1f}%  exec {ex}")
