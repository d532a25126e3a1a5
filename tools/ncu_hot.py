"""Top stall sites of an .ncu-rep at SASS level (read on the CPU box).  python tools/ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
tot = sum(int(r[col["# Samples"]] or 0) for r in data if len(r) > 5)
print("total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ranked = sorted((r for r in data if len(r) > 5), key=lambda r: -int(r[col["# Samples"]] or 0))[:N]
idx = {id(r): i for i, r in enumerate(data)}
for r in ranked:
    s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print("%5.1f%% #%4d %-70s exec=%-8s %s" % (100.0 * s / tot, idx[id(r)], r[col["Source"]][:70], r[col["Instructions Executed"]], " ".join("%s:%d" % (n, c) for c, n in top if c)))
