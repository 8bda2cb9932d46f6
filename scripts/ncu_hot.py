"""Top stalled SASS/source lines of an ncu report: python scripts/ncu_hot.py gpurun_out/prof_X.ncu-rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: v = float(r[ci["# Samples"]])
    except ValueError: continue
    top = sorted(((float(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    data.append((v, r[ci["Source"]][:95], ",".join("%s=%d" % (c[6:], x) for x, c in top if x > 0)))
tot = sum(d[0] for d in data)
print("total samples", tot)
for v, s, t in sorted(data, reverse=True)[:n]:
    print("%5.1f%%  %-95s %s" % (100 * v / tot, s, t))
