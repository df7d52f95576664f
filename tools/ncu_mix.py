"""Executed-instruction mix of one kernel launch of an ncu report (needs --import-source on / -lineinfo not required).
usage: python tools/ncu_mix.py rep.ncu-rep [launch_index] [--regions]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True).stdout.decode("utf-8", "replace")
rows = list(csv.reader(io.StringIO(txt)))
b = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
hdr = rows[b[which] + 1]
print(rows[b[which]][1][:90])
data = [r for r in rows[b[which] + 2:b[which + 1]] if len(r) == len(hdr)]
ie = hdr.index("Instructions Executed"); si = hdr.index("Source"); ns = hdr.index("# Samples")
byop = collections.Counter(); bys = collections.Counter()
for r in data:
    op = [o for o in r[si].split() if not o.startswith("@")]
    byop[op[0].split(".")[0]] += int(r[ie]); bys[op[0].split(".")[0]] += int(r[ns])
tot = sum(byop.values()); ts = sum(bys.values())
print("warp instructions", tot, "samples", ts)
for k, v in byop.most_common(22):
    print("  %-8s %10d %5.1f%%   samples %5.1f%%" % (k, v, 100 * v / tot, 100 * bys[k] / max(ts, 1)))
if "--regions" in sys.argv:
    prev = None; start = 0; acc = 0; sacc = 0
    for j, r in enumerate(data + [None]):
        c = int(r[ie]) if r else None
        if c != prev:
            if prev is not None:
                print("  [%4d..%4d] x%-9d = %10d warp-inst (%4.1f%%) samples %4.1f%%  %s" % (start, j - 1, prev, acc, 100 * acc / tot, 100 * sacc / max(ts, 1), data[start][si].strip()[:50]))
            prev = c; start = j; acc = 0; sacc = 0
        if r: acc += c; sacc += int(r[ns])
for i, c in enumerate(hdr):
    if c.startswith("stall_") and "Not Issued" not in c:
        s = sum(int(r[i]) for r in data if r[i].isdigit())
        if s and 100 * s / ts > 2: print("  %s %.1f%%" % (c, 100 * s / ts))
