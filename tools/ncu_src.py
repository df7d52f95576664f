"""Summarise an `ncu --page source --csv --print-source sass` dump: samples by opcode, by stall reason, top instructions.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_src.py src.csv [launch_index]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# split into kernels: a kernel block starts with a "Kernel Name" row followed by a header row
blocks = []
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        blocks.append(i)
blocks.append(len(rows))
b0, b1 = blocks[which], blocks[which + 1]
print(rows[b0][1][:100])
hdr = rows[b0 + 1]
data = [r for r in rows[b0 + 2:b1] if len(r) == len(hdr)]
si = hdr.index("Source"); k = hdr.index("# Samples")
byop = collections.Counter(); tot = 0
for r in data:
    if r[k].isdigit():
        op = [o for o in r[si].split() if not o.startswith("@")]
        byop[(op[0] if op else "").split(".")[0]] += int(r[k]); tot += int(r[k])
print("samples", tot, "instructions", len(data))
print(byop.most_common(20))
for i, c in enumerate(hdr):
    if c.startswith("stall_") and "Not Issued" not in c:
        s = sum(int(r[i]) for r in data if r[i].isdigit())
        if s: print(f"  {c}: {s} ({100*s/tot:.1f}%)")
lsb = hdr.index("stall_long_sb")
print("top long_sb:")
for s, j, src in sorted(((int(r[lsb]) if r[lsb].isdigit() else 0, j, r[si]) for j, r in enumerate(data)), reverse=True)[:8]:
    print("   ", s, j, src)
print("top samples:")
for s, j, src in sorted(((int(r[k]) if r[k].isdigit() else 0, j, r[si]) for j, r in enumerate(data)), reverse=True)[:12]:
    print("   ", s, j, src)
