"""Summarise an `ncu --page source --csv` dump: executed instructions and stall samples per opcode."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
iS, iX, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
ops = collections.Counter(); smp = collections.Counter()
tot_x = tot_s = 0
lines = []
for r in rows[hdr_i + 1:]:
    if len(r) <= iX: continue
    try: x = int(r[iX]); s = int(r[iS])
    except ValueError: continue
    toks = r[iSrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    ops[op] += x; smp[op] += s; tot_x += x; tot_s += s
    lines.append((s, x, r[iSrc].strip()))
print("total warp-instructions executed", tot_x, "samples", tot_s)
print("%-10s %14s %7s %9s %7s" % ("opcode", "executed", "%", "samples", "%"))
for op, x in ops.most_common(28):
    print("%-10s %14d %6.1f%% %9d %6.1f%%" % (op, x, 100.0 * x / tot_x, smp[op], 100.0 * smp[op] / max(tot_s, 1)))
if len(sys.argv) > 2:
    print("--- top lines by samples")
    for s, x, src in sorted(lines, reverse=True)[: int(sys.argv[2])]:
        print("%7d %10d  %s" % (s, x, src))
