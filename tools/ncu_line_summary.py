"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line: executed warp
instructions and stall samples (top N lines)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ex = collections.Counter(); sm = collections.Counter(); txt = {}
hdr = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iX, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= iX: continue
    try: line = int(r[0]); x = int(r[iX] or 0); s = int(r[iS] or 0)
    except ValueError: continue
    if r[2] == "":      # a CUDA source line row (no SASS address): remember its text
        txt.setdefault(line, r[1].strip())
        continue
    ex[line] += x; sm[line] += s
    txt.setdefault(line, "")
tx, ts = sum(ex.values()), sum(sm.values())
print("total warp-instructions", tx, "samples", ts)
for line, x in ex.most_common(top):
    print("%5d %6.2f%% instr %6.2f%% samples  %s" % (line, 100.0 * x / max(tx, 1), 100.0 * sm[line] / max(ts, 1), txt.get(line, "")[:110]))
