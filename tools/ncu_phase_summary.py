"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump of one kernel: stall samples,
executed warp instructions and excess shared-memory wavefronts by source line and by line range (phase)."""
import csv, collections, sys
path, srcfile = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
hdr = None; nk = 0
ex = collections.Counter(); sm = collections.Counter(); cf = collections.Counter()
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1; continue
    if r and r[0] == "Line No":
        hdr = r; iX = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iE = hdr.index("L1 Wavefronts Shared Excessive"); continue
    if nk > 1: break
    if hdr is None or len(r) <= iX: continue
    try: line = int(r[0]); x = int(r[iX] or 0); s = int(r[iS] or 0); e = int(r[iE] or 0)
    except ValueError: continue
    if r[2] == "": continue
    ex[line] += x; sm[line] += s; cf[line] += e
src = open(srcfile).read().split("\n")
tx, ts, tc = sum(ex.values()), sum(sm.values()), sum(cf.values())
print("total warp-instructions", tx, "samples", ts, "excess shared wavefronts", tc)
print("--- top lines by samples")
for line in sorted(sm, key=lambda l: -sm[l])[:top]:
    print("%5d %5.2f%% samples %5.2f%% instr %5.1f%% conf  %s" % (line, 100 * sm[line] / ts, 100 * ex[line] / tx, 100 * cf[line] / max(tc, 1), src[line - 1].strip()[:100]))
print("--- top lines by excess shared wavefronts")
for line in sorted(cf, key=lambda l: -cf[l])[:8]:
    if cf[line]: print("%5d %5.1f%% conf  %s" % (line, 100 * cf[line] / max(tc, 1), src[line - 1].strip()[:100]))
# phases: functions found by name in the source file
import re
marks = []
for i, l in enumerate(src):
    m = re.match(r"^(?:template <.*>\s*)?(?:__device__ __forceinline__|__global__|__device__|static|inline|__host__ __device__).*?\b(\w+)\s*\(", l)
    if m and not l.startswith(" "): marks.append((i + 1, m.group(1)))
marks.append((len(src) + 1, "end"))
print("--- by function (definition line ranges)")
for (a, name), (b, _) in zip(marks, marks[1:]):
    s_ = sum(v for l, v in sm.items() if a <= l < b); x_ = sum(v for l, v in ex.items() if a <= l < b)
    if s_ > 0.003 * ts: print("%-28s lines %5d-%5d samples %5.1f%% instr %5.1f%%" % (name, a, b - 1, 100 * s_ / ts, 100 * x_ / tx))
