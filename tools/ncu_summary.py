"""Turns an .ncu-rep into a small markdown summary for profiles/ (run in the build container, no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_forward_kernel.md "title"
"""
import collections, csv, io, subprocess, sys

rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on)", ""]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    lines.append(f"## {d.get('Kernel Name', '?')[:110]}  (launch id {d.get('ID', '?')})")
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    for k in want:
        if k in d:
            lines.append(f"| {k} | {d[k]} | {u[k]} |")
    st = []
    for k, v in d.items():
        if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio"):
            try: st.append((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    lines.append("")
    lines.append("warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    lines.append("")
# SASS opcode mix, one table per captured launch: the source page lists the launches one after the other, each introduced
# by a "Kernel Name" row and an "Address" header row (summing over all of them, as the first version of this tool did, gives
# a total that belongs to no launch)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sections, cur = [], None
for r in srows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "?", "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
seen = set()
for sec in sections:
    h = sec["hdr"]
    if h is None:
        continue
    iS, iX, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    ops, smp = collections.Counter(), collections.Counter()
    for r in sec["rows"]:
        try: x, s = int(r[iX]), int(r[iS])
        except (ValueError, IndexError): continue
        toks = r[iSrc].split()
        if not toks: continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        ops[op] += x; smp[op] += s
    tx, ts = sum(ops.values()), max(sum(smp.values()), 1)
    if (sec["name"], tx) in seen:      # the page repeats a launch (views); one table each
        continue
    seen.add((sec["name"], tx))
    lines += [f"SASS opcode mix of `{sec['name'][:80]}` (warp-level instructions executed: {tx} = smsp__inst_executed.sum of that launch; stall samples):", "",
              "| opcode | executed | share | samples |", "|---|---|---|---|"]
    for op, x in ops.most_common(12):
        lines.append(f"| {op} | {x} | {100.0 * x / max(tx, 1):.1f}% | {100.0 * smp[op] / ts:.1f}% |")
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
