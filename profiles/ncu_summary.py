"""Summarise an .ncu-rep (raw page + source page) into the few numbers DESIGN.md cites.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [nquartets]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
nq = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or "issue_stalled" in h and "per_issue_active" in h:
        print(f"{h:75s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, samp = collections.Counter(), collections.Counter()
tot = tots = 0
for r in rows[2:]:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    op = m.group(2).split(".")[0] if m else "?"
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    ops[op] += n; samp[op] += s; tot += n; tots += s
print(f"warp-instructions {tot}" + (f" = {tot / nq:.0f} per quartet" if nq else ""))
for op, n in ops.most_common(16):
    print(f"  {op:8s} {100 * n / tot:5.1f}% inst  {100 * samp[op] / max(1, tots):5.1f}% samples" + (f"  {n / nq:8.0f}/quartet" if nq else ""))
