#!/usr/bin/env python
"""Turn an `ncu --set full` capture of one resolve launch into the JSON record bench.py reads for `roofline.traffic`
and the `binding` block (profiles/resolve_ncu_<config>[_full].json), and print a short text summary.

    python profiles/tools/ncu_to_json.py gpurun_out/x.ncu-rep C4 [--full] [--note "..."]
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
L2_STREAM_PEAK = 13.9e12     # bytes/s L2 -> SM, measured in round 1 (loads-only build of the resolve kernel, r01_v6_resolve_ncu.md)


def metrics(rep: str, index: int = -1) -> dict:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:][index]
    m = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            m[h] = v
            continue
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u)
        m[h] = x * scale if scale else x
    return m


def main():
    rep, config = sys.argv[1], sys.argv[2]
    full = "--full" in sys.argv
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    m = metrics(rep)
    rec = {
        "kernel": m.get("Kernel Name"),
        "duration_s_under_ncu": m.get("gpu__time_duration.sum"),
        "dram_bytes_per_launch": m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0),
        "l2_to_sm_bytes_per_launch": 32.0 * m.get("lts__t_sectors_srcunit_tex_op_read.sum", 0),
        "l2_hit_rate_pct": m.get("lts__t_sector_hit_rate.pct"),
        "l2_stream_peak_bytes_per_s": L2_STREAM_PEAK,
        "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": m.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": m.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": m.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": m.get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_pct": m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "smem_wavefronts_pct": m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "warp_instructions": m.get("smsp__inst_executed.sum"),
        "registers_per_thread": m.get("launch__registers_per_thread"),
        "source": f"ncu --set full --clock-control none, {Path(rep).name}" + (f"; {note}" if note else ""),
    }
    path = ROOT / "profiles" / f"resolve_ncu_{config}{'_full' if full else ''}.json"
    path.write_text(json.dumps(rec, indent=1) + "\n")
    print(json.dumps(rec, indent=1))
    stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: v for k, v in m.items()
              if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and isinstance(v, float)}
    print("stalls per issue:", {k: round(v, 2) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]})


if __name__ == "__main__":
    main()
