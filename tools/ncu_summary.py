"""Markdown summary of one `ncu --set full` report (first profiled launch) -- the files under profiles/.
usage: python tools/ncu_summary.py report.ncu-rep "title" <algorithmic bytes per launch> [out.md]"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, title, alg = sys.argv[1], sys.argv[2], float(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    name = vals[col["Kernel Name"]] if "Kernel Name" in col else "?"
    lines = [f"# ncu --set full summary: {title}", "",
             f"Kernel: `{name}`.  Report: `gpurun_out/{rep.split('/')[-1]}` (scratch; `ncu --set full --clock-control none --import-source on`).", "",
             "| metric | unit | value |", "|---|---|---|"]
    for m in METRICS:
        if m in col:
            lines.append(f"| `{m}` | {units[col[m]]} | {vals[col[m]]} |")
    stalls = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(vals[col[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    lines += ["", "Top warp stall reasons (cycles per issued instruction): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8])]
    try:
        rd = float(vals[col["dram__bytes_read.sum"]]) * TO_BYTES[units[col["dram__bytes_read.sum"]]]
        wr = float(vals[col["dram__bytes_write.sum"]]) * TO_BYTES[units[col["dram__bytes_write.sum"]]]
        lines += ["", f"DRAM traffic per launch: read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = **{(rd + wr) / 1e6:.1f} MB**; "
                      f"algorithmic {alg / 1e6:.1f} MB (ratio {(rd + wr) / alg:.3f})."]
        print(f"TRAFFIC {int(rd + wr)}")
    except Exception:
        pass
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
