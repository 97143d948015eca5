"""Writes the markdown summary + traffic json of an `ncu --set full` report (raw page) for profiles/.
Usage: python tools/ncu_summary.py report.ncu-rep out.md [traffic.json n_env] ["header line"]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
traffic_json = sys.argv[3] if len(sys.argv) > 3 else None
n_env = int(sys.argv[4]) if len(sys.argv) > 4 else 0
header = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__average_warps_issue_stalled_"
lines = [header, ""] if header else []
traffic = None
to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in data:
    name = r[hdr.index("Kernel Name")]
    lines.append(f"## {name}")
    for k in KEEP:
        if k in hdr:
            lines.append(f"- {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(h[len(STALL):].replace("_per_issue_active.ratio", ""), float(r[i])) for i, h in enumerate(hdr)
          if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
    tot = sum(v for _, v in st) or 1.0
    lines.append("- warp stall mix (share of stalled+selected warp-cycles): " +
                 ", ".join(f"{n} {100 * v / tot:.1f}%" for n, v in sorted(st, key=lambda t: -t[1])[:8]))
    lines.append("")
    if traffic is None:
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic = float(r[i_r]) * to_bytes[units[i_r]] + float(r[i_w]) * to_bytes[units[i_w]]
        tname = name
open(out, "w").write("\n".join(lines))
if traffic_json:
    json.dump({"n_env": n_env, "kernel": tname, "dram_bytes_per_launch": traffic,
               "source": f"ncu --set full --clock-control none ({rep.split('/')[-1]})"}, open(traffic_json, "w"))
print("wrote", out, traffic)
