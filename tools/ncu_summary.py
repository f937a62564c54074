#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, on the CPU box) into a small JSON + text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name [frames_per_launch]"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_imma_src_int8_sparsity_off.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__cycles_active.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                try:
                    v = float(v)
                except ValueError:
                    pass
                d[k] = {"value": v, "unit": units[hdr.index(k)]}
        if frames:
            t_ms = d["gpu__time_duration.sum"]["value"]
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["gpu__time_duration.sum"]["unit"], 1.0)
            t_ms *= scale
            d["derived"] = {"frames_per_launch": frames, "frames_per_s": frames / (t_ms * 1e-3),
                            "warp_instr_per_frame": d["smsp__inst_executed.sum"]["value"] / frames,
                            "smem_wavefronts_per_frame": d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]["value"] / frames,
                            "tensor_operand_wavefronts_per_frame": d.get("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", {"value": 0})["value"] / frames,
                            "dram_bytes_per_frame": (_b(d["dram__bytes_read.sum"]) + _b(d["dram__bytes_write.sum"])) / frames,
                            "dram_bytes_per_launch": _b(d["dram__bytes_read.sum"]) + _b(d["dram__bytes_write.sum"])}
        res.append(d)
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        for d in res:
            f.write(d["kernel"] + "\n")
            for k, v in d.items():
                if k == "kernel":
                    continue
                if k == "derived":
                    for kk, vv in v.items():
                        f.write(f"  derived.{kk:60s} {vv:,.3f}\n")
                else:
                    f.write(f"  {k:68s} {v['value']} {v['unit']}\n")
            f.write("\n")
    print(open(out + ".txt").read())


def _b(m):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(m["unit"], 1)
    return m["value"] * mult


if __name__ == "__main__":
    main()
