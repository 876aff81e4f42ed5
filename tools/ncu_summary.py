"""Summarise `ncu -i <rep> --page raw --csv` (run here, no GPU needed): one block per captured launch with the counters
DESIGN.md quotes, and — with --traffic-json — the per-kernel-family DRAM bytes per launch that bench.py's
roofline.traffic reads (profiles/r02_roofline_traffic.json).

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv
    python tools/ncu_summary.py /tmp/x.csv [--traffic-json profiles/r02_roofline_traffic.json --source profiles/<file>]
"""
import argparse
import csv
import json
import os
import re

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX (shared-memory pipe) %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "  tensor-core operand wavefronts % of the smem pipe"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "  LSU (ld/st.shared, mbarrier) wavefronts %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory path active %"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (realtime) %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
]
FAMILIES = ["c12_kernel", "ws2conv_kernel", "ws2x_conv_kernel", "ptcw_conv_kernel", "ptc2_conv_kernel", "xf_kernel", "tc_gemm_kernel",
            "rvk_conv2_kernel"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--traffic-json")
    ap.add_argument("--source", default="")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    name_i = ix["Kernel Name"]
    fam_bytes = {}
    for r in data:
        name = r[name_i]
        short = re.sub(r"\(.*", "", name)
        print(f"== {short}")
        for key, label in KEYS:
            if key in ix:
                print(f"   {label:<62s} {r[ix[key]]} {units[ix[key]]}")
        for fam in FAMILIES:
            if fam in name:
                def mb(k):
                    v = float(r[ix[k]])
                    u = units[ix[k]].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                fam_bytes.setdefault(fam, []).append(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
    if args.traffic_json:
        cur = json.load(open(args.traffic_json)) if os.path.exists(args.traffic_json) else {}
        for fam, v in fam_bytes.items():
            cur[fam] = {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v),
                        "source": args.source or os.path.basename(args.csv),
                        "what": "ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum, average over the captured launches"}
        json.dump(cur, open(args.traffic_json, "w"), indent=1, sort_keys=True)
        print("wrote", args.traffic_json)


if __name__ == "__main__":
    main()
