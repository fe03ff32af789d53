#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into markdown for profiles/: per captured launch the
duration, DRAM traffic, pipe utilisation, occupancy and the warp-stall breakdown.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/rNN_xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s\n" % title)
    print("Source: `%s` (ncu --set full --clock-control none), read with `ncu -i ... --page raw --csv`.\n" % rep)
    for r in data:
        print("## launch %s: `%s`\n" % (r[col["ID"]], r[col["Kernel Name"]]))
        print("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in col:
                print("| %s (`%s`) | %s | %s |" % (label, k, r[col[k]], units[col[k]]))
        stalls = []
        for h, i in col.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls)
        if tot > 0:
            print("\nWarp-state samples (pc sampling, % of all samples, top 8): " +
                  ", ".join("%s %.1f%%" % (n, 100 * v / tot) for v, n in sorted(stalls, reverse=True)[:8]))
        print()


if __name__ == "__main__":
    main()
