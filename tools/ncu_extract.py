"""Turn `ncu -i X.ncu-rep --page raw --csv` into the short per-kernel summaries kept under profiles/:
    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/x.csv
    python tools/ncu_extract.py /tmp/x.csv "title line" > profiles/rNN_x.txt
and, with --lines, aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by source line
(samples, instructions, live lanes) to see where a kernel spends its issue slots:
    python tools/ncu_extract.py --lines /tmp/x_src.csv [top_n]"""
import csv
import sys

KEYS = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']


def summary(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print("# " + title)
    print("# ncu --set full --clock-control none --import-source on")
    ik = hdr.index('Kernel Name')
    for r in rows[2:]:
        print("kernel:", r[ik])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:85s} {units[i]:16s} {r[i]}")
        print("-----")


def lines(path, top):
    hdr, agg = None, {}
    for r in csv.reader(open(path)):
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        try:
            ln, smp = int(r[0]), int(r[hdr.index("# Samples")])
            ex, tex = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("Thread Instructions Executed")])
        except ValueError:
            continue
        a = agg.setdefault(ln, [0, 0, 0, r[1]])
        a[0] += smp; a[1] += ex; a[2] += tex
    tot, totex = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"samples {tot}, warp instructions {totex}")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{ln:5d}  samples {100 * a[0] / tot:5.1f} %  inst {100 * a[1] / totex:5.1f} %  lanes {a[2] / max(1, a[1]):4.1f}  {a[3][:100]}")


if __name__ == "__main__":
    if sys.argv[1] == "--lines":
        lines(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
    else:
        summary(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
