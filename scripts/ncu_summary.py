"""Summarise `ncu --set full` captures (.ncu-rep) as markdown tables for profiles/ (development aid).
usage: python scripts/ncu_summary.py <file.ncu-rep> <out.md> "<title>" """
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'smsp__sass_inst_executed_op_shared_ld.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def main():
    rep, out, title = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{rep.split('/')[-1]}` (ncu --set full --clock-control none; values are per launch)\n")
        for vals in rows[2:]:
            d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
            if "nan" in d.get("smsp__inst_executed.sum", ("", ""))[1]:
                continue                                    # a capture that was cut off mid-replay
            f.write(f"\n## {d.get('Kernel Name', ('', '?'))[1]}\n\n| metric | unit | value |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            f.write("\nwarp issue-stall reasons (warps per issue-active cycle): ")
            f.write(", ".join(f"{h.split('stalled_')[1].split('_per')[0]} {float(d[h][1]):.3f}" for h in hdr
                              if 'issue_stalled' in h and 'per_issue_active' in h and float(d[h][1]) >= 0.02) + "\n")


if __name__ == "__main__":
    main()
