"""Text summary of an `ncu --set full --import-source on` report for profiles/ (run here, no GPU needed):
  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_<kernel>_ncu.txt
Prints the launch's key raw metrics, the stall reasons above 0.2 per issue, and the SASS regions (runs of instructions
with the same execution count) that hold more than 1 % of the executed instructions or of the stall samples."""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active']


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("--- launch")
        for w in WANT:
            if w in d:
                print(f"{w:88s} {d[w]} {units[hdr.index(w)]}")
        for k in hdr:
            if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
                try:
                    if float(d[k]) > 0.2:
                        print(f"{k:88s} {d[k]}")
                except ValueError:
                    pass
    rows = page(rep, "source", ("--print-source", "sass"))
    hdr = rows[1]
    isrc, iex, ithr, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
    data = [(r[isrc], int(r[iex] or 0), float(r[ithr] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > iex]
    tot, tots = sum(d[1] for d in data) or 1, sum(d[3] for d in data) or 1
    print(f"--- SASS regions (of {tot} executed warp-instructions, {tots} stall samples)")
    segs, start = [], 0
    for i in range(1, len(data) + 1):
        if i == len(data) or abs(data[i][1] - data[start][1]) > 0.15 * max(data[start][1], 1):
            segs.append((start, i))
            start = i
    for a, b in segs:
        n, sm = sum(d[1] for d in data[a:b]), sum(d[3] for d in data[a:b])
        if n / tot > 0.01 or sm / tots > 0.01:
            print(f"[{a:4d},{b:4d}) {b - a:3d} instr x {data[a][1] / 1e6:8.2f} M  = {n * 100 / tot:5.2f} % of instructions, {sm * 100 / tots:5.2f} % of samples, "
                  f"{data[a][2]:4.1f} threads/instr   first: {data[a][0].strip()[:60]}")


if __name__ == "__main__":
    main()
