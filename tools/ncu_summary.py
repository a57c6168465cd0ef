"""Summarise ncu outputs kept under gpurun_out/ (launch list CSV, full-capture .ncu-rep) as text tables."""
import collections, csv, io, subprocess, sys

def launches(path, top=30):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        if row.get('Metric Name') != 'gpu__time_duration.sum': continue
        v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
        a = agg.setdefault(row['Kernel Name'], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{len(agg)} kernels, {tot:.0f} us")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{a[1]:10.1f} us {a[0]:5d}x {100 * a[1] / tot:5.1f}%  {k[:100]}")

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']

def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx['Kernel Name']][:28] for r in rows[2:]]
    print(f"{'metric':88s}" + ''.join(f"{n:>30s}" for n in names))
    for m in WANT:
        if m in idx:
            print(f"{m + ' [' + units[idx[m]] + ']':88s}" + ''.join(f"{r[idx[m]]:>30s}" for r in rows[2:]))

if __name__ == '__main__':
    (launches if sys.argv[1] == 'launches' else full)(sys.argv[2])
