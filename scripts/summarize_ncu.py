"""Turn ncu outputs (brought back in gpurun_out/) into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches_summary.md
    python scripts/summarize_ncu.py report   gpurun_out/prof.ncu-rep  > profiles/rNN_kernel_summary.md
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    name_i, val_i = rows[hdr].index('Kernel Name'), rows[hdr].index('Metric Value')
    fam = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows[hdr + 1:]:
        if len(r) <= val_i:
            continue
        try:
            ns = float(r[val_i].replace(',', ''))
        except ValueError:
            continue
        name = re.sub(r'^void ', '', r[name_i]).split('(')[0]
        name = re.sub(r'<.*', '', name) if not name.startswith('ngan::') else name
        fam[name][0] += 1
        fam[name][1] += ns
        total += ns
    print(f'ncu launch list `{path}`: {sum(v[0] for v in fam.values())} launches, {total / 1e6:.3f} ms of kernel time '
          '(cold-cache, serialised: compare SHARES, not absolutes)\n')
    print('| kernel | launches | total us | share |')
    print('|---|---:|---:|---:|')
    for k, (n, ns) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {n} | {ns / 1e3:.1f} | {ns / total:.1%} |')


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tc.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']


def report(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f'ncu --set full capture `{path}` (per launch)\n')
    for r in rows[2:]:
        print(f'### `{r[hdr.index("Kernel Name")][:110]}`  grid {r[hdr.index("Grid Size")]} block {r[hdr.index("Block Size")]}\n')
        print('| metric | value | unit |')
        print('|---|---:|---|')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'| {w} | {r[i]} | {units[i]} |')
        print()


if __name__ == '__main__':
    {'launches': launches, 'report': report}[sys.argv[1]](sys.argv[2])
