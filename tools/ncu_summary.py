#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-kernel CSV kept under profiles/.

  python tools/ncu_summary.py gpurun_out/r1p_prof.ncu-rep profiles/r1p_ncu_full_summary.csv

Writes rows `kernel,metric,value,unit` for the metrics the roofline discussion in DESIGN.md uses and refreshes
profiles/roofline_traffic.json (DRAM bytes per launch of the naming kernel = bench.py's `roofline.traffic`)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
           'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'launch__cluster_size']
TO_BYTES = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    text = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    seen = {}
    lines = [('kernel', 'metric', 'value', 'unit')]
    traffic = None
    for r in body:
        name = re.sub(r'^void\s+', '', r[col['Kernel Name']]).split('(')[0]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:
            continue                                # first captured launch of each kernel
        for m in METRICS:
            if m in col:
                lines.append((name, m, r[col[m]], units[col[m]]))
        if name.startswith('name_topk_kernel'):
            rd = float(r[col['dram__bytes_read.sum']]) * TO_BYTES[units[col['dram__bytes_read.sum']]]
            wr = float(r[col['dram__bytes_write.sum']]) * TO_BYTES[units[col['dram__bytes_write.sum']]]
            traffic = dict(name_topk_kernel_dram_bytes_per_launch=int(rd + wr), source=f'{out} (ncu --set full, C2, 1 launch)',
                           read_bytes=int(rd), write_bytes=int(wr))
    with open(out, 'w', newline='') as f:
        csv.writer(f).writerows(lines)
    if traffic is not None:
        with open(os.path.join(os.path.dirname(out), 'roofline_traffic.json'), 'w') as f:
            json.dump(traffic, f, indent=1)
    print(f'{len(lines) - 1} rows, kernels: {sorted(seen)}; traffic: {traffic}')


if __name__ == '__main__':
    main()
