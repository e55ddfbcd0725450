#!/usr/bin/env python
"""Selected counters of the lattice / beam kernels from an `ncu --set full` report, as JSON.

usage: ncu_summary.py <report.ncu-rep> <out.json> [traffic.json <config>]

Takes the LONGEST captured launch of each kernel (a capture window also holds the retry pass of the
lattice kernel and launches that exit early because a buffer overflowed during warm-up).  With the last two arguments it also records, under the key <config> of traffic.json, the
DRAM bytes per launch that bench.py reports as `roofline.traffic` for that configuration
(configurations without a capture report null).
"""
import csv
import io
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']


def to_bytes(value, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    return float(value) * scale


def main():
    report, out = sys.argv[1:3]
    raw = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    unit_of_name = dict(zip(hdr, units))
    last = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d['Kernel Name']
        key = 'beam_kernel' if 'beam_kernel' in name else ('lattice_kernel' if 'lattice_kernel' in name else None)
        if key:
            def duration(row):
                try:
                    return float(row['gpu__time_duration.sum']) * {'us': 1e-3, 'ms': 1.0, 'ns': 1e-6, 's': 1e3}.get(unit_of_name.get('gpu__time_duration.sum', 'ms'), 1.0)
                except (KeyError, ValueError):
                    return 0.0
            if key not in last or duration(d) > duration(last[key]):
                last[key] = d
    unit_of = dict(zip(hdr, units))
    summary, traffic = [], {}
    for key, d in last.items():
        item = {'Kernel Name': d['Kernel Name']}
        for w in WANT:
            if w in d:
                item[w] = '%s %s' % (d[w], unit_of.get(w, ''))
        for h in hdr:
            if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h and float(d[h] or 0) >= 0.05:
                item[h] = d[h]
        summary.append(item)
        traffic[key] = int(to_bytes(d['dram__bytes_read.sum'], unit_of['dram__bytes_read.sum']) +
                           to_bytes(d['dram__bytes_write.sum'], unit_of['dram__bytes_write.sum']))
    with open(out, 'w') as f:
        json.dump(summary, f, indent=1)
    if len(sys.argv) > 4:
        path, config = sys.argv[3], sys.argv[4]
        try:
            with open(path) as f:
                table = json.load(f)
        except Exception:
            table = {}
        traffic['_source'] = '%s: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full' % out
        table[config] = traffic
        with open(path, 'w') as f:
            json.dump(table, f, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == '__main__':
    main()
