#!/usr/bin/env python
"""Attribute executed warp instructions / stall samples of one kernel to CUDA source lines.

usage: ncu_by_line.py <report.ncu-rep> <kernel-substring> <lib.so> [top]

Joins `ncu --page source --csv` (per-SASS-instruction counters) with `nvdisasm -g` line info of
the same cubin by instruction offset inside the kernel's .text section.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(report, kernel):
    # `kernel`: a regex on the kernel name, or `id:N` = the N-th profiled launch of the report
    sel = ['--kernel-id', ':::' + kernel[3:]] if kernel.startswith('id:') else ['-k', 'regex:' + kernel]
    out = subprocess.run(['ncu', '-i', report, '--page', 'source', '--csv'] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks = []
    cur = None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] is not None and r:
            cur['rows'].append(r)
    return blocks[0]


def line_info(lib, mangled_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    info = {}
    active = False
    cur = ('?', 0)
    for line in dis.split('\n'):
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
        if m:
            active = mangled_substr in m.group(1)
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', line)
        if m:
            info[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return info


def main():
    report, kernel, lib = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    blk = sass_rows(report, kernel)
    ix = {h: i for i, h in enumerate(blk['hdr'])}
    # section name of exactly this instantiation: beam_kernel<(int)1, (int)32, (bool)0> -> beam_kernelILi1ELi32ELb0EE
    full = blk['name']
    base_name = re.search(r'(\w+)\s*(<|\()', re.sub(r'^void\s+', '', full).replace('lt::', '')).group(1)
    mangled = base_name
    m = re.search(r'<([^>]*)>', full)
    if m:
        parts = []
        for arg in m.group(1).split(','):
            t, v = re.match(r'\s*\((\w+)\)(-?\d+)', arg).groups()
            parts.append({'int': 'Li', 'bool': 'Lb', 'unsigned int': 'Lj'}.get(t, 'Li') + v + 'E')
        mangled = base_name + 'I' + ''.join(parts) + 'E'
    info = line_info(lib, mangled)
    base = int(blk['rows'][0][ix['Address']], 16) if blk['rows'][0][ix['Address']].startswith('0x') else int(blk['rows'][0][ix['Address']])
    by_line = collections.defaultdict(lambda: [0.0, 0.0])
    tot_i = tot_s = 0.0
    for r in blk['rows']:
        a = r[ix['Address']]
        off = (int(a, 16) if a.startswith('0x') else int(a)) - base
        inst = float(r[ix['Instructions Executed']] or 0)
        samp = float(r[ix['# Samples']] or 0)
        key = info.get(off, (('?', 0), ''))[0]
        by_line[key][0] += inst
        by_line[key][1] += samp
        tot_i += inst
        tot_s += samp
    print('kernel %s: %.0f warp instructions, %.0f samples' % (blk['name'], tot_i, tot_s))
    srcs = {}
    for (fn, ln), (inst, samp) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        if fn not in srcs:
            path = os.path.join(os.path.dirname(os.path.abspath(lib)), 'csrc', fn)
            if not os.path.exists(path):     # a copy of the library kept elsewhere: use this repo's sources
                path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'lattice_based_tagger_b200', 'csrc', fn)
            srcs[fn] = open(path).read().split('\n') if os.path.exists(path) else []
        text = srcs[fn][ln - 1].strip()[:90] if 0 < ln <= len(srcs[fn]) else ''
        print('%5.2f%% inst %5.2f%% stall  %s:%d  %s' % (100 * inst / tot_i, 100 * samp / max(1, tot_s), fn, ln, text))


if __name__ == '__main__':
    main()
