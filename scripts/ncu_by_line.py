"""Per-source-line view of an ncu capture without the GUI: joins the SASS rows of `ncu --page source --csv` (instructions
executed, stall samples per instruction) with the file:line of every instruction from `nvdisasm -g` on the library's cubin
(built with -lineinfo), and prints the hottest lines / files.

usage: python scripts/ncu_by_line.py <report.ncu-rep> <kernel-name-substring> [--top 40] [--lib babyai_b200/libbabyai_b200.so]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
    out = subprocess.run(['nvdisasm', '-g', os.path.join(d, cub)], capture_output=True, text=True).stdout
    res, cur, on = [], ('?', 0), False
    for ln in out.splitlines():
        if ln.startswith('//---') and '.text.' in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', ln)
        if m:
            res.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('kernel')
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--lib', default='babyai_b200/libbabyai_b200.so')
    ap.add_argument('--nth', type=int, default=0, help='which matching launch of the report')
    ap.add_argument('--mangled', default=None, help='substring of the mangled name in the cubin (e.g. k_rollout_ctaILb0), default: the kernel name')
    a = ap.parse_args()
    txt = subprocess.run(['ncu', '-i', a.report, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    # the csv holds one block per profiled launch: "Kernel Name",<name> / header / rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None and row:
            cur['rows'].append(row)
    blocks = [b for b in blocks if a.kernel in b['name']]
    if not blocks:
        sys.exit('no launch of %r in the report' % a.kernel)
    b = blocks[min(a.nth, len(blocks) - 1)]
    hdr, rows = b['rows'][0], b['rows'][1:]
    ci = {h: i for i, h in enumerate(hdr)}
    sass = sass_lines(a.lib, a.mangled or a.kernel.split('<')[0])
    if len(sass) != len(rows):
        print('warning: %d SASS instructions in the library vs %d in the report (different build?)' % (len(sass), len(rows)))
    n = min(len(sass), len(rows))
    by_line = collections.defaultdict(lambda: [0, 0, 0])
    line_stalls = collections.defaultdict(collections.Counter)
    by_file = collections.defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    stalls = collections.Counter()
    scols = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
    for k in range(n):
        r = rows[k]
        ins = int(float(r[ci['Instructions Executed']] or 0))
        smp = int(float(r[ci['# Samples']] or 0))
        thr = int(float(r[ci['Thread Instructions Executed']] or 0))
        key = sass[k][1]
        by_line[key][0] += ins; by_line[key][1] += smp; by_line[key][2] += thr
        by_file[key[0]][0] += ins; by_file[key[0]][1] += smp
        tot_i += ins; tot_s += smp
        for h in scols:
            v = int(float(r[ci[h]] or 0))
            stalls[h] += v
            line_stalls[key][h[6:]] += v
    print('%s\n  %d SASS instructions, %.2f M warp instructions executed, %d stall samples' % (b['name'][:100], n, tot_i / 1e6, tot_s))
    print('  stall samples: ' + ', '.join('%s %.1f%%' % (h[6:], 100.0 * v / max(1, sum(stalls.values()))) for h, v in stalls.most_common(8)))
    print('by file:')
    for f, (i, s) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        print('  %-22s %6.2f%% of instructions  %6.2f%% of samples' % (f, 100.0 * i / tot_i, 100.0 * s / max(1, tot_s)))
    print('hottest lines (by warp instructions executed):')
    for (f, l), (i, s, t) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print('  %-20s:%-5d %6.2f%% instr  %6.2f%% samples  %5.1f thr/instr' % (f, l, 100.0 * i / tot_i, 100.0 * s / max(1, tot_s), t / max(1, i)))
    print('hottest lines (by stall samples) and their main stall reasons:')
    for (f, l), (i, s, t) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:a.top // 2]:
        top = ', '.join('%s %d' % kv for kv in line_stalls[(f, l)].most_common(3))
        print('  %-20s:%-5d %6.2f%% samples  %6.2f%% instr   %s' % (f, l, 100.0 * s / max(1, tot_s), 100.0 * i / tot_i, top))


if __name__ == '__main__':
    main()


def _unused():
    pass
