#!/usr/bin/env python
"""Attribute an ncu report's per-SASS samples / executed instructions to CUDA source lines.
ncu's CSV source page is SASS-only; nvdisasm -g gives the line of every SASS instruction of the
same object file, in the same order.
    python scripts/ncu_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [top_n]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def sass_lines(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines, cur, active = [], None, False
    for ln in out.split('\n'):
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
        if m:
            active = kernel in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File ".*?", line (\d+)', ln)
        if m:
            cur = int(m.group(1))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            lines.append(cur)
    return lines


def main():
    rep, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
    h = rows[hi]
    si, ii, bi = h.index('# Samples'), h.index('Instructions Executed'), h.index('stall_barrier')
    sass = [r for r in rows[hi + 1:] if len(r) > ii]
    lines = sass_lines(obj, kernel)
    if len(lines) != len(sass):
        print('warning: %d SASS rows in the report vs %d in the object' % (len(sass), len(lines)))
    cu = os.path.splitext(obj)[0] + '.cu'
    src = open(cu).read().split('\n') if os.path.exists(cu) else []
    agg = {}
    for r, l in zip(sass, lines):
        a = agg.setdefault(l, [0, 0, 0])
        a[0] += int(r[si]); a[1] += int(r[ii]); a[2] += int(r[bi] or 0)
    ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print('total samples %d, warp instructions %d' % (ts, ti))
    for l, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src[l - 1].strip()[:100] if l and l <= len(src) else ''
        print('%5.1f%% samples (%4.1f%% barrier) %5.1f%% instr  L%-5s %s' % (100 * a[0] / ts, 100 * a[2] / ts, 100 * a[1] / ti, l, text))


if __name__ == '__main__':
    main()
