#!/usr/bin/env python
"""Executed warp instructions of one kernel by source line: joins the SASS page of an ncu report
(`ncu -i rep --page source --csv`) with the line table of the same cubin (`nvdisasm -g -c`).

    python profiles/ncu_lines.py <sass_page.csv> <nvdisasm.sass> <mangled kernel substring> <cells per launch> [top]
"""
import collections
import csv
import os
import re
import sys


def main():
    page, sass, pat, cells = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 45
    rows = list(csv.reader(open(page)))
    hdr = rows[1]
    ia, isrc, ie = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed')
    data = [(int(r[ia], 16), r[isrc].strip(), int(r[ie])) for r in rows[2:] if len(r) > ie and r[ia].startswith('0x')]
    # the page lists every profiled launch one after another: keep the first
    base = data[0][0]
    first = [data[0]]
    for d in data[1:]:
        if d[0] == base:
            break
        first.append(d)
    cur, in_fn, line_of = [('?', 0)], False, {}
    for ln in open(sass):
        if ln.startswith('\t.section\t.text.'):
            in_fn = pat in ln
            continue
        if not in_fn:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', ln)
        if m:
            chain = [(os.path.basename(m.group(1)), int(m.group(2)))]
            for mm in re.finditer(r'inlined at "(.*?)", line (\d+)', m.group(3)):
                chain.append((os.path.basename(mm.group(1)), int(mm.group(2))))
            cur = chain
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*);', ln)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2))
    by, byop, tot = collections.Counter(), collections.Counter(), 0
    for a, s, e in first:
        ch, ins = line_of.get(a - base, ([('?', 0)], s))
        # attribute to the outermost frame that lies in the kernel's own file, else the innermost
        own = [c for c in ch if c[0].startswith('score_')]
        key = own[-1] if own else ch[0]
        by[key] += e
        tot += e
        byop[re.sub(r'^@!?U?P\d+\s+', '', ins).split()[0].split('.')[0]] += e
    print("total warp instructions %d = %.1f per cell" % (tot, tot / cells))
    for k, v in byop.most_common(16):
        print("  %-12s %7.2f per cell %5.1f%%" % (k, v / cells, 100 * v / tot))
    print("-- by source line (outermost frame in the kernel file)")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:top]:
        print("  %-20s:%-4d %7.2f per cell %5.1f%%" % (k[0], k[1], v / cells, 100 * v / tot))


if __name__ == "__main__":
    main()
