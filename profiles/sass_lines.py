#!/usr/bin/env python
"""Static SASS histogram of one kernel by source line (no GPU needed).

    python profiles/sass_lines.py nasbowl_b200/libnbw.so score_pool.sm_100a.cubin 'score_pool_kernelILi8ELi6ELi2E'

Extracts the cubin with cuobjdump, disassembles it with `nvdisasm -g -c` (the library is built with
-lineinfo) and counts instructions per (file, line) and per opcode class.  Static counts are not
executed counts (loops, predication), but they show where the instruction budget of a straight-line
kernel body sits before GPU time is spent on it.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    so, cubin, pat = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", cubin, os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    in_fn = False
    cur = ("?", 0)
    by_line = collections.Counter()
    by_op = collections.Counter()
    total = 0
    for ln in sass.splitlines():
        if ln.startswith("\t.section\t.text."):
            in_fn = re.search(pat, ln) is not None
            continue
        if not in_fn:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(2)
            if op in ("NOP",):
                continue
            total += 1
            by_line[cur] += 1
            by_op[op.split(".")[0]] += 1
    print("kernel /%s/: %d static instructions" % (pat, total))
    print("-- by opcode")
    for op, c in by_op.most_common(25):
        print("  %-10s %5d" % (op, c))
    print("-- by source line")
    for (f, l), c in by_line.most_common(top):
        print("  %-22s:%-4d %5d" % (f, l, c))


if __name__ == "__main__":
    main()
