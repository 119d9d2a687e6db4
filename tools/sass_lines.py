"""Joins the per-instruction dynamic counts of an ncu source-page CSV (`ncu -i X.ncu-rep --page source --csv
--print-source sass --kernel-name regex:K`) with nvdisasm line info of the cubin, and prints dynamic instruction
counts per CUDA source line and per opcode.  Development tool.

    python tools/sass_lines.py <sass.csv> <cubin> <mangled-substring> <units (e.g. voxel-warps)>
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    sass_csv, cubin, func, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    iS, iI = hdr.index('Source'), hdr.index('Instructions Executed')
    data, seen = [], set()
    for r in rows[2:]:
        if len(r) > 10 and r[0].startswith('0x'):
            if r[0] in seen:
                break
            seen.add(r[0])
            data.append((r[iS].strip(), int(r[iI])))
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    lines, cur, infunc, insts = [], None, False, []
    for ln in txt.splitlines():
        if ln.startswith('.text.') or ln.lstrip().startswith('.section'):
            infunc = func in ln
            continue
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            insts.append(cur)
    print("ncu instructions: %d, nvdisasm instructions: %d" % (len(data), len(insts)))
    n = min(len(data), len(insts))
    per_line = collections.Counter()
    tot = sum(c for _, c in data)
    for k in range(n):
        per_line[insts[k]] += data[k][1]
    src_cache = {}
    print("total dynamic warp-instructions: %d = %.1f per unit" % (tot, tot / units))
    for (key, c) in per_line.most_common(60):
        text = ''
        if key:
            f = key[0]
            if f not in src_cache:
                try:
                    p = subprocess.run(['find', '/root/repo/brainfm_b200', '/root/repo/include', '-name', f],
                                       capture_output=True, text=True).stdout.split()[0]
                    src_cache[f] = open(p).read().splitlines()
                except Exception:
                    src_cache[f] = []
            if key[1] - 1 < len(src_cache[f]):
                text = src_cache[f][key[1] - 1].strip()
        print("%6.2f%% %6.1f  %s:%s  %s" % (100 * c / tot, c / units, key[0] if key else '?', key[1] if key else '?', text[:100]))


if __name__ == '__main__':
    main()
