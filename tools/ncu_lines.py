#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep.

Joins ncu's SASS-level source page (--page source --csv) with the line table of the shipped cubin
(nvdisasm -g on the cubin extracted from libfreud_b200.so; needs -lineinfo at compile time).
usage: ncu_lines.py report.ncu-rep kernel_regex mangled_substring [top_n] [launch_index]"""
import collections, csv, glob, os, re, subprocess, sys, tempfile

rep, kern, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
skip = sys.argv[5] if len(sys.argv) > 5 else '0'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(root, 'freud_b200', 'libfreud_b200.so')], cwd=tmp,
               capture_output=True)
lines_of = {}
for cubin in glob.glob(os.path.join(tmp, '*.cubin')):
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    cur_fn, cur_line, offsets = None, None, None
    for ln in txt.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
        if m:
            cur_fn = m.group(1)
            offsets = lines_of.setdefault(cur_fn, {})
            cur_line = None
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
        if m and cur_fn is not None:
            offsets[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
cands = [f for f in lines_of if mangled in f]
assert len(cands) == 1, f'{len(cands)} functions match {mangled}: {cands[:5]}'
table = lines_of[cands[0]]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}', '--launch-skip',
                      skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ii, si, ti = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for r in rows[2:]:
    try:
        addr, inst, samp, tinst = int(r[ia], 16), int(r[ii]), int(r[si]), int(r[ti])
    except (ValueError, IndexError):
        continue
    base = addr if base is None else base
    line = table.get(addr - base, (None, ''))[0] or ('?', 0)
    a = agg[line]
    a[0] += inst
    a[1] += samp
    a[2] += tinst
    tot_i += inst
    tot_s += samp
src_cache = {}
def src(f, n):
    if f not in src_cache:
        p = os.path.join(root, 'freud_b200', 'csrc', f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    l = src_cache[f]
    return l[n - 1].strip()[:100] if 0 < n <= len(l) else ''
print(f'{cands[0][:80]}: {tot_i} warp instructions, {tot_s} stall samples')
for line, (inst, samp, tinst) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100*inst/max(tot_i,1):5.1f}% inst {100*samp/max(tot_s,1):5.1f}% samp  thr/inst {tinst/max(inst,1):4.1f}  {line[0]}:{line[1]:<4d} {src(*line)}')
