#!/usr/bin/env python
"""Attribute executed instructions / stall samples of one kernel in an .ncu-rep to CUDA source lines.
usage: ncu_lines.py report.ncu-rep lib.so mangled_kernel_substring [top_n]"""
import collections, csv, re, subprocess, sys, tempfile, os, glob

rep, lib, sub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = glob.glob(os.path.join(tmp, '*.cubin'))[0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel with their source line
lines = []
infn = False
cur = None
for l in dis:
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        infn = sub in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ei = hdr.index('Instructions Executed'); ss = hdr.index('Warp Stall Sampling (All Samples)')
body = [r for r in rows[2:] if len(r) > ei and r[ei].isdigit()]
if len(body) != len(lines):
    print('warning: %d profiled instructions vs %d disassembled' % (len(body), len(lines)))
agg = collections.defaultdict(lambda: [0, 0])
for r, ln in zip(body, lines):
    a = agg[ln]; a[0] += int(r[ei]); a[1] += int(r[ss] or 0)
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
src = {}
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        p = [q for q in glob.glob('/root/repo/raynet_b200/csrc/*') + glob.glob('/root/repo/include/*') if os.path.basename(q) == f]
        src[f] = open(p[0]).read().splitlines() if p else []
    text = src[f][n - 1].strip()[:95] if src[f] and n <= len(src[f]) else ''
    print('%5.1f%% inst %5.1f%% stall  %s:%d  %s' % (100.0 * a[0] / tot, 100.0 * a[1] / max(tots, 1), f, n, text))
