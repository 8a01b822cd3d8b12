#!/usr/bin/env python
"""Dynamic warp-instruction counts per SOURCE LINE of one kernel: joins the per-SASS-instruction executed
counts of an ncu report (--page source) with nvdisasm's line table of the same libvsl.so build.

  python profiles/line_profile.py <report.ncu-rep> <kernel-regex> <mangled-function-substring> [libvsl.so]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kre, fsub = sys.argv[1:4]
lib = sys.argv[4] if len(sys.argv) > 4 else 'tf_depth_estimation_b200/libvsl.so'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ie, isrc = hdr.index('Instructions Executed'), hdr.index('Source')
ismp = hdr.index('# Samples')
ilsb = hdr.index('stall_long_sb')
body = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) > 10:
        body.append((r[isrc].strip(), int(r[ie]), int(r[ismp] or 0), int(r[ilsb] or 0)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, on = [], None, False
for l in dis.splitlines():
    if l.startswith('//--------------------- .text.'):
        on = fsub in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
        lines.append(cur)
assert len(lines) == len(body), (len(lines), len(body))
tot = sum(b[1] for b in body)
stot = sum(b[2] for b in body)
per = collections.Counter()
smp = collections.Counter()
lsb = collections.Counter()
for ln, (_, c, sm, ls) in zip(lines, body):
    per[ln] += c
    smp[ln] += sm
    lsb[ln] += ls
print('total executed warp-instr', tot)
acc = 0
for ln, c in sorted(per.items(), key=lambda kv: (kv[0] or ('', 0))):
    if c * 200 >= tot or os.environ.get("LP_ALL"):
        print('%6.2f%%  %9d  %s:%s   samples %5.2f%% (long_sb %5.2f%%)' % (100.0 * c / tot, c, ln[0] if ln else '?', ln[1] if ln else '?', 100.0 * smp[ln] / max(stot, 1), 100.0 * lsb[ln] / max(stot, 1)))
