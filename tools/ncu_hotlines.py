#!/usr/bin/env python3
"""Attribute an ncu --set full --import-source capture of one kernel to device functions and source lines.
usage: python tools/ncu_hotlines.py <report.ncu-rep> <kernel-substring> [object.o] [top-N]
Reads the SASS page of the report (stall samples, executed instructions, active lanes per instruction) and maps it to the
function labels / line info of the same kernel in the in-tree object (nvdisasm -c -g): the build must match the capture."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
obj = sys.argv[3] if len(sys.argv) > 3 else 'so101_sim_b200/csrc/_obj/scene_kernel_f32.o'
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout.split('\n')
rows = list(csv.reader(out[1:]))
hdr = rows[0]
isamp, iex, ith = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
body = []
for r in rows[1:]:   # a report with several launches repeats the header per launch: the first launch is attributed
  if len(r) > ith and r[isamp] == '# Samples':
    break
  if len(r) > ith and r[isamp].strip().isdigit():
    body.append(r)
data = [(int(r[isamp]), int(r[iex]), int(r[ith])) for r in body]
with tempfile.TemporaryDirectory() as td:
  subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=td, capture_output=True)
  cub = [f for f in os.listdir(td) if f.endswith('.cubin')][0]
  lines = subprocess.run(['nvdisasm', '-c', '-g', os.path.join(td, cub)], capture_output=True, text=True).stdout.split('\n')
secs = [i for i, l in enumerate(lines) if l.startswith('//--------------------- .text.')]
a = b = None
for k, i in enumerate(secs):
  if kern in lines[i]:
    a, b = i, (secs[k + 1] if k + 1 < len(secs) else len(lines))
    break
cur, fn, per = None, '?', []
for l in lines[a:b]:
  m = re.search(r'//## File "([^"]+)", line (\d+)', l)
  if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
  m = re.match(r'^(\$?[_\$A-Za-z0-9\.]+):$', l)
  if m and not l.startswith('.L_') and not l.startswith('.text'): fn = m.group(1)
  if re.match(r'^\s+/\*[0-9a-f]{4,}\*/', l): per.append((fn, cur))
assert len(per) == len(data), (len(per), len(data), 'the object does not match the capture')
tot, totex = sum(d[0] for d in data), sum(d[1] for d in data)
byfn, byline = collections.defaultdict(lambda: [0, 0, 0, 0]), collections.defaultdict(lambda: [0, 0, 0])
for (f, ln), d in zip(per, data):
  byfn[f][0] += d[0]; byfn[f][1] += d[1]; byfn[f][2] += d[2]; byfn[f][3] += 1
  byline[ln][0] += d[0]; byline[ln][1] += d[1]; byline[ln][2] += d[2]
print(f'{len(data)} instructions, {tot} samples, {totex / 1e6:.1f} M warp instructions executed, {sum(d[2] for d in data) / totex:.1f} active lanes')
for f, v in sorted(byfn.items(), key=lambda kv: -kv[1][0]):
  name = subprocess.run(['c++filt', f.split('$')[-1] if '$' in f else f], capture_output=True, text=True).stdout.strip()
  name = re.sub(r'<.*', '', name.replace('so101::', ''))
  print(f'{v[3]:5d} instr  samples {100 * v[0] / tot:5.1f}%  executed {100 * v[1] / totex:5.1f}%  lanes {v[2] / max(v[1], 1):5.1f}  {name[:60]}')
src = {}
for (f, ln), v in sorted(((k, v) for k, v in byline.items() if k), key=lambda kv: -kv[1][0])[:top]:
  if f not in src:
    p = os.path.join('so101_sim_b200/csrc', f)
    src[f] = open(p).read().split('\n') if os.path.exists(p) else None
  text = src[f][ln - 1].strip()[:95] if src[f] and ln <= len(src[f]) else ''
  print(f'{100 * v[0] / tot:5.1f}%  {v[1] / 1e6:6.1f} M  lanes {v[2] / max(v[1], 1):4.1f}  {f}:{ln}  {text}')
