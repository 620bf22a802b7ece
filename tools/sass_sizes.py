#!/usr/bin/env python3
"""Instruction counts per (sub)function of a cubin disassembly (nvdisasm -c): which code has to fit the instruction cache."""
import re, subprocess, sys
lines = open(sys.argv[1]).read().split('\n')
marks = []
for i, l in enumerate(lines):
  m = re.match(r'^(\$?[_\$A-Za-z0-9\.]+):$', l)
  if m and not l.startswith('.L_') and not l.startswith('.text'):
    marks.append((i, m.group(1)))
def count(a, b): return sum(1 for l in lines[a:b] if re.match(r'^\s+/\*[0-9a-f]{4,}\*/', l))
for (i, n), (j, _) in zip(marks, marks[1:] + [(len(lines), '')]):
  name = n.split('$')[-1] if '$' in n else n
  name = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
  kern = next((k for k in ('narrow', 'solve_big', 'scene_solve_kernel', 'begin', 'reset') if k in n), '?')
  c = count(i, j)
  if c > 60: print(f'{c:6d} {kern:18s} {name[:110]}')
