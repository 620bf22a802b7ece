#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the env-step path into tests/golden/*.json.

The reference (read-only checkout, argument 1, default /root/reference) pins nothing numerically in its tests for the
SO100 path; the only numeric pins are printed cell outputs of two notebooks:
  so101_rl.ipynb                      -> KAT-1: observation after ONE env.step (state before = delayed_physics_state)
  examples/so101_rl_breakdown.ipynb   -> KAT-2: reset observation, observation keys, action spec bounds
This script parses those printed outputs (it executes nothing of the reference) and writes small JSON fixtures that travel
with the repo; `/root/reference` does not exist on the GPU box.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
NUM = r'[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?'


def outputs(nb_path):
  try:
    nb = json.load(open(nb_path))
  except json.JSONDecodeError:
    # examples/so101_rl_breakdown.ipynb is not valid JSON as committed: fall back to its raw text (JSON string escapes undone)
    raw = open(nb_path).read().replace('\\n",\n', '\n').replace('\\n', '\n').replace('\\"', '"')
    raw = re.sub(r'\n\s*"', '\n', raw)
    yield -1, raw
    return
  for ci, c in enumerate(nb['cells']):
    for o in c.get('outputs', []):
      t = ''.join(o.get('text', [])) if 'text' in o else ''.join(o.get('data', {}).get('text/plain', []))
      if t:
        yield ci, t


def arrays(text, names):
  """'name': array([...]) occurrences -> {name: [floats]} (first occurrence of each name)."""
  got = {}
  for n in names:
    m = re.search(r"'%s':\s*array\(\[(.*?)\]" % re.escape(n), text, re.S)
    if m:
      got[n] = [float(x) for x in re.findall(NUM, m.group(1))]
  return got


def main():
  os.makedirs(OUT, exist_ok=True)
  names = ['commanded_joints_pos', 'joints_pos', 'joints_vel', 'physics_state', 'undelayed_joints_pos', 'undelayed_joints_vel',
           'delayed_physics_state']
  # ---- KAT-1
  src = os.path.join(REF, 'so101_rl.ipynb')
  best = None
  for ci, t in outputs(src):
    a = arrays(t, names)
    if 'physics_state' in a and len(a['physics_state']) == 38 and any(abs(x) > 1 for x in a['physics_state'][20:26]):
      best = (ci, a)
  assert best, 'KAT-1 output not found'
  ci, a = best
  kat1 = dict(source='so101_rl.ipynb cell %d (printed observation after one env.step; run from the repo root so that '
                     'calibration/red_arm.json homing offsets are applied)' % ci,
              calibration_offsets=[28, 42, 18, -21, 1009, -158], action=[0, 0, 0, 0, 0, 0.5], **a)
  json.dump(kat1, open(os.path.join(OUT, 'kat1_so101_rl.json'), 'w'), indent=1)
  # ---- KAT-2
  src = os.path.join(REF, 'examples', 'so101_rl_breakdown.ipynb')
  kat2 = None
  keys = None
  spec = {}
  for ci, t in outputs(src):
    a = arrays(t, names)
    if kat2 is None and 'physics_state' in a and len(a['physics_state']) == 38 and 'commanded_joints_pos' in a:
      kat2 = dict(source='examples/so101_rl_breakdown.ipynb cell %d (reset observation, zero calibration offsets)' % ci, **a)
    m = re.search(r"odict_keys\(\[(.*?)\]\)|Observation keys:\s*\[(.*?)\]", t, re.S)
    if m and keys is None:
      keys = re.findall(r"'([a-z_]+)'", m.group(1) or m.group(2))
    rng = re.findall(r"action\[(\d)\] = (\w+)\s*\| range: \[\s*(%s),\s*(%s)\]" % (NUM, NUM), t)
    if rng and not spec:  # printed with 2 decimals
      spec = dict(names=[r[1] for r in rng], minimum=[float(r[2]) for r in rng], maximum=[float(r[3]) for r in rng], printed_decimals=2)
  assert kat2, 'KAT-2 output not found'
  kat2['observation_keys'] = keys
  kat2['action_spec'] = spec
  json.dump(kat2, open(os.path.join(OUT, 'kat2_reset_observation.json'), 'w'), indent=1)
  print('wrote', os.listdir(OUT))


if __name__ == '__main__':
  main()
