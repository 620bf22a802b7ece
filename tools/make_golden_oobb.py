#!/usr/bin/env python3
"""Golden vectors for the reward geometry, produced by RUNNING the reference's own code:
so101_sim/utils/oobb_utils.py (transform_oobb :175-199, overlap_oobb_oobb :251-273, overlap_aabb_oobb :202-248) is imported
from the read-only checkout and evaluated on seeded random box pairs -> tests/golden/oobb_overlap.json.

oobb_utils imports `mujoco` for four quaternion helpers (mju_rotVecQuat, mju_mulQuat, mju_negQuat, mju_mat2Quat); the package
is not installed here, so this script registers a stand-in module with numpy versions of exactly those four documented
functions ([upstream] engine_util_spatial.c semantics: quaternions are (w, x, y, z), results are written in place).  Everything
else - corner enumeration, the 6-axis separating-axis test, its strict comparisons - is the reference's code.
`/root/reference` does not exist on the GPU box: only the JSON travels.
usage: python tools/make_golden_oobb.py [/root/reference]"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'oobb_overlap.json')


def _mul(a, b):
  return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                   a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def mju_mulQuat(res, q1, q2):
  res[:] = _mul(np.asarray(q1, dtype=float), np.asarray(q2, dtype=float))


def mju_negQuat(res, q):
  res[:] = [q[0], -q[1], -q[2], -q[3]]


def mju_rotVecQuat(res, vec, quat):
  q = np.asarray(quat, dtype=float)
  v = np.array([0.0, vec[0], vec[1], vec[2]])
  res[:] = _mul(_mul(q, v), np.array([q[0], -q[1], -q[2], -q[3]]))[1:]


def mju_mat2Quat(quat, mat):  # not exercised by the functions sampled here; present so that the import succeeds
  raise NotImplementedError


stub = types.ModuleType('mujoco')
stub.MjModel = stub.MjData = object  # (type annotations of get_oobb only)
stub.mju_mulQuat, stub.mju_negQuat, stub.mju_rotVecQuat, stub.mju_mat2Quat = mju_mulQuat, mju_negQuat, mju_rotVecQuat, mju_mat2Quat
sys.modules['mujoco'] = stub
sys.modules['ref_oobb_utils'] = None
spec = importlib.util.spec_from_file_location('ref_oobb_utils', os.path.join(REF, 'so101_sim', 'utils', 'oobb_utils.py'))
ou = importlib.util.module_from_spec(spec)
sys.modules['ref_oobb_utils'] = ou  # (dataclasses looks the module up while the classes are created)
spec.loader.exec_module(ou)


def rquat(rs, small=False):
  if small:
    ax = rs.normal(size=3); ax /= np.linalg.norm(ax)
    a = rs.uniform(-0.3, 0.3)
    return np.concatenate([[np.cos(a / 2)], np.sin(a / 2) * ax])
  q = rs.normal(size=4)
  return q / np.linalg.norm(q)


def main():
  rs = np.random.RandomState(20251017)
  cases = []
  # the task's own boxes (so100_hand_over.py:87-93,104-116) and the compiled banana / pen root boxes, plus random ones
  containers = [((-0.0255, -0.0675, 0.0525), (0.03, 0.03, 0.015)), ((0.0, 0.0, 0.015996), (0.027996, 0.027996, 0.015)),
                ((0.0, 0.0, 0.15), (0.06, 0.06, 0.009996))]
  objects = [(0.0323, 0.0439, 0.1062), (0.00803, 0.00908, 0.09279)]
  for k in range(240):
    cpos, chalf = containers[k % 3] if k < 180 else (rs.uniform(-0.1, 0.1, 3), rs.uniform(0.005, 0.08, 3))
    ohalf = objects[k % 2] if k < 180 else rs.uniform(0.005, 0.1, 3)
    body_pos = rs.uniform(-0.3, 0.3, 3) + [0, 0, 0.42]
    body_quat = rquat(rs, small=k % 4 != 0)
    box = ou.Oobb(position=np.array(cpos, dtype=float), rotation=np.array([1.0, 0, 0, 0]), half_extents=np.array(chalf, dtype=float))
    ws = ou.transform_oobb(box, body_pos, body_quat)
    # object box: centred near the container box (so that both outcomes occur), with a spread that grows over the cases
    spread = 0.02 + 0.12 * rs.uniform()
    opos = ws.position + rs.normal(size=3) * spread
    oq = rquat(rs, small=False)
    obj = ou.Oobb(position=opos, rotation=oq, half_extents=np.array(ohalf, dtype=float))
    res = bool(ou.overlap_oobb_oobb(obj, ws))
    cases.append(dict(container_pos=list(map(float, cpos)), container_half=list(map(float, chalf)), body_pos=body_pos.tolist(),
                      body_quat=body_quat.tolist(), ws_pos=ws.position.tolist(), ws_quat=ws.rotation.tolist(),
                      obj_pos=opos.tolist(), obj_quat=oq.tolist(), obj_half=list(map(float, ohalf)), overlap=res))
  n_true = sum(c['overlap'] for c in cases)
  json.dump(dict(source='so101_sim/utils/oobb_utils.py transform_oobb + overlap_oobb_oobb, executed by tools/make_golden_oobb.py',
                 seed=20251017, cases=cases), open(OUT, 'w'))
  print(f'{len(cases)} cases, {n_true} overlapping -> {OUT} ({os.path.getsize(OUT)} B)')


if __name__ == '__main__':
  main()
