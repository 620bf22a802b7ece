#!/usr/bin/env python3
"""Builds tests/golden/kat_primitives.blob: the reference's SO100 scene (so101_sim/assets/so100/scene_pbr.xml, read from the
read-only checkout at build time) with TWO PRIMITIVE free props - a box and a capsule carrying the YCB props' collision
class (tests/golden/kat_scene/*.xml) - attached exactly as so100_hand_over.py:159-206 attaches the banana and the bowl.
The analytic known-answer tests (tests/test_oracle_analytic.py, tests/test_scene_gpu.py) run closed-form contact cases on it:
rest penetration from solref / solimp, normal force = m g, sliding deceleration = mu g, contact geometry of box-box and
capsule-box pairs.  The blob is committed because /root/reference does not exist on the GPU box."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import compile_model as cm

def build(ref_root='/root/reference'):
  assets = os.path.join(ref_root, 'so101_sim', 'assets')
  kat = os.path.join(HERE, '..', 'tests', 'golden', 'kat_scene')
  m = cm.Model()
  cm.parse_mjcf(m, os.path.join(assets, 'so100', 'scene_pbr.xml'))
  for nm, f in (('kat_box/', 'kat_box.xml'), ('kat_capsule/', 'kat_capsule.xml')):
    cm.parse_mjcf(m, os.path.join(kat, f), prefix=nm, attach_free=True, mesh_scale=1.0)
  A = cm.finalize(m, {})
  nb = A['nbody']
  A['nprop'] = 2
  A['prop_body'] = np.array([nb - 2, nb - 1], dtype=np.int32)
  A['reward_obj_box'] = cm.body_root_box(A, nb - 2)
  A['reward_box_pos'] = np.zeros(3); A['reward_box_half'] = np.array([0.01, 0.01, 0.01])
  A['prop_mass_standin'] = 0
  return A

if __name__ == '__main__':
  A = build()
  out = os.path.join(HERE, '..', 'tests', 'golden', 'kat_primitives.blob')
  cm.write_blob(out, A)
  print(f"nq={A['nq']} nv={A['nv']} nbody={A['nbody']} ngeom={A['ngeom']} -> {out} ({os.path.getsize(out)} B)")
  for i in (A['nbody'] - 2, A['nbody'] - 1):
    print('body', i, A['_names_body'][i], 'mass', A['body_mass'][i], 'inertia', A['body_inertia'][i], 'invweight0', A['body_invweight0'][i], 'ipos', A['body_ipos'][i])
