"""Pins the CPU oracle against the reference's own known-answer vectors."""
import numpy as np
import pytest

from oracle.oracle import OracleSim

import json
import os

# KAT-1: reference so101_rl.ipynb:219-229 — one env.step from arm qpos = 0 with calibration offsets applied
# (run from the repo root, calibration/red_arm.json:5-40) and action [0,0,0,0,0,0.5].  The fixture is the notebook's
# printed output, extracted by tools/make_golden.py.
_G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
KAT1 = json.load(open(os.path.join(_G, 'kat1_so101_rl.json')))
KAT2 = json.load(open(os.path.join(_G, 'kat2_reset_observation.json')))
KAT1_OFFSETS = KAT1['calibration_offsets']
KAT1_ACTION = KAT1['action']
KAT1_COMMANDED = KAT1['commanded_joints_pos']
KAT1_QPOS = np.array(KAT1['physics_state'][:6])
KAT1_QVEL = np.array(KAT1['physics_state'][20:26])


@pytest.mark.parametrize('model', ['so100_arm', 'so100_handover_banana'])
def test_kat1_arm_state_after_one_control_step(model):
  s = OracleSim(model, collide=False)
  r = s.control_step(KAT1_ACTION, offsets=KAT1_OFFSETS)
  np.testing.assert_allclose(s.ctrl[:6], KAT1_COMMANDED)
  # printed with 9 significant digits in the notebook
  np.testing.assert_allclose(s.qpos[:6], KAT1_QPOS, rtol=2e-9)
  np.testing.assert_allclose(s.qvel[:6], KAT1_QVEL, rtol=2e-9)
  assert r == 0.0


def test_kat1_sensitivity_friction_rows_matter():
  """Dropping the always-on friction-loss rows must break the KAT (shows the pin has teeth)."""
  s = OracleSim('so100_arm', collide=False)
  s.control_step(KAT1_ACTION, offsets=[0] * 6)  # unsaturated actuators -> different state
  assert np.abs(s.qpos[:6] - KAT1_QPOS).max() > 1e-3


def test_kat1_fixture_is_the_notebook_printout():
  assert KAT1_COMMANDED == [28, 42, 18, -21, 1009, -157.5] and KAT1['joints_pos'] == [0.0] * 6 and KAT1['joints_vel'] == []
  assert len(KAT1['physics_state']) == 38 and len(KAT1['delayed_physics_state']) == 38
  np.testing.assert_allclose(KAT1_QPOS[0], 5.85192160e-02, rtol=1e-9)
  # state BEFORE the step (delayed_physics_state): arm at qpos 0, props at rest on the table
  before = np.array(KAT1['delayed_physics_state'])
  assert np.all(before[:6] == 0) and abs(before[8] - 0.4217) < 1e-3 and abs(before[15] - 0.4226) < 1e-3


def test_kat1_props_stay_at_rest_in_the_oracle():
  """The notebook's pre-step state (so101_rl.ipynb:219-229, props resting on the table) must be a rest state of the oracle's
  contact pipeline too.  After the one control step the notebook prints, the oracle's prop state is within 4.4e-6 m /
  5.2e-5 (quaternion components) of the printed post-step state (bars 1e-5 / 1e-4); left alone for 2 s the props move by
  <= 2.7e-5 m and tilt by <= 8.5e-4 (bars 5e-5 / 1.5e-3): MuJoCo's rest pose of the real YCB props is reproduced to ~30 um
  although the oracle's manifold is its own restatement of multiccd and the prop inertias come from stand-in meshes (the rest
  penetration n * imp^2 / (1 - imp) * K * d = g does not depend on the mass, only on the number of support points n)."""
  s = OracleSim('so100_handover_banana', collide=True)
  before = np.array(KAT1['delayed_physics_state'])
  s.set_state(before[:20], before[20:])
  s.control_step(KAT1_ACTION, offsets=KAT1_OFFSETS)
  after = np.array(KAT1['physics_state'])
  np.testing.assert_allclose(s.qpos[:6], after[:6], rtol=2e-9)   # arm: exact pin
  for a0 in (6, 13):
    assert np.abs(s.qpos[a0:a0 + 3] - after[a0:a0 + 3]).max() < 1e-5, s.qpos[a0:a0 + 3] - after[a0:a0 + 3]
    assert np.abs(s.qpos[a0 + 3:a0 + 7] - after[a0 + 3:a0 + 7]).max() < 1e-4
  s = OracleSim('so100_handover_banana', collide=True)
  s.set_state(before[:20], before[20:])
  for _ in range(100):
    s.control_step(np.zeros(6))
  for a0 in (6, 13):
    assert np.abs(s.qpos[a0:a0 + 3] - before[a0:a0 + 3]).max() < 5e-5, s.qpos[a0:a0 + 3] - before[a0:a0 + 3]
    assert np.abs(s.qpos[a0 + 3:a0 + 7] - before[a0 + 3:a0 + 7]).max() < 1.5e-3
  assert np.abs(s.qvel[6:]).max() < 1e-3


def test_kat2_tilted_bowl_is_a_rest_state_of_the_oracle():
  """KAT-2 (examples/so101_rl_breakdown.ipynb:274-298): the reset observation has the bowl resting TILTED on the static cylinder
  obstacle (z = 0.43058764, quaternion (0.99953667, 0.02739444, -0.01326563, ...)) - a contact configuration with a curved
  primitive (hull vs cylinder through GJK / EPA) next to table contacts.  Held with the home command for 2 s the oracle keeps
  the bowl there: position within 1.2e-4 m, quaternion within 3.6e-4 (bars 2e-4 / 6e-4); the banana within 2.7e-5 / 8e-4."""
  p = np.array(KAT2['physics_state'])
  s = OracleSim('so100_handover_banana', collide=True)
  s.set_state(p[:20], p[20:]); s.forward()
  pairs = {(c['geom1'], c['geom2']) for c in s.contacts()}
  cyl = [g for g in range(len(s.meta['geom_type'])) if s.meta['geom_type'][g] == 3 and abs(s.meta['geom_size'].reshape(-1, 3)[g][0] - 0.08) < 1e-9]
  assert len(cyl) == 1 and any(g1 == cyl[0] for g1, _ in pairs), 'the tilted bowl must touch the static cylinder'
  for _ in range(100):
    s.control_step(np.array(KAT2['commanded_joints_pos']))
  assert np.abs(s.qpos[13:16] - p[13:16]).max() < 2e-4 and np.abs(s.qpos[16:20] - p[16:20]).max() < 6e-4, s.qpos[13:20] - p[13:20]
  assert np.abs(s.qpos[6:9] - p[6:9]).max() < 5e-5 and np.abs(s.qpos[9:13] - p[9:13]).max() < 1.5e-3
  assert abs(s.qpos[15] - 0.43058764) < 2e-4   # still resting on the obstacle, not on the table (0.4226)


def test_kat2_reset_observation_semantics():
  """examples/so101_rl_breakdown.ipynb:274-298 — zero offsets: commanded = HOME_CTRL, arm qpos = 0, joints_vel empty."""
  from so101_sim_b200.task_suite import OBSERVATION_KEYS, SO100_HOME_CTRL
  np.testing.assert_allclose(KAT2['commanded_joints_pos'], SO100_HOME_CTRL, atol=1e-5)
  assert KAT2['joints_pos'] == [0.0] * 6 and KAT2['joints_vel'] == [] and KAT2['undelayed_joints_vel'] == []
  assert KAT2['physics_state'] == KAT2['delayed_physics_state'] and KAT2['physics_state'][:6] == [0.0] * 6
  assert tuple(k for k in KAT2['observation_keys'] if not k.endswith('_cam')) == OBSERVATION_KEYS
  # printed action bounds (2 decimals): +-3.14 for the five arm joints, [0, 0.08] for the jaw
  assert KAT2['action_spec']['minimum'] == [-3.14] * 5 + [0.0] and KAT2['action_spec']['maximum'] == [3.14] * 5 + [0.08]


def test_overlap_matches_the_reference_oobb_utils_golden():
  """tests/golden/oobb_overlap.json was produced by executing the reference's own so101_sim/utils/oobb_utils.py
  (transform_oobb + overlap_oobb_oobb, the 6-axis SAT with strict comparisons) on 240 seeded box pairs built around the
  task's overlap boxes (tools/make_golden_oobb.py).  The oracle's restatement must return the same booleans, and its
  quaternion transform must reproduce the reference's world-space container boxes."""
  import json
  from oracle.oracle import overlap_oobb_oobb
  g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'oobb_overlap.json')))
  assert len(g['cases']) == 240
  n_true = 0
  for c in g['cases']:
    got = overlap_oobb_oobb(c['obj_pos'], c['obj_quat'], c['obj_half'], c['ws_pos'], c['ws_quat'], c['container_half'])
    assert got == c['overlap'], c
    n_true += got
    # transform_oobb (oobb_utils.py:175-199): position = body_pos + R(body_quat) container_pos, rotation = body_quat (x identity)
    w, x, y, z = c['body_quat']
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    np.testing.assert_allclose(np.asarray(c['body_pos']) + R @ np.asarray(c['container_pos']), c['ws_pos'], atol=1e-12)
    np.testing.assert_allclose(c['ws_quat'], c['body_quat'], atol=1e-15)
  assert 60 < n_true < 180  # both outcomes are well represented
