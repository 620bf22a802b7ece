"""Pins the CPU oracle against the reference's own known-answer vectors."""
import numpy as np
import pytest

from oracle.oracle import OracleSim

# KAT-1: reference so101_rl.ipynb:219-229 — one env.step from arm qpos = 0 with calibration offsets applied
# (run from the repo root, calibration/red_arm.json:5-40) and action [0,0,0,0,0,0.5].
KAT1_OFFSETS = [28, 42, 18, -21, 1009, -158]
KAT1_ACTION = [0, 0, 0, 0, 0, 0.5]
KAT1_COMMANDED = [28, 42, 18, -21, 1009, -157.5]
KAT1_QPOS = np.array([5.85192160e-02, 5.80983147e-02, 6.58658498e-02, -8.00624348e-02, 7.67682376e-02, -7.65953670e-02])
KAT1_QVEL = np.array([5.32876236, 5.27619008, 5.98870371, -7.27788632, 6.97910317, -6.96330033])


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
