"""Host-side mirror of the reference's calibration helper (scripts/so101_calibration.py:12-142).

The reference reads `calibration/red_arm.json` relative to the CWD and silently falls back to zero offsets
(so101_calibration.py:15,39-56); whether the offsets apply therefore depends on where the process was started
(SURVEY.md fact 7).  Here the 6-vector is an explicit parameter: pass a JSON path to get the reference's behaviour
when run from its repo root, or nothing for zero offsets.
"""
from __future__ import annotations

import json
import os

import numpy as np

# so101_calibration.py:23-31
JOINT_MAPPING = {'shoulder_pan': 0, 'shoulder_lift': 1, 'elbow_flex': 2, 'wrist_flex': 3, 'wrist_roll': 4, 'gripper': 5}


class SO101Calibration:
  def __init__(self, calibration_file: str | None = None):
    self.calibration_file = calibration_file
    self.calibration_data = {}
    self.homing_offsets = np.zeros(6)
    self.load_calibration()

  def load_calibration(self) -> bool:
    """so101_calibration.py:36-56 — missing/unreadable file leaves zero offsets and returns False."""
    try:
      if not self.calibration_file or not os.path.exists(self.calibration_file):
        return False
      with open(self.calibration_file) as f:
        self.calibration_data = json.load(f)
      for name, idx in JOINT_MAPPING.items():
        if name in self.calibration_data:
          self.homing_offsets[idx] = self.calibration_data[name].get('homing_offset', 0)
      return True
    except Exception:
      return False

  def apply_calibration_to_position(self, joint_positions: np.ndarray) -> np.ndarray:
    """so101_calibration.py:62-77"""
    joint_positions = np.asarray(joint_positions)
    if len(joint_positions) != 6:
      raise ValueError(f'Expected 6 joint positions, got {len(joint_positions)}')
    return joint_positions + self.homing_offsets

  def apply_calibration_to_action(self, action: np.ndarray) -> np.ndarray:
    return self.apply_calibration_to_position(action)
