#!/usr/bin/env python3
"""T3 parity hook: compare the float64 oracle with the REAL reference stack when a box has it.

Needs `mujoco`, `dm_control` and the reference checkout (argument 1, default /root/reference).  In the build container
none of them is importable, so this prints a skip line and exits 0.  It runs the reference's own
`task_suite.create_task_env('SO100HandOverBanana', cameras=())` for a few control steps from its reset state and replays the
same state / actions through the oracle, printing the arm and prop deviations per step.
"""
import os
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
try:
  import mujoco  # noqa: F401
  import dm_control  # noqa: F401
except Exception as e:  # pragma: no cover
  print(f'skipped: reference stack not importable here ({e})')
  sys.exit(0)
if not os.path.isdir(ref):
  print(f'skipped: reference checkout {ref} not present')
  sys.exit(0)
sys.path.insert(0, ref); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from so101_sim import task_suite
from oracle.oracle import OracleSim

env = task_suite.create_task_env('SO100HandOverBanana', time_limit=30.0, cameras=(), random_state=np.random.RandomState(0))
ts = env.reset()
state = ts.observation['physics_state']
o = OracleSim('so100_handover_banana', collide=True)
o.set_state(state[:20], state[20:])
rs = np.random.RandomState(1)
spec = env.action_spec()
for t in range(20):
  a = rs.uniform(spec.minimum, spec.maximum) * 0.3
  ts = env.step(a)
  r = o.control_step(a.astype(np.float64))
  s = ts.observation['physics_state']
  print(f'step {t}: |dq arm| {np.abs(s[:6] - o.qpos[:6]).max():.3e}  |dq props| {np.abs(s[6:20] - o.qpos[6:]).max():.3e}  reward ref {ts.reward} oracle {r}')
