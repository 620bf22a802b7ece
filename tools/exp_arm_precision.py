#!/usr/bin/env python3
"""Error-budget experiment for the float32 arm path: which parts must run in float64 so that a 100-step
rollout stays within 1e-4 relative of the float64 oracle.  SO101_ARM_MODE: 0 = state rounded to float32 every substep (round 1),
1 = float64 state + Euler, 2 = + float64 actuator model (product default), 3 = + float64 M^-1 solve, 4 = + float64 FK/CRB/RNE.
Prints one JSON line per mode: max relative error over 32 envs at control steps 10/25/50/100 and the 4096-env step time."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run(mode):
  import numpy as np, torch
  from oracle.oracle import OracleSim
  from so101_sim_b200.task_suite import create_batched_task_env
  dev = 'cuda:0'
  env = create_batched_task_env('SO100ArmOnly', num_envs=32, time_limit=30.0, seed=0, device=dev)
  q0, _ = env.sample_arm_initial_states(seed=0)
  env.reset()
  g = torch.Generator(device=dev); g.manual_seed(1)
  spec = env.action_spec()
  lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
  acts = (lo + torch.rand(100, 32, 6, generator=g, device=dev) * (hi - lo)) * 0.3
  sims = []
  for e in range(32):
    o = OracleSim('so100_arm', collide=False); o.set_state(q0[e].double().cpu().numpy(), np.zeros(6)); sims.append(o)
  out = {}
  for t in range(100):
    env.step(acts[t])
    for e, o in enumerate(sims):
      o.control_step(acts[t, e].double().cpu().numpy())
    if t + 1 in (10, 25, 50, 100):
      q, v = env.get_state(torch.float64)
      q, v = q.cpu().numpy(), v.cpu().numpy()
      eq = max(np.abs(q[e] - o.qpos).max() / max(1, np.abs(o.qpos).max()) for e, o in enumerate(sims))
      ev = max(np.abs(v[e] - o.qvel).max() / max(1, np.abs(o.qvel).max()) for e, o in enumerate(sims))
      out[t + 1] = (float(f'{eq:.3g}'), float(f'{ev:.3g}'))
  env.close()
  env = create_batched_task_env('SO100ArmOnly', num_envs=4096, time_limit=30.0, seed=0, device=dev)
  env.sample_arm_initial_states(seed=0); env.reset()
  a = torch.zeros(4096, 6, device=dev)
  for _ in range(5): env.step(a)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(50): env.step(a)
  e1.record(); torch.cuda.synchronize()
  print(json.dumps(dict(arm_mode=mode, rel_err_qpos_qvel_by_step=out, us_per_step_4096=1e3 * e0.elapsed_time(e1) / 50)))

if __name__ == '__main__':
  if len(sys.argv) > 1:
    run(int(sys.argv[1]))
  else:
    for m in range(5):
      subprocess.run([sys.executable, __file__, str(m)], env=dict(os.environ, SO101_ARM_MODE=str(m)))
