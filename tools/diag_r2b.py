#!/usr/bin/env python3
"""Diagnostics (round 2): (1) per-env / per-component error of the f32 and f64 scene paths against the oracle over 25 steps;
(2) banana-in-bowl drop at num_envs = 1 and 4: per-step error, contact counts, drop counters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.oracle import OracleSim
from so101_sim_b200.task_suite import create_batched_task_env
dev = 'cuda:0'
def actions(env, steps, seed=1, scale=0.3):
  g = torch.Generator(device=dev); g.manual_seed(seed)
  spec = env.action_spec()
  lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
  return (lo + torch.rand(steps, env.num_envs, 6, generator=g, device=dev) * (hi - lo)) * scale

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', 'scene'):
  for prec in ('f32', 'f64'):
    env = create_batched_task_env('SO100HandOverBanana', num_envs=4, time_limit=30.0, seed=0, device=dev, precision=prec, reset_rounds=0)
    env.sample_prop_initial_states(seed=5, clearance=0.002, settle_steps=0)
    q0, v0 = env.get_state(torch.float64)
    acts = actions(env, 25, scale=0.1)
    sims = []
    for e in range(4):
      o = OracleSim('so100_handover_banana', collide=True); o.set_state(q0[e].cpu().numpy(), v0[e].cpu().numpy()); sims.append(o)
    for t in range(25):
      env.step(acts[t])
      q, v = env.get_state(torch.float64); q = q.cpu().numpy()
      ncon = env.debug_read('ncon').flatten().tolist()
      row = []
      for e, o in enumerate(sims):
        o.control_step(acts[t, e].double().cpu().numpy())
        d = np.abs(q[e] - o.qpos)
        row.append('e%d arm %.1e ban p %.1e q %.1e bowl p %.1e q %.1e nc %d/%d' % (e, d[:6].max(), d[6:9].max(), d[9:13].max(), d[13:16].max(), d[16:20].max(), int(ncon[e]), o.info('ncon')))
      if t in (0, 1, 2, 4, 9, 24):
        print(prec, 'step', t + 1, ' | '.join(row))
    print(prec, env.counters())
    env.close()
if which in ('all', 'drop'):
  for N in (1, 4):
    env = create_batched_task_env('SO100HandOverBanana', num_envs=N, time_limit=30.0, seed=0, device=dev, precision='f64', reset_rounds=0)
    q = torch.tensor(env.model['qpos0'], dtype=torch.float64).repeat(N, 1)
    q[:, :6] = 0
    q[:, 13:16] = torch.tensor([-0.25, -0.05, 0.4226], dtype=torch.float64); q[:, 16] = 1; q[:, 17:20] = 0
    q[:, 6:9] = torch.tensor([-0.25 - 0.0255, -0.05 - 0.0675, 0.4226 + 0.06], dtype=torch.float64)
    q[:, 9] = float(np.cos(0.4)); q[:, 10:12] = 0; q[:, 12] = float(np.sin(0.4))
    env.set_initial_state(q, torch.zeros(N, 18, dtype=torch.float64)); env.reset()
    o = OracleSim('so100_handover_banana', collide=True); o.set_state(q[0].numpy(), np.zeros(18))
    zero = torch.zeros(N, 6, device=dev)
    for t in range(12):
      env.step(zero); o.control_step(np.zeros(6))
      qq, vv = env.get_state(torch.float64); qq = qq.cpu().numpy()
      d = np.abs(qq[0] - o.qpos)
      print('drop N=%d step %d: err arm %.1e ban %.1e bowl %.1e  ncon gpu %d oracle(refreshed) %d  ban z %.4f oracle %.4f' % (N, t + 1, d[:6].max(), d[6:13].max(), d[13:].max(), int(env.debug_read('ncon')[0, 0]), o.info('ncon'), qq[0, 8], o.qpos[8]), env.counters())
    env.close()
