#!/usr/bin/env python3
"""Substep-by-substep comparison of the float64 GPU scene kernel with the float64 oracle (debug aid, needs a GPU)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np, torch
from oracle.oracle import OracleSim
from so101_sim_b200.task_suite import create_batched_task_env

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prec = sys.argv[4] if len(sys.argv) > 4 else 'f64'
env = create_batched_task_env('SO100HandOverBanana', num_envs=N, time_limit=30.0, control_timestep=0.002, precision=prec)
env.sample_prop_initial_states(seed=seed, clearance=float(os.environ.get('SO101_DBG_CLEARANCE', '0.002')), settle_steps=0)
q0, v0 = env.get_state(torch.float64)
env.debug_contacts()
sims = []
for e in range(N):
  o = OracleSim('so100_handover_banana', collide=True); o.set_state(q0[e].cpu().numpy(), v0[e].cpu().numpy()); sims.append(o)
g = torch.Generator(device='cpu'); g.manual_seed(1)
bad = set()
for t in range(steps):
  a = ((torch.rand(N, 6, generator=g) * 2 - 1) * 0.5).to('cuda:0')
  for e in range(N):
    sims[e].ctrl[:] = a[e].double().cpu().numpy()
  # oracle contacts BEFORE the substep are the ones used inside it: run substep and grab the list from the position stage
  env.step(a)
  q, v = env.get_state(torch.float64)
  gc = env.debug_contacts()
  it = env.debug_read('solver_iter')[:, 0].cpu().numpy()
  for e in range(N):
    if os.environ.get('SO101_DBG_ENV') and e == int(os.environ['SO101_DBG_ENV']) and t == int(os.environ['SO101_DBG_STEP']):
      os.environ['SO101_ORACLE_DBG'] = os.environ.get('SO101_DBG_PAIR', '19,30'); sys.stdout.flush()
    else:
      os.environ.pop('SO101_ORACLE_DBG', None)
    sims[e].substep()
    oc = sims[e].contacts()
    dq = np.abs(q[e].cpu().numpy() - sims[e].qpos).max(); dv = np.abs(v[e].cpu().numpy() - sims[e].qvel).max()
    tol = 1e-9 if prec == 'f64' else 1e-4
    if (dq > tol or len(gc[e]) != len(oc)) and e not in bad:
      bad.add(e)
      print(f'--- env {e} substep {t}: dq={dq:.3e} dv={dv:.3e} ncon gpu={len(gc[e])} oracle={len(oc)} iters gpu={it[e]} oracle={sims[e].info("solver_iter")}')
      print('   q0 props', q0[e, 6:].cpu().numpy().round(4))
      go = {}; oo = {}
      for c in gc[e]: go.setdefault((c[0], c[1]), []).append(c)
      for c in oc: oo.setdefault((c['geom1'], c['geom2']), []).append(c)
      for k in sorted(set(go) | set(oo)):
        a_, b_ = go.get(k, []), oo.get(k, [])
        same = len(a_) == len(b_) and all(abs(x[2] - y['dist']) < 1e-6 and np.abs(x[3] - y['pos']).max() < 1e-5 for x, y in zip(a_, b_))
        if not same:
          print('   pair', k, 'gpu', [(round(x[2], 7), x[3].round(5).tolist(), x[4].round(4).tolist()) for x in a_])
          print('        ', ' ', 'ora', [(round(y['dist'], 7), y['pos'].round(5).tolist(), y['frame'][0].round(4).tolist()) for y in b_])
print('diverged envs:', sorted(bad), 'counters', env.counters())
mx = max(np.abs(env.get_state(torch.float64)[0][e].cpu().numpy() - sims[e].qpos).max() for e in range(N))
print('final max |dq| =', mx)
