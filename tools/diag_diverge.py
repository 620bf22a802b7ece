#!/usr/bin/env python3
"""Diagnostics: how often does the float32 path end an episode by divergence under the bench's random actions, and does float64
(the same kernels in double, and the CPU oracle) diverge from the same pre-step state?  Usage: diag_diverge.py [envs] [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so101_sim_b200.task_suite import create_batched_task_env
dev = 'cuda:0'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 110

def actions(env, n, envs, seed):
  g = torch.Generator(device=dev); g.manual_seed(seed)
  spec = env.action_spec()
  lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
  return (lo + torch.rand(n, envs, len(spec.minimum), generator=g, device=dev) * (hi - lo)) * 0.3

os.makedirs('gpurun_out', exist_ok=True)
env32 = create_batched_task_env('SO100HandOverBanana', num_envs=N, time_limit=30.0, seed=0, device=dev, precision='f32', nursery_envs=0)
q0, v0 = env32.get_state(torch.float64)
acts = actions(env32, STEPS, N, 1234)
events = {}
out = {}
for prec in ('f32', 'f64'):
  env = env32 if prec == 'f32' else create_batched_task_env('SO100HandOverBanana', num_envs=N, time_limit=30.0, seed=0, device=dev, precision='f64', placement='none')
  env.set_initial_state(q0, v0); env.reset()
  ev = []
  alive = torch.ones(N, dtype=torch.bool, device=dev)   # only the first episode of every env counts (same actions, same start in both)
  for t in range(STEPS):
    q, v = env.get_state(torch.float64)
    ts = env.step(acts[t])
    div = (ts.step_type == 2) & (ts.discount == 0) & (ts.reward == 0) & alive
    ended = (ts.step_type == 2) & alive
    for e in torch.nonzero(div).flatten().tolist():
      ev.append(dict(step=t, env=e, q=q[e].cpu().numpy(), v=v[e].cpu().numpy(), a=acts[t, e].cpu().numpy(), ncon=int(env.debug_read('ncon').flatten()[e])))
    alive &= ~ended
  events[prec] = ev
  out[prec] = dict(diverged_first_episode=len(ev), counters=env.counters())
  print(prec, 'diverged', len(ev), 'of', N, 'in', STEPS, 'steps;', env.counters(), flush=True)
  if prec == 'f64': env.close()
# whole histories (start state + actions) of the first float64 divergences, for a substep-by-substep look in the CPU oracle
sel = [e['env'] for e in events['f64'][:24]]
np.savez('gpurun_out/diag_diverge_histories.npz', envs=np.array(sel), q0=q0[sel].cpu().numpy(), v0=v0[sel].cpu().numpy(),
         acts=acts[:, sel].cpu().numpy(), steps=np.array([e['step'] for e in events['f64'][:24]]))
both = set(e['env'] for e in events['f32']) & set(e['env'] for e in events['f64'])
print('envs diverged in both precisions:', len(both))
# the pre-step state of the float32 divergences, replayed for one control step in float64 on the device and in the CPU oracle
ev = events['f32'][:64]
if ev:
  from oracle.oracle import OracleSim
  envr = create_batched_task_env('SO100HandOverBanana', num_envs=len(ev), time_limit=30.0, seed=0, device=dev, precision='f64', placement='none')
  Q = torch.tensor(np.stack([e['q'] for e in ev])); V = torch.tensor(np.stack([e['v'] for e in ev]))
  envr.set_initial_state(Q, V); envr.reset()
  A = torch.tensor(np.stack([e['a'] for e in ev]), device=dev)
  ts = envr.step(A)
  d64 = ((ts.step_type == 2) & (ts.discount == 0)).cpu().numpy()
  nor = 0
  rows = []
  for i, e in enumerate(ev):
    o = OracleSim('so100_handover_banana', collide=True); o.set_state(e['q'], e['v'])
    o.forward()
    depth = min([c['dist'] for c in o.contacts()] + [0.0])
    o.control_step(e['a'].astype(np.float64))
    od = bool(o.info('diverged'))
    nor += od
    rows.append(dict(step=e['step'], env=e['env'], ncon=e['ncon'], max_qvel_arm=float(np.abs(e['v'][:6]).max()), max_qvel_prop=float(np.abs(e['v'][6:]).max()),
                     min_dist_pre=float(depth), f64_device_diverges=bool(d64[i]), oracle_diverges=od,
                     prop_z=[float(e['q'][8]), float(e['q'][15])]))
  print('replayed', len(ev), 'float32 divergences from their pre-step state: float64 device diverges', int(d64.sum()), ', oracle diverges', nor)
  for r in rows[:24]: print(json.dumps(r))
  out['replay'] = rows
json.dump(out, open('gpurun_out/diag_diverge.json', 'w'), indent=1, default=float)
