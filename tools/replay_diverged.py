#!/usr/bin/env python3
"""Developer probe: replay saved diverging envs (tools/data/diverged.npz) on the GPU in f32 and f64 next to the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.oracle import OracleSim
from so101_sim_b200.task_suite import create_batched_task_env
d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'diverged.npz'))
want = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [739]
ks = [i for i, e in enumerate(d['env']) if int(e) in want]
for k in ks:
  T = int(d['end_step'][k]) + 2
  sims = {}
  for prec in ('f32', 'f64'):
    env = create_batched_task_env('SO100HandOverBanana', num_envs=2, time_limit=30.0, seed=0, device='cuda:0', precision=prec)
    q0 = torch.tensor(np.stack([d['q0'][k]] * 2), dtype=torch.float32); v0 = torch.tensor(np.stack([d['v0'][k]] * 2), dtype=torch.float32)
    env.set_initial_state(q0, v0); env.reset()
    sims[prec] = env
  o = OracleSim('so100_handover_banana', collide=True)
  o.set_state(q0[0].double().numpy(), v0[0].double().numpy())
  print(f'=== env {d["env"][k]} (ended at {d["end_step"][k]} in the batch run)')
  for t in range(T):
    a = torch.tensor(d['acts'][t, k]).float().repeat(2, 1).cuda()
    row = [f't={t:3d}']
    for prec, env in sims.items():
      ts = env.step(a)
      q, v = env.get_state(torch.float64)
      row.append(f'{prec}: |v|max {float(v[0].abs().max()):10.3g} ncon {int(env.debug_read("ncon")[0,0]):3d} it {int(env.debug_read("solver_iter")[0,0]):3d} st {int(ts.step_type[0])} armq {q[0,:3].cpu().numpy().round(2).tolist()} banana z {float(q[0,8]):.3f} bowl z {float(q[0,15]):.3f}')
    o.control_step(d['acts'][t, k].astype(np.float64))
    row.append(f'oracle: |v|max {np.abs(o.qvel).max():10.3g} ncon {len(o.contacts()):3d} armq {o.qpos[:3].round(2).tolist()} banana z {o.qpos[8]:.3f} bowl z {o.qpos[15]:.3f}')
    if True: print(' | '.join(row))
