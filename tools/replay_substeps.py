#!/usr/bin/env python3
"""Developer probe: replay one saved env substep by substep in f32 (control_timestep = physics timestep) and print what happens
around the first velocity jump."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so101_sim_b200.task_suite import create_batched_task_env
d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'diverged.npz'))
want = int(sys.argv[1]) if len(sys.argv) > 1 else 739
prec = sys.argv[2] if len(sys.argv) > 2 else 'f32'
k = [i for i, e in enumerate(d['env']) if int(e) == want][0]
env = create_batched_task_env('SO100HandOverBanana', num_envs=2, time_limit=30.0, seed=0, device='cuda:0', precision=prec, control_timestep=0.002)
q0 = torch.tensor(np.stack([d['q0'][k]] * 2), dtype=torch.float32); v0 = torch.tensor(np.stack([d['v0'][k]] * 2), dtype=torch.float32)
env.set_initial_state(q0, v0); env.reset(); env.debug_contacts()
T = int(d['end_step'][k]) + 1
hist = []
for t in range(T):
  a = torch.tensor(d['acts'][t, k]).float().repeat(2, 1).cuda()
  for sub in range(10):
    ts = env.step(a)
    q, v = env.get_state(torch.float64)
    con = env.debug_contacts()[0]
    hist.append((t, sub, float(v[0].abs().max()), v[0].cpu().numpy().copy(), q[0].cpu().numpy().copy(), int(env.debug_read('solver_iter')[0, 0]), con, int(ts.step_type[0])))
vm = np.array([h[2] for h in hist])
jump = next((i for i in range(1, len(vm)) if vm[i] > 3 * max(vm[max(0, i - 20):i].max(), 50.0)), len(vm) - 1)
print('first jump at substep index', jump, 'control step', hist[jump][0], 'sub', hist[jump][1])
for i in range(max(0, jump - 4), min(len(hist), jump + 3)):
  t, sub, m, v, q, it, con, st = hist[i]
  print(f'--- t={t} sub={sub} |v|max={m:.4g} iters={it} st={st} ncon={len(con)}')
  print('   qvel arm', v[:6].round(2).tolist(), 'banana', v[6:12].round(2).tolist(), 'bowl', v[12:].round(2).tolist())
  print('   banana pos', q[6:9].round(4).tolist(), 'bowl pos', q[13:16].round(4).tolist(), 'arm q', q[:6].round(3).tolist())
  for c in con: print(f'     geoms ({c[0]},{c[1]}) dist {c[2]:.3e} pos {c[3].round(4).tolist()} n {c[4].round(3).tolist()}')
