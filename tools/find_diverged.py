#!/usr/bin/env python3
"""Developer probe: run a random-action rollout, record the initial state and actions of the first envs that diverge on the
GPU (f32), and save them (gpurun_out/diverged.npz) so that the float64 oracle can replay them on the CPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from so101_sim_b200.task_suite import create_batched_task_env
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
prec = sys.argv[3] if len(sys.argv) > 3 else 'f32'
env = create_batched_task_env('SO100HandOverBanana', num_envs=envs, time_limit=30.0, seed=0, device='cuda:0', precision=prec)
q0, v0 = env.sample_prop_initial_states(seed=0, spawn_z=0.45, settle_steps=50)
q0, v0 = env.get_state(torch.float64)
g = torch.Generator(device='cuda:0'); g.manual_seed(1)
spec = env.action_spec()
lo, hi = torch.tensor(spec.minimum, device='cuda:0'), torch.tensor(spec.maximum, device='cuda:0')
acts, first_last = [], torch.full((envs,), -1, dtype=torch.long, device='cuda:0')
maxv = torch.zeros(envs, device='cuda:0')
for t in range(steps):
  a = (lo + torch.rand(envs, 6, generator=g, device='cuda:0') * (hi - lo)) * 0.3
  acts.append(a.cpu())
  ts = env.step(a)
  q, v = env.get_state()
  maxv = torch.maximum(maxv, v.abs().max(dim=1).values * (first_last < 0))
  newly = (ts.step_type == 2) & (first_last < 0)
  first_last[newly] = t
print('counters', env.counters(), 'envs that ended:', int((first_last >= 0).sum()))
idx = torch.nonzero(first_last >= 0).flatten()[:8].cpu()
fast = torch.topk(maxv, 4).indices.cpu()
sel = torch.unique(torch.cat([idx, fast]))
acts = torch.stack(acts)  # [T, N, 6]
os.makedirs('gpurun_out', exist_ok=True)
np.savez('gpurun_out/diverged.npz', env=sel.numpy(), q0=q0.cpu().numpy()[sel], v0=v0.cpu().numpy()[sel], acts=acts[:, sel].numpy(),
         end_step=first_last.cpu().numpy()[sel], maxv=maxv.cpu().numpy()[sel])
print('saved', sel.tolist(), 'end steps', first_last.cpu().numpy()[sel].tolist(), 'max |qvel| before end', maxv.cpu().numpy()[sel].round(1).tolist())
