#!/usr/bin/env python3
"""Developer probe: which buffer drops contacts, how many envs diverge, contact-count histogram over a long random-action rollout."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so101_sim_b200.task_suite import create_batched_task_env
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 120
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
env = create_batched_task_env('SO100HandOverBanana', num_envs=envs, time_limit=30.0, seed=0, device='cuda:0')
env.sample_prop_initial_states(seed=0, spawn_z=0.45, settle_steps=50)
def cat():
  out = torch.empty(8, dtype=torch.float32, device='cuda:0')
  env._check(env._lib.so101_debug_read(env._h, b'dropcat', ctypes.c_void_p(out.data_ptr()), 8, env._stream()))
  return [int(x) for x in out.tolist()]
print('after settle:', dict(zip(['NOUT', 'CANDCAP', 'PAIRCAP', 'QUEUE', 'CONBUF', 'NBLK'], cat())), env.counters())
g = torch.Generator(device='cuda:0'); g.manual_seed(1)
spec = env.action_spec()
lo, hi = torch.tensor(spec.minimum, device='cuda:0'), torch.tensor(spec.maximum, device='cuda:0')
for t in range(steps):
  a = (lo + torch.rand(envs, 6, generator=g, device='cuda:0') * (hi - lo)) * scale
  ts = env.step(a)
  if t % 20 == 19:
    ncon = env.debug_read('ncon').flatten()
    h = torch.histc(ncon.float(), bins=13, min=0, max=104)
    q, v = env.get_state()
    print(t, dict(zip(['NOUT', 'CANDCAP', 'PAIRCAP', 'QUEUE', 'CONBUF', 'NBLK'], cat())), env.counters(), 'ncon hist/8:', [int(x) for x in h.tolist()],
          'last:', int((ts.step_type == 2).sum()), 'min arm z-ish q1:', float(q[:, 1].min()), 'max |qvel|', float(v.abs().max()))
