#!/usr/bin/env python3
"""ncu target: BASELINE config 3 (SO100HandOverBanana, 16384 envs, device placements) in the random-action regime.  The profiled
region (cudaProfilerStart/Stop -> run ncu with --profile-from-start off) is `steps` control steps after `warmup` random-action
steps, launched eagerly (SO101_GRAPH=0) so that every kernel is an individual launch.
usage: ncu --profile-from-start off ... python tools/ncu_target.py [envs] [warmup] [steps]"""
import os, sys
os.environ.setdefault('SO101_GRAPH', '0')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so101_sim_b200.task_suite import create_batched_task_env
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 30
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = 'cuda:0'
env = create_batched_task_env('SO100HandOverBanana', num_envs=envs, time_limit=30.0, seed=0, device=dev, placement='device', nursery_envs=0)
env.reset()
g = torch.Generator(device=dev); g.manual_seed(1)
spec = env.action_spec()
lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
acts = (lo + torch.rand(warmup + steps, envs, 6, generator=g, device=dev) * (hi - lo)) * 0.3
for t in range(warmup):
  env.step(acts[t])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for t in range(steps):
  env.step(acts[warmup + t])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled', steps, 'control steps;', env.counters())
