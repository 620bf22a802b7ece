#!/usr/bin/env python3
"""compute-sanitizer target: a few control steps of a 64-env banana scene (collisions, all solver tiers reachable, on-device
placements with nursery envs and auto-resets), of the two-arm build and of the arm-only kernel, through the public API.
usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_probe.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from so101_sim_b200.task_suite import create_batched_task_env
dev = 'cuda:0'
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
def actions(env):
  g = torch.Generator(device=dev); g.manual_seed(1)
  spec = env.action_spec()
  lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
  return lambda: (lo + torch.rand(env.num_envs, env.nu, generator=g, device=dev) * (hi - lo)) * 0.3
# contact scene, host-installed state (eager + graph replay)
env = create_batched_task_env('SO100HandOverBanana', num_envs=64, time_limit=30.0, seed=0, device=dev, reset_rounds=0)
env.sample_prop_initial_states(seed=3, clearance=0.001, settle_steps=0)
a = actions(env)
for t in range(steps):
  ts = env.step(a())
torch.cuda.synchronize()
print('scene ok', env.counters(), float(ts.observation['physics_state'].abs().max()))
env.close()
# on-device placements: sample / reject / settle, nursery ring, auto-reset through a short time limit
env = create_batched_task_env('SO100HandOverBanana', num_envs=16, time_limit=0.04, seed=1, device=dev, nursery_envs=8)
a = actions(env)
for t in range(2 * steps):
  ts = env.step(a())
torch.cuda.synchronize()
print('placements ok', env.placement_stats(), env.counters())
env.close()
# the two-launch narrow phase of large batches (EPA state machine with a pair cursor, manifold kernel), forced on at 64 envs
os.environ['SO101_NARROW_SPLIT'] = '1'
env = create_batched_task_env('SO100HandOverBanana', num_envs=64, time_limit=30.0, seed=0, device=dev, reset_rounds=0)
env.sample_prop_initial_states(seed=3, clearance=0.001, settle_steps=0)
a = actions(env)
for t in range(steps):
  ts = env.step(a())
ck = env.save_checkpoint(); env.load_checkpoint(ck)
torch.cuda.synchronize()
print('two-launch narrow phase + checkpoint ok', env.counters())
env.close()
del os.environ['SO101_NARROW_SPLIT']
# two-arm build
env = create_batched_task_env('SO100TwoArmHandOverBanana', num_envs=16, time_limit=30.0, seed=0, device=dev, nursery_envs=0)
a = actions(env)
for t in range(steps):
  ts = env.step(a())
torch.cuda.synchronize()
print('two-arm ok', env.counters())
env.close()
env = create_batched_task_env('SO100ArmOnly', num_envs=64, time_limit=30.0, seed=0, device=dev)
env.sample_arm_initial_states(seed=0); env.reset()
for t in range(steps):
  ts = env.step(torch.zeros(64, 6, device=dev))
torch.cuda.synchronize()
print('arm ok', env.counters())
env.close()
