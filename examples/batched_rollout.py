#!/usr/bin/env python3
"""Minimal use of the batched drop-in (needs a B200 and the built library: python -c "import __graft_entry__ as g; g.build()").

The reference's single-env loop (examples/so101_rl_breakdown.ipynb:425-440)
    env = task_suite.create_task_env('SO100HandOverBanana', time_limit=30.0, cameras=())
    ts = env.reset()
    for _ in range(steps): ts = env.step(action)           # numpy (6,) -> dm_env.TimeStep
becomes, for N environments in lockstep on one GPU:"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from so101_sim_b200.lerobot import BatchedDatasetGenerator, BatchedSO101LeRobotWrapper, DatasetConfig  # noqa: E402
from so101_sim_b200.task_suite import create_batched_task_env  # noqa: E402

N, dev = 4096, 'cuda:0'
env = create_batched_task_env('SO100HandOverBanana', num_envs=N, time_limit=30.0, seed=0, cameras=(), device=dev)
env.randomize_resets(rounds=4)                  # 4 sampled-and-settled prop placements per env; auto-resets cycle through them
ts = env.reset()                                # BatchedTimeStep: step_type u8[N], reward f32[N], discount f32[N], observation dict
spec = env.action_spec()
lo, hi = torch.tensor(spec.minimum, device=dev), torch.tensor(spec.maximum, device=dev)
ret = torch.zeros(N, device=dev)
for t in range(200):
  action = (lo + torch.rand(N, 6, device=dev) * (hi - lo)) * 0.3
  ts = env.step(action)                         # float32 [N, 6] on the device; no host sync inside
  ret += ts.reward
print('observation keys:', list(ts.observation.keys()))
print('mean return after 200 steps:', float(ret.mean()), '| counters:', env.counters())
env.close()

# LeRobot-style records from the scripted pick-and-place generator (twin of examples/automated_lerobot_dataset_generator.py)
w = BatchedSO101LeRobotWrapper(num_envs=256, device=dev, seed=0, reset_rounds=2)
ep = BatchedDatasetGenerator(DatasetConfig(), w, seed=0).generate_episodes()
print('episode batch: actions', tuple(ep['actions'].shape), 'states', tuple(ep['observations']['state'].shape),
      'scripted successes', int(ep['episode_metadata']['success'].sum()), 'of', w.num_envs)
w.env.close()
