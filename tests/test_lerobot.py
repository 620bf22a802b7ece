"""Batched LeRobot adapter + scripted dataset generator (so101_sim_b200/lerobot.py) against the reference's formulas
(scripts/so101_lerobot_wrapper.py:77-122, examples/automated_lerobot_dataset_generator.py:52-260).  The reference modules
import dm_control and cannot be imported here, so the expected values restate its arithmetic in numpy with line citations."""
import math

import numpy as np
import pytest
import torch

from so101_sim_b200.lerobot import (BatchedDatasetGenerator, BatchedSO101LeRobotWrapper, BatchedTrajectoryPlanner, DatasetConfig,
                                    convert_to_lerobot_format)


def test_lerobot_format_matches_the_reference_layout():
  jp = torch.arange(12, dtype=torch.float64).reshape(2, 6)
  o = convert_to_lerobot_format(jp, None, frame_index=0, episode_index=3, device='cpu')
  # key set of so101_lerobot_wrapper.py:77-122 without the camera keys (state-only)
  assert set(o) == {'observation.state', 'action', 'timestamp', 'frame_index', 'episode_index', 'index', 'task_index', 'task'}
  assert o['observation.state'].dtype == torch.float32 and o['observation.state'].shape == (2, 6)
  assert torch.equal(o['action'], torch.zeros(2, 6))  # zero action echoed on reset (:107-110)
  a = torch.full((2, 6), 0.25)
  o = convert_to_lerobot_format(jp, a, frame_index=7, episode_index=3, device='cpu')
  assert torch.equal(o['action'], a)
  np.testing.assert_allclose(o['timestamp'].numpy(), np.float32(7 * 0.1))  # (sic) frame_index * 0.1 (:113)
  assert o['frame_index'].tolist() == [7, 7] and o['index'].tolist() == [7, 7] and o['episode_index'].tolist() == [3, 3]
  assert o['task_index'].tolist() == [0, 0] and o['frame_index'].dtype == torch.long and o['task'] == 'SO100 manipulation task'


def _ref_ik(p, offsets):  # automated_lerobot_dataset_generator.py:188-212
  x, y, z = p
  base = np.arctan2(x, y)
  h, v = np.sqrt(x * x + y * y), z - 0.42
  sh = -np.arctan2(v, h) - 0.5
  el = np.pi / 2 + np.arctan2(v, h)
  return np.array([base, sh, el, np.pi / 2 - sh - el, 0.0, 0.0]) + offsets


def test_planner_matches_the_reference_geometry():
  cfg = DatasetConfig()
  offs = np.array([28, 42, 18, -21, 1009, -158], dtype=np.float64)  # calibration/red_arm.json homing offsets
  g = torch.Generator(); g.manual_seed(0)
  pl = BatchedTrajectoryPlanner(cfg, offs, g, 'cpu')
  banana, bowl = pl.randomize_spawn_positions(64)
  # :57-78 spawn arcs: banana 0.15..0.25 m inside +-60 degrees of +y, bowl 0.12..0.30 m; both at table height
  d = banana[:, :2].norm(dim=1)
  assert float(d.min()) >= 0.15 - 1e-6 and float(d.max()) <= cfg.banana_spawn_radius + 1e-6
  assert float(torch.atan2(banana[:, 0], banana[:, 1]).abs().max()) <= math.pi / 3 + 1e-6
  db = bowl[:, :2].norm(dim=1)
  assert float(db.min()) >= 0.12 - 1e-6 and float(db.max()) <= cfg.bowl_spawn_radius + 1e-6
  assert torch.all(banana[:, 2] == 0.42) and torch.all(bowl[:, 2] == 0.42)
  start = pl.randomize_robot_start_pose(64)
  lo = np.array([l for l, _ in pl.JOINT_LIMITS]); hi = np.array([h for _, h in pl.JOINT_LIMITS])
  assert np.all(start.numpy() >= lo - 1e-6) and np.all(start.numpy() <= hi + 1e-6)
  assert np.abs(start.numpy() - np.clip(np.array(pl.BASE_POSE), lo, hi)).max() <= cfg.robot_pose_variation + 1e-6
  ik = pl.inverse_kinematics_approximate(banana + torch.tensor([0.0, 0.0, 0.1]))
  for e in range(8):
    np.testing.assert_allclose(ik[e].numpy(), _ref_ik(banana[e].numpy() + [0, 0, 0.1], offs), rtol=1e-5, atol=1e-4)
  pick = pl.plan_pickup_trajectory(start, banana)
  place = pl.plan_placement_trajectory(pick[-1], bowl)
  assert len(pick) == 8 + 5 + 3 + 5 and len(place) == 8 + 4 + 3 + 5  # :124-143, :156-184
  # each segment ends exactly on its target (the noise fades with 1 - smoothstep, :222-224)
  lift = _ref_ik(banana[0].numpy() + [0, 0, cfg.lift_height], offs); lift[5] = 0.05 + offs[5] * 0  # gripper set after calibration (:139-141)
  got = pick[-1][0].numpy()
  np.testing.assert_allclose(got[:5], lift[:5], rtol=1e-5, atol=1e-4)
  assert got[5] == np.float32(0.05)
  assert place[-1][0, 5] == 0.0 and place[11][0, 5] == np.float32(0.05)  # released / still closed
  # intermediate waypoints stay within the noise envelope of the smooth-step path
  t = 1 / 8
  st = t * t * (3 - 2 * t)
  target = pl.inverse_kinematics_approximate(banana + torch.tensor([0.0, 0.0, cfg.approach_height]))
  dev = (pick[0] - (start + st * (target - start))).abs()
  assert float(dev[:, :5].max()) < 0.02 * 6 and float(dev[:, 5].max()) == 0.0
  # grasp-failure test (:238-241): distance of the first three joint values to the banana position, or an open gripper
  f = pl.detect_grasp_failure(torch.tensor([[0.0, 0.2, 0.42], [0.0, 0.2, 0.42], [0.5, 0.2, 0.42]]), torch.tensor([[0.0, 0.2, 0.42]] * 3),
                              torch.tensor([0.05, 0.0, 0.05]))
  assert f.tolist() == [False, True, True]


class _FakeEnv:
  """Stands in for BatchedEnvironment on CPU: joints_pos follows the action with a one-step lag."""
  def __init__(self, n):
    self.num_envs, self.nq = n, 20
    self.q = torch.zeros(n, 6)

  def reset(self):
    from so101_sim_b200.task_suite import BatchedTimeStep
    self.q = torch.zeros(self.num_envs, 6)
    return BatchedTimeStep(torch.zeros(self.num_envs, dtype=torch.uint8), torch.zeros(self.num_envs), torch.ones(self.num_envs), {'joints_pos': self.q})

  def step(self, a):
    from so101_sim_b200.task_suite import BatchedTimeStep
    out = BatchedTimeStep(torch.ones(self.num_envs, dtype=torch.uint8), torch.zeros(self.num_envs), torch.ones(self.num_envs), {'joints_pos': self.q})
    self.q = a.clone()
    return out


def test_generator_episode_records_on_a_fake_env():
  n = 16
  w = BatchedSO101LeRobotWrapper.__new__(BatchedSO101LeRobotWrapper)
  w.device, w.cameras, w.env, w.num_envs, w.episode_index, w.frame_index = 'cpu', (), _FakeEnv(n), n, 0, 0
  gen = BatchedDatasetGenerator(DatasetConfig(), w, seed=3)
  ep = gen.generate_episodes()
  T = ep['actions'].shape[0]
  assert T == 50  # 21 + 20 + settling, capped by max_episode_length (:392-445)
  assert ep['observations']['state'].shape == (T + 1, n, 6) and ep['rewards'].shape == (T, n) and ep['dones'].shape == (T + 1, n)
  ok = ep['episode_metadata']['success']
  L = ep['length']
  assert torch.all(L[ok] == T) and torch.all(L[~ok] == 21)
  assert torch.equal(ep['rewards'].sum(0), ok.float())             # a single final reward of 1 for successes (:446-449)
  assert torch.all(ep['dones'].sum(0) == 1) and torch.all(ep['dones'][L, torch.arange(n)])
  assert ep['episode_metadata']['failure_reason'] == ['grasp_failed' if not s else None for s in ok.tolist()]
  assert w.frame_index == 1 + T and w.episode_index == 1
  # the wrapper echoes the action and reports the (lagged) joint observation
  o = w.step(torch.ones(6))
  assert o['action'].shape == (n, 6) and torch.equal(o['observation.state'], ep['actions'][-1])


@pytest.mark.gpu
def test_wrapper_and_generator_on_the_gpu(built):
  w = BatchedSO101LeRobotWrapper(num_envs=8, device='cuda:0', seed=0, reset_rounds=2)
  o = w.reset()
  assert o['observation.state'].shape == (8, 6) and o['observation.state'].device.type == 'cuda'
  assert float(o['observation.state'].abs().max()) == 0.0  # arm resets to qpos = 0 (so100_task.py:308-313)
  ep = BatchedDatasetGenerator(DatasetConfig(), w, seed=1).generate_episodes()
  assert ep['actions'].shape == (50, 8, 6) and torch.isfinite(ep['observations']['state']).all()
  # observation.state is joints_pos delayed by 5 control steps (so100_task.py:196-198): the arm starts at 0, so the first 5
  # recorded states after the reset are still 0 while the arm is already moving
  assert float(ep['observations']['state'][:5].abs().max()) == 0.0 and float(ep['observations']['state'][8].abs().max()) > 0.0
  w.env.close()


@pytest.mark.gpu
def test_reset_pool_cycles_placements(built):
  """Every reset of an env moves on to its next sampled-and-settled placement (initialize_episode re-samples the props in the
  reference, so100_hand_over.py:208-229,320-323); the pool wraps around."""
  from so101_sim_b200.task_suite import create_batched_task_env
  env = create_batched_task_env('SO100HandOverBanana', num_envs=4, time_limit=0.04, seed=0, device='cuda:0', reset_rounds=0)
  Q, V = env.randomize_resets(rounds=3, seed=7, settle_steps=10)
  assert Q.shape == (3, 4, 20) and not torch.allclose(Q[0, :, 6:8], Q[1, :, 6:8])
  # placement ranges of so100_hand_over.py:37-55
  assert float(Q[..., 6].min()) >= 0.2 - 5e-3 and float(Q[..., 6].max()) <= 0.3 + 5e-3 and float(Q[..., 13].min()) >= -0.3 - 5e-3
  seen = []
  zero = torch.zeros(4, 6, device='cuda:0')
  ts = env.reset()  # randomize_resets() already consumed entry 0 with its own reset(): this one starts episode 1
  for _ in range(4):
    seen.append(ts.observation['physics_state'][:, :20].clone())
    assert ts.step_type.tolist() == [0] * 4
    ts = env.step(zero); ts = env.step(zero)
    assert ts.step_type.tolist() == [2] * 4  # time limit 0.04 s = 2 control steps... LAST
    ts = env.step(zero)                      # auto-reset -> FIRST, next pool entry
  for k, s in enumerate(seen):
    assert torch.allclose(s, Q[(k + 1) % 3], atol=1e-6), k
  env.close()
