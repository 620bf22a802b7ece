"""On-device episode initialisation: the batched twin of SO100HandOver.initialize_episode
(so100_hand_over.py:320-323: three PropPlacers :208-229 with the distributions :37-55, arm reset so100_task.py:304-320) -
Philox sampling, collision rejection with the real narrow phase, frozen-arm settle, and a nursery of hidden envs that keeps
producing fresh settled placements for the auto-resets."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
HOME = [0.0, -1.57079, 1.57079, 1.57079, -1.57079, 0.0]


def _make(n, **kw):
  from so101_sim_b200.task_suite import create_batched_task_env
  args = dict(task_name='SO100HandOverBanana', num_envs=n, time_limit=30.0, seed=11, device=DEV)
  args.update(kw)
  return create_batched_task_env(**args)


def _yaw(q):  # rotation angle about z of a (w, x, y, z) quaternion batch
  return 2 * torch.atan2(q[:, 3], q[:, 0])


def test_reset_state_follows_the_reference_placement_semantics(built):
  """KAT-2 semantics (examples/so101_rl_breakdown.ipynb:274-298) for a whole batch: commanded = HOME_CTRL, arm qpos = 0 and arm
  qvel = 0 (the arm is FROZEN while the props settle), props inside their placement boxes (so100_hand_over.py:37-55) up to the
  settle drift, resting on the table (or, for a container that landed on the static cylinder obstacle, on it: KAT-2's
  z = 0.4306), object yaw within +-0.1 pi, residual prop velocities below the settle tolerance unless the 2 s ran out."""
  env = _make(256, nursery_envs=0)
  st = env.placement_stats()
  ts = env.reset()
  ps = ts.observation['physics_state'].double()
  assert ts.step_type.eq(0).all() and ts.reward.eq(0).all() and ts.discount.eq(1).all()
  np.testing.assert_allclose(ts.observation['commanded_joints_pos'][0].cpu().numpy(), HOME, atol=1e-5)
  assert float(ps[:, :6].abs().max()) == 0.0 and float(ps[:, 20:26].abs().max()) == 0.0
  assert float(ts.observation['joints_pos'].abs().max()) == 0.0
  bx, by, bz = ps[:, 6], ps[:, 7], ps[:, 8]
  cx, cy, cz = ps[:, 13], ps[:, 14], ps[:, 15]
  tol = 0.06   # props roll / slide while settling from the 3 cm drop (an object that lands on the static capsule obstacle rolls off it)
  assert float(bx.min()) > 0.2 - tol and float(bx.max()) < 0.3 + tol and float(by.abs().max()) < 0.1 + tol
  assert float(cx.min()) > -0.3 - tol and float(cx.max()) < -0.2 + tol and float(cy.abs().max()) < 0.1 + tol
  assert float(bz.min()) > 0.418 and float(bz.max()) < 0.48 and float(cz.min()) > 0.420 and float(cz.max()) < 0.45   # (an object may end up leaning on the static capsule, top at 0.46)
  assert abs(float(bz.median()) - 0.4217) < 1e-3                      # most objects rest on the table at the reference's rest height (so101_rl.ipynb:221)
  assert float((cz - 0.4226).abs().min()) < 1e-3                      # most containers rest flat on the table ...
  yaw = _yaw(ps[:, 9:13]).abs()
  assert 0.2 < float(yaw.max()) and float(yaw.median()) < 0.1 * np.pi     # ... and the objects keep (roughly) their sampled yaw (one that rolls off the capsule turns)
  spread = float(bx.max() - bx.min())
  assert spread > 0.07                                                 # the whole 10 cm box is used
  print('placement stats after create:', st, '| residual prop speed max', float(ps[:, 26:].abs().max()))
  assert st['attempts_exhausted'] == 0
  env.close()


def test_placements_are_a_function_of_the_seed(built):
  a, b, c = _make(32, nursery_envs=0), _make(32, nursery_envs=0), _make(32, nursery_envs=0, seed=12)
  pa, pb, pc = (e.reset().observation['physics_state'].clone() for e in (a, b, c))
  assert torch.equal(pa, pb)                           # same seed -> same sampled and settled states
  assert float((pa[:, 6:8] - pc[:, 6:8]).abs().min()) > 0   # another seed -> other placements
  assert float((pa[0, 6:8] - pa[1, 6:8]).abs().max()) > 0   # envs draw independently
  # a second initialize_placements() draws new placements
  a.initialize_placements()
  assert float((a.reset().observation['physics_state'][:, 6:8] - pa[:, 6:8]).abs().min()) > 0
  for e in (a, b, c):
    e.close()


def test_container_spawned_inside_the_obstacle_is_resampled(built):
  """[upstream] PropPlacer(ignore_collisions=False) for the container (so100_hand_over.py:216-221): a spawn pose that penetrates
  the static cylinder obstacle is rejected and re-drawn.  Forced here by lowering the container's spawn box INTO the obstacle
  (z = 0.43 < the cylinder's top at 0.45) over its footprint: every first sample collides; with the box widened to the whole
  placement range the sampler finds the free part."""
  from so101_sim_b200.task_suite import BatchedEnvironment, SO100HandOver
  task = SO100HandOver('banana', control_timestep=0.02, cameras=())
  task.PLACE_LO = (task.PLACE_LO[0], (-0.3, -0.1, 0.43)); task.PLACE_HI = (task.PLACE_HI[0], (-0.2, 0.1, 0.43))
  env = BatchedEnvironment(task, 64, 30.0, 5, DEV, None, 'f32', 100, 1e-6)
  st = env.initialize_placements()
  assert st['rejected_samples'] > 0 and st['attempts_exhausted'] == 0
  ps = env.reset().observation['physics_state']
  # the obstacle sits at (-0.2, 0.1) with radius 0.08: no settled container centre is left inside it
  d = torch.hypot(ps[:, 13] + 0.2, ps[:, 14] - 0.1)
  print('rejected samples:', st['rejected_samples'], 'min distance of a container origin from the obstacle axis:', float(d.min()))
  assert float(ps[:, 15].max()) < 0.43 + 1e-3
  env.close()


def test_every_episode_starts_from_a_fresh_placement(built):
  """Auto-resets draw from the nursery's ring: consecutive episodes of an env start from different, never repeated, settled
  placements (the reference re-samples in every initialize_episode); the arm always restarts at qpos 0 with the home command."""
  env = _make(4, time_limit=0.6, nursery_envs=48)   # 31-step episodes; 48 nursery envs supply ~0.5 placements per control step
  zero = torch.zeros(4, 6, device=DEV)
  ts = env.reset()
  firsts = [ts.observation['physics_state'].clone()]
  for _ in range(150):     # let the nursery fill its ring
    ts = env.step(zero)
    if int(ts.step_type[0]) == 0:
      firsts.append(ts.observation['physics_state'].clone())
  st = env.placement_stats()
  print('nursery:', st, 'episodes seen:', len(firsts))
  assert len(firsts) >= 4 and st['published'] >= 8
  late = firsts[-2:]
  assert float((late[0][:, 6:8] - late[1][:, 6:8]).abs().min()) > 1e-4      # fresh placement per episode
  for f in firsts:
    assert float(f[:, :6].abs().max()) == 0.0 and float(f[:, 20:26].abs().max()) == 0.0
    assert float(f[:, 6].min()) > 0.17 and float(f[:, 6].max()) < 0.33 and float(f[:, 13].min()) > -0.33 and float(f[:, 13].max()) < -0.17
  assert st['consumed'] >= 4 and env.counters()['diverged'] == 0
  env.close()
